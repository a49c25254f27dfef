"""Attribute-view CNN score (MultiKE_model.py:34-63) + loss + full backward on the GPU against the
torch-autograd oracle (oracle/attr_cnn.py, float64), incl. the global batch normalisation, the
three weighting modes of the reference graphs and the Adagrad updates of tables and parameters."""
import numpy as np
import pytest
import torch

from oracle import attr_cnn as oc
from oracle.tf_semantics import l2_normalize

pytestmark = pytest.mark.gpu


# B = 5000 is args.json's attribute_batch_size: the 8-way split-batch reduction of the parameter gradient
# (n >= 2048) and the multi-iteration grid-stride passes only run at that size
@pytest.mark.parametrize("dim,B,weighted,scale", [(75, 257, True, 1.0), (75, 64, False, 2.0), (32, 50, True, 1.0),
                                                  (75, 5000, True, 1.0), (75, 5000, False, 2.0), (75, 2048, True, 1.0)])
def test_attr_cnn_step_matches_oracle(dim, B, weighted, scale):
    from multike_b200 import tables as T
    from multike_b200.attr_view import AttrCNN, param_layout
    rng = np.random.default_rng(dim + B)
    n_ent, n_attr, n_val = (400, 30, 300) if B < 1000 else (20000, 300, 6000)
    ent0 = rng.normal(0, 0.02, (n_ent, dim))
    att0 = rng.normal(0, 0.05, (n_attr, dim))
    val0 = rng.normal(0, 0.3, (n_val, dim))
    gen = torch.Generator().manual_seed(3)
    theta0 = oc.init_theta(dim, generator=gen)
    lay = oc.layout(dim)
    assert {k: v for k, v in lay.items()} == {k: v for k, v in param_layout(dim).items()}
    # non-trivial gamma/beta/biases so that every parameter gradient is exercised
    theta0[:2 * dim] += torch.tensor(rng.normal(0, 0.1, 2 * dim))
    theta0[lay["b1"][0]:lay["b1"][0] + 2] = torch.tensor([0.05, -0.03])
    theta0[lay["b2"][0]:lay["b2"][0] + 2] = torch.tensor([-0.02, 0.04])
    theta0[lay["bd"][0]:] = torch.tensor(rng.normal(0, 0.05, dim))
    ih, ia, iv = rng.integers(0, n_ent, B), rng.integers(0, n_attr, B), rng.integers(0, n_val, B)
    w = rng.choice([1.0, 0.9, 0.6], B) if weighted else None

    ent = T.EmbeddingTable(n_ent, dim, True, "cuda", init=ent0, flags=True, grad_replicas=1)
    att = T.EmbeddingTable(n_attr, dim, False, "cuda", init=att0)          # small table: replicas, no flags
    val = T.EmbeddingTable(n_val, dim, False, "cuda", init=val0, trainable=False)
    cnn = AttrCNN(dim, theta=theta0.float())
    acc = T.new_loss_accumulator()
    cnn.fwd_bwd(ent, att, val, ih, ia, iv, acc, w=w, scale=scale)
    torch.cuda.synchronize()

    Ve, Va = torch.tensor(ent0, requires_grad=True), torch.tensor(att0, requires_grad=True)
    th = theta0.clone().requires_grad_(True)
    E = l2_normalize(Ve, 1)
    Ed = E.detach().clone().requires_grad_(True)
    loss = oc.attribute_cnn_loss(Ed[ih], Va[ia], torch.tensor(val0)[iv], None if w is None else torch.tensor(w), th, dim,
                                 scale=scale)
    gE, gA, gT = torch.autograd.grad(loss, [Ed, Va, th])
    assert float(acc.item()) == pytest.approx(float(loss), rel=2e-5)
    scale_t = float(gT.abs().max())
    np.testing.assert_allclose(cnn.grad.double().cpu().numpy(), gT.numpy(), rtol=2e-3, atol=2e-5 * max(scale_t, 1.0))
    np.testing.assert_allclose(ent.grad_sum()[:, :dim].double().cpu().numpy(), gE.numpy(), rtol=1e-3, atol=2e-6)
    np.testing.assert_allclose(att.grad_sum()[:, :dim].double().cpu().numpy(), gA.numpy(), rtol=2e-3, atol=2e-6)
    # per-layer view of the parameter gradient (helps when something is off)
    for name in ("gamma", "beta", "k1", "b1", "k2", "b2", "wd", "bd"):
        off, shape = lay[name]
        n = int(np.prod(shape))
        got, want = cnn.grad[off:off + n].double().cpu().numpy(), gT[off:off + n].numpy()
        assert np.abs(got - want).max() <= 2e-3 * np.abs(want).max() + 1e-6, name
    # phase 2: tables (normalise-backward for ent, raw for attr) and the dense parameters
    lr = 0.001
    ent.apply_adagrad("attribute", lr)
    att.apply_adagrad("attribute", lr)
    cnn.apply_adagrad("attribute", lr)
    loss_full = oc.attribute_cnn_loss(E[ih], Va[ia], torch.tensor(val0)[iv], None if w is None else torch.tensor(w), th, dim,
                                      scale=scale)
    gV, = torch.autograd.grad(loss_full, [Ve])
    np.testing.assert_allclose(ent.raw(), ent0 - lr * gV.numpy() / np.sqrt(0.1 + gV.numpy() ** 2), rtol=0, atol=5e-6)
    np.testing.assert_allclose(att.raw(), att0 - lr * gA.numpy() / np.sqrt(0.1 + gA.numpy() ** 2), rtol=0, atol=5e-6)
    want_t = theta0.numpy() - lr * gT.numpy() / np.sqrt(0.1 + gT.numpy() ** 2)
    np.testing.assert_allclose(cnn.theta.cpu().numpy(), want_t, rtol=0, atol=5e-6)
    assert float(cnn.grad.abs().max()) == 0.0 and float(ent.grad.abs().max()) == 0.0
