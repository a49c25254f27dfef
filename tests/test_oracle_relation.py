"""The dense (autograd-through-the-whole-table) oracle against (i) its frozen float64 answers and
(ii) the hand-derived sparse form the CUDA kernels implement."""
import numpy as np
import pytest
import torch

from oracle import losses as ol
from oracle import relation_view as orv
from oracle.tf_semantics import ADAGRAD_INIT, l2_normalize


@pytest.mark.parametrize("fname", ["relation_step_d75.npz", "relation_step_d128.npz"])
def test_fp32_oracle_reproduces_fp64_golden(golden, fname):
    g = golden(fname)
    pos, neg, lr = g["pos"], g["neg"], float(g["lr"])
    ent, rel = orv.DenseTable(g["ent0"], True, torch.float32), orv.DenseTable(g["rel0"], True, torch.float32)
    loss, ge, gr = orv.relation_view_step(ent, rel, pos[:, 0], pos[:, 1], pos[:, 2], neg[:, 0], neg[:, 1], neg[:, 2], lr)
    assert loss == pytest.approx(float(g["loss"]), rel=1e-5)
    # row 5 sits on the 1e-12 clamp: its gradient is amplified by 1e6, keep the check relative
    scale = np.abs(g["grad_ent"]).max(axis=1, keepdims=True) + 1e-30
    assert np.max(np.abs(ge.numpy() - g["grad_ent"]) / scale) < 2e-3
    keep = np.ones(len(g["ent0"]), bool)
    keep[5] = False
    np.testing.assert_allclose(ent.var.numpy()[keep], g["ent1"][keep], rtol=0, atol=2e-6)
    np.testing.assert_allclose(rel.var.numpy(), g["rel1"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("fname", ["relation_step_d75.npz", "relation_step_d128.npz"])
def test_sparse_form_equals_dense_autograd(golden, fname):
    """scatter-add of +-c*d rows, then per-row (g - u(u.g))/|v|, then Adagrad on touched rows only
    == dense TF semantics (float64, tight)."""
    g = golden(fname)
    pos, neg_ent, side, K = g["pos"], g["neg_ent"], g["neg_side"], int(g["K"])
    V, R = torch.tensor(g["ent0"]), torch.tensor(g["rel0"])
    E, Rn = l2_normalize(V, 1), l2_normalize(R, 1)
    G, Gr = torch.zeros_like(V), torch.zeros_like(R)
    loss = 0.0
    for i, (h, r, t) in enumerate(pos):
        d = E[h] + Rn[r] - E[t]
        s = (d * d).sum()
        loss += torch.log(1 + torch.exp(s))
        c = 2 * torch.sigmoid(s)
        G[h] += c * d; Gr[r] += c * d; G[t] -= c * d
        for j in range(K):
            e = int(neg_ent[i, j])
            hs = (int(side[i]) >> j) & 1
            d = (E[e] + Rn[r] - E[t]) if hs else (E[h] + Rn[r] - E[e])
            s = (d * d).sum()
            loss += torch.log(1 + torch.exp(-s))
            c = -2 * torch.sigmoid(-s)
            Gr[r] += c * d
            if hs:
                G[e] += c * d; G[t] -= c * d
            else:
                G[h] += c * d; G[e] -= c * d
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-12)
    np.testing.assert_allclose(G.numpy(), g["view_grad_ent"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(Gr.numpy(), g["view_grad_rel"], rtol=1e-10, atol=1e-13)

    def project(v, gv):
        ss = (v * v).sum(1, keepdim=True)
        inv = torch.rsqrt(torch.clamp(ss, min=1e-12))
        coef = torch.where(ss >= 1e-12, (v * gv).sum(1, keepdim=True) * inv * inv, torch.zeros_like(ss))
        return (gv - v * coef) * inv

    gV, gRv = project(V, G), project(R, Gr)
    np.testing.assert_allclose(gV.numpy(), g["grad_ent"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(gRv.numpy(), g["grad_rel"], rtol=1e-8, atol=1e-12)
    untouched = (G.abs().sum(1) == 0)
    assert untouched.any() and float(gV[untouched].abs().max()) == 0.0      # => Adagrad no-op there
    acc = torch.full_like(V, ADAGRAD_INIT) + gV * gV
    V1 = V - float(g["lr"]) * gV * torch.rsqrt(acc)
    np.testing.assert_allclose(V1.numpy(), g["ent1"], rtol=0, atol=1e-14)
    assert np.array_equal(V1.numpy()[untouched.numpy()], g["ent0"][untouched.numpy()])


def test_losses_known_answers():
    """Forward known-answer tests for losses.py (hand-computed)."""
    z = torch.zeros(2, 3)
    one = torch.tensor([[1.0, 0, 0], [0, 2.0, 0]])
    # pos distance = one  -> ||.||^2 = (1, 4); neg distance = 0
    got = ol.relation_logistic_loss(one, z, z, z, z, z)
    want = np.log(1 + np.exp(1.0)) + np.log(1 + np.exp(4.0)) + 2 * np.log(2.0)
    assert float(got) == pytest.approx(want, rel=1e-6)
    w = torch.tensor([0.5, 2.0])
    got = ol.attribute_logistic_loss(one, z, z, w, one, z, z, w)
    want = 0.5 * np.log(1 + np.exp(1.0)) + 2 * np.log(1 + np.exp(4.0)) + 0.5 * np.log(1 + np.exp(-1.0)) + 2 * np.log(
        1 + np.exp(-4.0))
    assert float(got) == pytest.approx(want, rel=1e-6)
    assert float(ol.relation_logistic_loss_wo_negs(one, z, z)) == pytest.approx(
        np.log(1 + np.exp(1.0)) + np.log(1 + np.exp(4.0)), rel=1e-6)
    assert float(ol.attribute_logistic_loss_wo_negs(one, z, z)) == float(ol.relation_logistic_loss_wo_negs(one, z, z))
    assert float(ol.logistic_loss_wo_negs(one, z, z, w)) == pytest.approx(
        0.5 * np.log(1 + np.exp(1.0)) + 2 * np.log(1 + np.exp(4.0)), rel=1e-6)
    assert float(ol.alignment_loss(one, z)) == pytest.approx(5.0)
    eye = torch.eye(3)
    assert float(ol.orthogonal_loss(2 * eye, eye)) == pytest.approx(27.0)
    # space mapping with identity map: global l2 normalisation divides by sqrt(5)
    got = ol.space_mapping_loss(one, z, eye, eye, 2.0)
    assert float(got) == pytest.approx(1.0 + 0.0 + 1e-4 * 3.0, rel=1e-6)


def test_l2_normalize_clamp_and_axes():
    x = torch.tensor([[3.0, 4.0], [0.0, 0.0]])
    y = l2_normalize(x, 1)
    assert torch.allclose(y[0], torch.tensor([0.6, 0.8])) and float(y[1].abs().sum()) == 0.0
    assert torch.allclose(l2_normalize(x), x / 5.0)
    tiny = torch.tensor([[1e-9, 0.0]], dtype=torch.float64, requires_grad=True)
    out = l2_normalize(tiny, 1)
    assert float(out[0, 0]) == pytest.approx(1e-3)          # x * rsqrt(1e-12)
    out.sum().backward()
    assert float(tiny.grad[0, 1]) == pytest.approx(1e6)     # no projection below the clamp


def test_structured_roundtrip(golden):
    g = golden("relation_step_d75.npz")
    K = int(g["K"])
    neg = orv.structured_to_negatives(g["pos"], g["neg_ent"], g["neg_side"], K)
    assert np.array_equal(neg, g["neg"])
    ent, side = orv.negatives_to_structured(g["pos"], neg, K)
    assert np.array_equal(orv.structured_to_negatives(g["pos"], ent, side, K), neg)
