"""The PyTorch C++ extension over the C-ABI (multike_b200/torch_ops.py, csrc/torch_ext): registration and argument checks
without a GPU; on the GPU the ops reproduce the golden relation step and the ctypes path's evaluator."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ops():
    from multike_b200 import torch_ops
    if not os.path.exists(torch_ops.EXT_LIB):
        torch_ops.build()
    return torch_ops.load()


def test_ops_are_registered_and_check_their_arguments(ops):
    for name in ("rel_step", "rows_apply_adagrad", "sim_rank"):
        assert hasattr(ops, name)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.rows_apply_adagrad(torch.zeros(4, 8), torch.zeros(4, 8), None, torch.zeros(4, 8), 0.1, 8, True)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.sim_rank(torch.zeros(4, 8), None, torch.zeros(5, 8), None, None, 8, True)


@pytest.mark.gpu
def test_ops_reproduce_the_golden_relation_step(ops):
    """MultiKE_model.py:123-131 + losses.py:4-12 + Adagrad through torch.ops.multike_b200: loss and post-Adagrad rows of
    tests/golden/relation_step_d75.npz (fp64 autograd oracle), tolerances of tests/test_gpu_parity.py"""
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "relation_step_d75.npz")))
    K, lr, dim, stride = int(g["K"]), float(g["lr"]), 75, 80
    dev = "cuda"

    def table(init):
        var = torch.zeros(init.shape[0], stride, device=dev)
        var[:, :dim] = torch.as_tensor(init.astype(np.float32), device=dev)
        return var, torch.zeros_like(var), torch.full_like(var, 0.1)

    ent, ent_g, ent_a = table(g["ent0"])
    rel, rel_g, rel_a = table(g["rel0"])
    touched = torch.zeros(ent.shape[0], dtype=torch.uint8, device=dev)
    pos = torch.as_tensor(g["pos"].astype(np.int32), device=dev).contiguous()
    neg_ent = torch.as_tensor(g["neg_ent"].astype(np.int32), device=dev).contiguous()
    neg_side = torch.as_tensor(g["neg_side"].astype(np.uint32).view(np.int32), device=dev).contiguous()
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    ops.rel_step(ent, ent_g, touched, rel, rel_g, pos, neg_ent, neg_side, K, None, 1.0, loss, dim, True, True, 0)
    ops.rows_apply_adagrad(ent, ent_g, touched, ent_a, lr, dim, True)
    ops.rows_apply_adagrad(rel, rel_g, None, rel_a, lr, dim, True)
    torch.cuda.synchronize()
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-5)
    keep = np.ones(ent.shape[0], bool)
    keep[5] = False                      # (the clamp row of the fixture, compared relatively in test_gpu_parity.py)
    assert np.abs(ent[:, :dim].cpu().numpy()[keep] - g["ent1"][keep]).max() < 2e-6
    assert np.abs(rel[:, :dim].cpu().numpy() - g["rel1"]).max() < 2e-6
    assert float(ent_g.abs().max()) == 0.0 and int(touched.max()) == 0


@pytest.mark.gpu
def test_sim_rank_op_equals_the_ctypes_path(ops):
    from multike_b200 import similarity as S
    gen = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(700, 80, device="cuda", generator=gen)
    b = torch.randn(1500, 80, device="cuda", generator=gen)
    gold = torch.randint(0, 1500, (700,), device="cuda", generator=gen, dtype=torch.int32)
    rank, top1 = ops.sim_rank(a, None, b, None, gold, 75, True)
    want_rank, want_top1 = S.sim_rank(a, b, gold=gold, normalize=True, dim=75)
    assert torch.equal(rank, want_rank) and torch.equal(top1, want_top1)
