"""Helpers shared by the -m gpu parity tests: build device tables from oracle/golden arrays and
run the C-ABI entry points (through multike_b200.tables, the ctypes wrappers)."""
import numpy as np
import torch

from multike_b200 import tables as T

DEV = "cuda"


def make_tables(ent0, rel0, ent_norm=True, rel_norm=True, flags=(True, True), rel_replicas=None):
    """flags: per table, True = touched-byte map, False = mke_table_t.touched NULL (phase 2 sweeps
    every row), None = the product default (by table size)."""
    ent0 = np.asarray(ent0, dtype=np.float32)
    rel0 = np.asarray(rel0, dtype=np.float32)
    ent = T.EmbeddingTable(ent0.shape[0], ent0.shape[1], ent_norm, DEV, init=ent0, name="ent", flags=flags[0],
                           grad_replicas=1)
    rel = T.EmbeddingTable(rel0.shape[0], rel0.shape[1], rel_norm, DEV, init=rel0, name="rel", flags=flags[1],
                           grad_replicas=rel_replicas)
    return ent, rel


def grad_np(table):
    torch.cuda.synchronize()
    return table.grad_sum()[:, : table.dim].double().cpu().numpy()


def pad_is_zero(table):
    torch.cuda.synchronize()
    if table.stride == table.dim:
        return True
    ok = float(table.var[:, table.dim:].abs().max()) == 0.0
    if table.grad is not None:
        ok = ok and float(table.grad[..., table.dim:].abs().max()) == 0.0
    return ok


def loss_value(acc):
    torch.cuda.synchronize()
    return float(acc.cpu()[0])


def torch_dense_step(var_e, var_r, pos, neg, lr, acc_e, acc_r, pos_w=None, pos_scale=1.0):
    """Dense TF-graph semantics with torch autograd on the GPU, in the dtype of var_e (the tests
    pass float64) -- checker for full-size cases where the CPU oracle would take minutes:
    MultiKE_model.py:123-131 + losses.py:4-12 + dense Adagrad.  Same code as
    oracle/relation_view.py::relation_view_step, device-agnostic.  Returns loss; updates var/acc
    in place."""
    ve = var_e.clone().requires_grad_(True)
    vr = var_r.clone().requires_grad_(True)

    def l2n(x):
        return x * torch.rsqrt(torch.clamp((x * x).sum(1, keepdim=True), min=1e-12))

    E, R = l2n(ve), l2n(vr)
    ph, pr, pt = pos[:, 0].long(), pos[:, 1].long(), pos[:, 2].long()
    d = E[ph] + R[pr] - E[pt]
    lp = torch.log(1 + torch.exp((d * d).sum(1)))
    if pos_w is not None:
        lp = lp * pos_w
    loss = pos_scale * lp.sum()
    if neg is not None and neg.numel():
        nh, nr, nt = neg[:, 0].long(), neg[:, 1].long(), neg[:, 2].long()
        d = E[nh] + R[nr] - E[nt]
        loss = loss + torch.log(1 + torch.exp(-(d * d).sum(1))).sum()
    ge, gr = torch.autograd.grad(loss, [ve, vr])
    with torch.no_grad():
        acc_e += ge * ge
        var_e -= lr * ge * torch.rsqrt(acc_e)
        acc_r += gr * gr
        var_r -= lr * gr * torch.rsqrt(acc_r)
    return float(loss.detach()), ge, gr
