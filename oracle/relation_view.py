"""Dense-semantics oracle of the relation-view training step (and its positives-only variants).

Follows MultiKE_model.py:114-132 (graph), :158-170 / :187-201 (ckge / ckgp variants),
base/initializers.py:22-26 (the table the graph reads is l2_normalize(var, 1)), losses.py and
MultiKE_model.py:15-31 (fresh Adagrad per graph).  Gradients come from torch autograd through the
normalisation of the WHOLE table and Adagrad is applied to EVERY row -- exactly what the TF graph
does -- so this file is independent of the hand-derived sparse form used by the CUDA kernels.
"""
import numpy as np
import torch

from . import losses
from .tf_semantics import ADAGRAD_INIT, adagrad_dense_, l2_normalize


def _t(x, dtype):
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def _idx(x):
    return torch.as_tensor(np.asarray(x), dtype=torch.long)


class DenseTable:
    """A variable, its normalised-view flag and one Adagrad accumulator per optimizer slot."""

    def __init__(self, var, normalised=True, dtype=torch.float32):
        self.var = _t(var, dtype).clone()
        self.normalised = normalised
        self.slots = {}

    def view(self, v=None):
        v = self.var if v is None else v
        return l2_normalize(v, 1) if self.normalised else v

    def acc(self, slot):
        if slot not in self.slots:
            self.slots[slot] = torch.full_like(self.var, ADAGRAD_INIT)
        return self.slots[slot]


def relation_view_step(ent, rel, ph, pr, pt, nh, nr, nt, lr, slot="relation", pos_w=None, pos_scale=1.0,
                       apply=True):
    """One session.run([relation_loss, relation_optimizer]) (MultiKE_model.py:304-310).

    With nh/nr/nt empty and pos_scale=2 it is the ckge step (:168); with pos_w also the ckgp step
    (:198).  Returns (loss, dense grad wrt ent.var, dense grad wrt rel.var).
    """
    dt = ent.var.dtype
    ph, pr, pt, nh, nr, nt = map(_idx, (ph, pr, pt, nh, nr, nt))
    ve = ent.var.clone().requires_grad_(True)
    vr = rel.var.clone().requires_grad_(True)
    E, R = ent.view(ve), rel.view(vr)
    if pos_w is None:
        loss = pos_scale * losses.relation_logistic_loss_wo_negs(E[ph], R[pr], E[pt])
    else:
        loss = pos_scale * losses.logistic_loss_wo_negs(E[ph], R[pr], E[pt], _t(pos_w, dt))
    if nh.numel():
        neg_score = -((E[nh] + R[nr] - E[nt]) ** 2).sum(dim=1)
        loss = loss + torch.log(1 + torch.exp(neg_score)).sum()
    ge, gr = torch.autograd.grad(loss, [ve, vr])
    if apply:
        with torch.no_grad():
            adagrad_dense_(ent.var, ent.acc(slot), ge, lr)
            adagrad_dense_(rel.var, rel.acc(slot), gr, lr)
    return float(loss.detach()), ge, gr


def view_gradients(ent, rel, ph, pr, pt, nh, nr, nt, pos_w=None, pos_scale=1.0):
    """d loss / d E and d loss / d R for the (normalised) views -- what phase 1 accumulates."""
    dt = ent.var.dtype
    ph, pr, pt, nh, nr, nt = map(_idx, (ph, pr, pt, nh, nr, nt))
    E = ent.view().detach().clone().requires_grad_(True)
    R = rel.view().detach().clone().requires_grad_(True)
    w = torch.ones(ph.numel(), dtype=dt) if pos_w is None else _t(pos_w, dt)
    loss = pos_scale * losses.logistic_loss_wo_negs(E[ph], R[pr], E[pt], w)
    if nh.numel():
        neg_score = -((E[nh] + R[nr] - E[nt]) ** 2).sum(dim=1)
        loss = loss + torch.log(1 + torch.exp(neg_score)).sum()
    gE, gR = torch.autograd.grad(loss, [E, R])
    return float(loss.detach()), gE, gR


def triple_scores(head, mid, tail, ih, im, it):
    """pos_score / neg_score tensors of losses.py:7-8 for arbitrary tables."""
    d = head.view()[_idx(ih)] + mid.view()[_idx(im)] - tail.view()[_idx(it)]
    return -(d * d).sum(dim=1)


def negatives_to_structured(pos, neg, K):
    """(h,r,t) negatives, positive-major -> (corrupted entity [n,K], head-side bit mask [n])."""
    pos = np.asarray(pos, dtype=np.int64).reshape(-1, 3)
    neg = np.asarray(neg, dtype=np.int64).reshape(-1, K, 3)
    n = pos.shape[0]
    ent = np.zeros((n, K), dtype=np.int32)
    side = np.zeros(n, dtype=np.uint32)
    for i in range(n):
        for j in range(K):
            nhj, _, ntj = neg[i, j]
            if ntj == pos[i, 2] and nhj != pos[i, 0]:
                ent[i, j] = nhj
                side[i] |= np.uint32(1 << j)
            elif nhj == pos[i, 0]:
                ent[i, j] = ntj
            else:
                raise ValueError("negative %d of positive %d corrupts both sides" % (j, i))
    return ent, side


def structured_to_negatives(pos, neg_ent, neg_side, K):
    pos = np.asarray(pos, dtype=np.int64).reshape(-1, 3)
    n = pos.shape[0]
    out = np.zeros((n, K, 3), dtype=np.int64)
    for j in range(K):
        hs = ((np.asarray(neg_side, dtype=np.uint64) >> np.uint64(j)) & np.uint64(1)).astype(bool)
        e = np.asarray(neg_ent).reshape(n, K)[:, j]
        out[:, j, 0] = np.where(hs, e, pos[:, 0])
        out[:, j, 1] = pos[:, 1]
        out[:, j, 2] = np.where(hs, pos[:, 2], e)
    return out.reshape(n * K, 3)
