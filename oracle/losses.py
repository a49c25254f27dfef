"""Restatement of code/losses.py (all eight functions) with torch tensors on CPU.

Every function follows the reference line by line (losses.py:<line> in each docstring); the
naive ``log(1 + exp(x))`` is kept on purpose (no softplus stabilisation in the reference).
"""
import torch

from .tf_semantics import l2_normalize


def _score(hs, ms, ts):
    d = hs + ms - ts
    return -(d * d).sum(dim=1)


def relation_logistic_loss(phs, prs, pts, nhs, nrs, nts):
    """losses.py:4-12"""
    pos_score = _score(phs, prs, pts)
    neg_score = _score(nhs, nrs, nts)
    pos_loss = torch.log(1 + torch.exp(-pos_score)).sum()
    neg_loss = torch.log(1 + torch.exp(neg_score)).sum()
    return pos_loss + neg_loss


def attribute_logistic_loss(phs, pas, pvs, pws, nhs, nas, nvs, nws):
    """losses.py:15-27"""
    pos_score = torch.log(1 + torch.exp(-_score(phs, pas, pvs))) * pws
    neg_score = torch.log(1 + torch.exp(_score(nhs, nas, nvs))) * nws
    return pos_score.sum() + neg_score.sum()


def relation_logistic_loss_wo_negs(phs, prs, pts):
    """losses.py:30-34"""
    return torch.log(1 + torch.exp(-_score(phs, prs, pts))).sum()


def attribute_logistic_loss_wo_negs(phs, pas, pvs):
    """losses.py:37-41"""
    return torch.log(1 + torch.exp(-_score(phs, pas, pvs))).sum()


def logistic_loss_wo_negs(phs, pas, pvs, pws):
    """losses.py:44-50"""
    return (torch.log(1 + torch.exp(-_score(phs, pas, pvs))) * pws).sum()


def orthogonal_loss(mapping, eye):
    """losses.py:61-63"""
    return ((mapping @ mapping.t() - eye) ** 2).sum()


def space_mapping_loss(view_embeds, shared_embeds, mapping, eye, orthogonal_weight, norm_w=0.0001):
    """losses.py:53-58 (note the axis-less l2_normalize: global Frobenius norm of the batch)"""
    mapped = l2_normalize(view_embeds @ mapping)
    map_loss = ((shared_embeds - mapped) ** 2).sum()
    norm_loss = (mapping ** 2).sum()
    return map_loss + orthogonal_weight * orthogonal_loss(mapping, eye) + norm_w * norm_loss


def alignment_loss(ents1, ents2):
    """losses.py:66-69"""
    d = ents1 - ents2
    return (d * d).sum()
