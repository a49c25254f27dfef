"""Mirror of code/base/alignment.py:8-79 (greedy_alignment) on the fused similarity/rank kernel.

Same signature and return value as the reference.  `nums_threads`, `metric` in
('inner', 'cosine'-with-normalize) and `accurate` are accepted; the result does not depend on
them (the kernel always ranks exactly; quick mode's argpartition gives the same Hits@k).
csls_k > 0 (unused by the MultiKE drivers: base/evaluation.py passes csls_k=0) is refused.
"""
import time

import numpy as np
import torch

from multike_b200 import similarity


def rank_metrics(rank, top_k):
    """calculate_rank's sums (base/alignment.py:141-163) from the 0-based gold ranks."""
    r = rank.to(torch.float64) + 1.0
    hits = [float((rank < k).sum().item()) for k in top_k]
    return float(r.mean().item()), float((1.0 / r).mean().item()), hits


def greedy_alignment(embed1, embed2, top_k, nums_threads, metric, normalize, csls_k, accurate):
    t = time.time()
    if csls_k and csls_k > 0:
        raise NotImplementedError("csls is not used by the MultiKE drivers (base/evaluation.py: csls_k=0)")
    if not (metric == 'inner' or (metric == 'cosine' and normalize)):
        raise NotImplementedError("metric %r: the MultiKE drivers rank by inner product" % (metric,))
    assert 1 in top_k
    rank, top1 = similarity.sim_rank(embed1, embed2, normalize=bool(normalize))
    num = int(rank.numel())
    mr, mrr, hits = rank_metrics(rank, top_k)
    alignment_rest = set(zip(range(num), top1.cpu().tolist()))
    assert len(alignment_rest) == num
    hits = np.array(hits) / num * 100
    for i in range(len(hits)):
        hits[i] = round(hits[i], 3)
    cost = time.time() - t
    if accurate:
        print("accurate results: hits@{} = {}%, mr = {:.3f}, mrr = {:.6f}, time = {:.3f} s ".
              format(top_k, hits, mr, mrr, cost))
    else:
        print("quick results: hits@{} = {}%, time = {:.3f} s ".format(top_k, hits, cost))
    hits1 = hits[0]
    return alignment_rest, hits1, mr, mrr
