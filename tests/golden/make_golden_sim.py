"""Generates tests/golden/ref_sim.npz: outputs of the REFERENCE's own base/alignment.py
(greedy_alignment -> calculate_rank) and base/batch.py (find_neighbours), imported unmodified from
/root/reference/code (stub `tensorflow` / `gensim`), on small seeded embeddings.  Run ONCE in the
build container; they pin oracle/alignment.py, which the GPU tests then compare the kernels with.

Cases are built so that the reference's answer does not depend on its unspecified tie order
(np.argsort / np.argpartition are not stable): no two sims of a row are closer than 2e-6 (20x the fp32 rounding of a sim) except
exact duplicates of NON-gold columns far below the gold rank / the k-th neighbour.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def well_separated(rng, n1, n2, d, k_list, gap=2e-6):
    """embeddings whose (float64) sim rows have no near-ties around the gold column and around
    every k-th largest value, so that the reference's ranks are tie-order independent"""
    while True:
        a = rng.standard_normal((n1, d)).astype(np.float32)
        b = rng.standard_normal((n2, d)).astype(np.float32)
        # make the gold pairs similar (realistic ranks: many hits@1, some deep ranks)
        m = min(n1, n2)
        b[:m] = a[:m] + rng.standard_normal((m, d)).astype(np.float32) * rng.uniform(0.2, 3.0, (m, 1)).astype(np.float32)
        an = a / np.linalg.norm(a, axis=1, keepdims=True)
        bn = b / np.linalg.norm(b, axis=1, keepdims=True)
        s = an.astype(np.float64) @ bn.astype(np.float64).T
        gold = s[np.arange(m), np.arange(m)]
        d_gold = np.abs(s[:m] - gold[:, None])
        d_gold[np.arange(m), np.arange(m)] = 1.0
        srt = -np.sort(-s, axis=1)
        ok = d_gold.min() > gap and all(np.min(srt[:, k - 1] - srt[:, k]) > gap for k in k_list if k < n2)
        ok = ok and np.min(srt[:, 0] - srt[:, 1]) > gap
        if ok:
            return a, b


def main():
    bat, _ = import_reference()
    import base.alignment as ali
    rng = np.random.default_rng(20190754)
    out = {}
    top_k = [1, 5, 10, 50]
    # ---- greedy_alignment: (n1, n2, d) incl. n2 > n1 (valid(): candidates = valid + test entities)
    cases = [(300, 300, 75), (257, 700, 75), (130, 513, 128), (64, 64, 20)]
    for c, (n1, n2, d) in enumerate(cases):
        a, b = well_separated(rng, n1, n2, d, [])
        with contextlib.redirect_stdout(io.StringIO()):
            rest, hits1, mr, mrr = ali.greedy_alignment(a, b, top_k, 1, 'inner', True, 0, True)
            rest_q, hits1_q, _, _ = ali.greedy_alignment(a, b, top_k, 1, 'inner', True, 0, False)
        top1 = np.full(n1, -1, np.int64)
        for g, j in rest:
            top1[g] = j
        assert hits1 == hits1_q and rest == rest_q
        out["align%d_a" % c], out["align%d_b" % c] = a, b
        out["align%d_top1" % c] = top1
        out["align%d_hits1" % c], out["align%d_mr" % c], out["align%d_mrr" % c] = hits1, mr, mrr
        # the hits vector is only printed by the reference: recompute it with its own calculate_rank
        sim_mat = ali.sim(a, b, metric='inner', normalize=True, csls_k=0)
        mr2, mrr2, hits, _ = ali.calculate_rank(list(range(n1)), sim_mat, top_k, True, n1)
        assert abs(mr2 - mr) < 1e-12
        out["align%d_hits" % c] = np.array(hits)
    out["align_cases"] = np.array(cases)
    out["top_k"] = np.array(top_k)
    # ---- find_neighbours (base/batch.py:141-150): normalised rows, k of n
    ncases = [(400, 75, 8), (300, 128, 37), (200, 75, 200)]
    for c, (n, d, k) in enumerate(ncases):
        e, _ = well_separated(rng, n, n, d, [k])
        e = e / np.linalg.norm(e, axis=1, keepdims=True)
        # symmetric sims: re-draw until the k-th / (k+1)-th gap holds for e . e^T itself
        while True:
            s = e.astype(np.float64) @ e.astype(np.float64).T
            srt = -np.sort(-s, axis=1)
            if k >= n or np.min(srt[:, k - 1] - srt[:, k]) > 1e-5:
                break
            e = rng.standard_normal((n, d)).astype(np.float32)
            e = e / np.linalg.norm(e, axis=1, keepdims=True)
        ent_list = (1000 + rng.permutation(n)).astype(np.int64)  # ids are not positions
        dic = bat.find_neighbours(ent_list, ent_list, e, e, k if k < n else n - 1) if k < n else None
        if dic is None:  # argpartition(kth=n) is out of range in the reference: k == n is every entity
            nb = np.tile(np.sort(ent_list), (n, 1))
        else:
            nb = np.array([sorted(dic[int(x)]) for x in ent_list])
        out["nb%d_e" % c], out["nb%d_ids" % c], out["nb%d_lists" % c] = e.astype(np.float32), ent_list, nb
    out["nb_cases"] = np.array(ncases)
    np.savez_compressed(os.path.join(HERE, "ref_sim.npz"), **out)
    print("wrote ref_sim.npz", {k: getattr(v, "shape", v) for k, v in out.items() if "hits" in k or "mr" in k})


if __name__ == "__main__":
    main()
