"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's
evaluator and neighbour search, in numpy.

  sim / greedy_alignment / calculate_rank   base/similarity.py:9-52, base/alignment.py:8-79, :141-163
  find_neighbours                           base/batch.py:141-150

Pinned: tests/test_oracle_alignment.py checks these functions against outputs of the reference's
own modules imported in the build container (tests/golden/ref_sim.npz, generator
tests/golden/make_golden_sim.py).  One deliberate sharpening: the reference sorts with
np.argsort / np.argpartition, whose order among EQUAL sims is unspecified; the restatement (and the
kernels) use the stable order -- equal sims rank by ascending column.
"""
import numpy as np


def normalize_rows(x):
    """sklearn.preprocessing.normalize(x) (l2, axis=1): zero rows stay zero."""
    x = np.asarray(x, dtype=np.float32)
    n = np.sqrt((x.astype(np.float64) ** 2).sum(1)).astype(np.float32)
    n[n == 0] = 1.0
    return x / n[:, None]


def sim(embed1, embed2, normalize=False, dtype=np.float32):
    """base/similarity.py:31-35 (metric='inner'): np.matmul of the (normalised) rows."""
    if normalize:
        embed1, embed2 = normalize_rows(embed1), normalize_rows(embed2)
    return np.matmul(np.asarray(embed1, dtype=dtype), np.asarray(embed2, dtype=dtype).T)


def gold_ranks(sim_mat, gold=None):
    """0-based position of the gold column in the stable descending order of every row
    (rank_index of base/alignment.py:153) and the arg-max column (rank[0], :151)."""
    n = sim_mat.shape[0]
    gold = np.arange(n) if gold is None else np.asarray(gold)
    sg = sim_mat[np.arange(n), gold][:, None]
    cols = np.arange(sim_mat.shape[1])[None, :]
    before = (sim_mat > sg) | ((sim_mat == sg) & (cols < gold[:, None]))
    return before.sum(1).astype(np.int64), np.argmax(sim_mat, axis=1)  # argmax: first maximum


def calculate_rank(sim_mat, top_k, gold=None):
    """base/alignment.py:141-163 -> (mr, mrr, hits counts, alignment_rest)."""
    rank, top1 = gold_ranks(sim_mat, gold)
    n = sim_mat.shape[0]
    g = np.arange(n) if gold is None else np.asarray(gold)
    hits = [int((rank < k).sum()) for k in top_k]
    return float((rank + 1).sum() / n), float((1.0 / (rank + 1)).sum() / n), hits, set(zip(g.tolist(), top1.tolist()))


def greedy_alignment(embed1, embed2, top_k, normalize=True):
    """base/alignment.py:8-79 -> (alignment_rest, hits percent rounded to 3 places, mr, mrr)."""
    mr, mrr, hits, rest = calculate_rank(sim(embed1, embed2, normalize), top_k)
    hits = np.round(np.array(hits) / len(embed1) * 100, 3)
    return rest, hits, mr, mrr


def find_neighbours(entity_list, embed, k, dtype=np.float32):
    """base/batch.py:141-150 for all rows at once: row i -> the ids of the k most similar rows,
    as the stable top-k (ascending column among the chosen; ties at the k-th sim: smallest column)."""
    ids = np.asarray(entity_list)
    s = np.matmul(np.asarray(embed, dtype=dtype), np.asarray(embed, dtype=dtype).T)
    order = np.argsort(-s, axis=1, kind="stable")[:, :k]
    return ids[np.sort(order, axis=1)]
