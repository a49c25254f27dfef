"""oracle/ref_batch.py must reproduce the REFERENCE's own batcher/sampler outputs bit for bit
(fixtures generated from /root/reference/code/base/batch.py and attr_batch.py, see
tests/golden/make_golden.py)."""
import random

import numpy as np

from oracle import ref_batch


def _tuples(a):
    return [tuple(int(x) for x in row) for row in a]


def test_relation_batches_match_reference(golden):
    g = golden("ref_batch_relation.npz")
    t1, t2 = _tuples(g["triples1"]), _tuples(g["triples2"])
    set1, set2 = set(t1) | set(_tuples(g["sup1"])), set(t2) | set(_tuples(g["sup2"]))
    n_ent = int(g["n_ent"])
    ents1, ents2 = list(range(n_ent)), list(range(n_ent, 2 * n_ent))
    nb1 = {int(k): [int(x) for x in v] for k, v in zip(g["nb1_keys"], g["nb1_vals"])}
    nb2 = {int(k): [int(x) for x in v] for k, v in zip(g["nb2_keys"], g["nb2_vals"])}
    B, K = int(g["B"]), int(g["K"])
    for case, (use_nb, seed, step) in enumerate(g["cases"]):
        random.seed(int(seed))
        np.random.seed(int(seed))
        pos, neg = ref_batch.relation_triple_batch(t1, t2, set1, set2, ents1, ents2, B, int(step),
                                                   nb1 if use_nb else None, nb2 if use_nb else None, K)
        assert np.array_equal(np.array(pos).reshape(-1, 3), g["case%d_pos" % case]), case
        assert np.array_equal(np.array(neg).reshape(-1, 3), g["case%d_neg" % case]), case
        assert len(neg) == K * len(pos)


def test_relation_tail_batches_are_short_or_empty(golden):
    """Epoch tail: slices are clipped (base/batch.py:48-50), kg2 may already be exhausted."""
    g = golden("ref_batch_relation.npz")
    cases = {int(step): c for c, (_, _, step) in enumerate(g["cases"])}
    full = len(g["case%d_pos" % cases[0]])
    assert full == int(g["B"])
    assert 0 < len(g["case%d_pos" % cases[11]]) < full      # both halves clipped / one empty
    assert len(g["case%d_pos" % cases[12]]) < full


def test_attribute_batches_match_reference(golden):
    g = golden("ref_batch_attribute.npz")
    a1 = [(int(h), int(a), int(v), float(w)) for h, a, v, w in g["a1"]]
    a2 = [(int(h), int(a), int(v), float(w)) for h, a, v, w in g["a2"]]
    ents1, ents2 = list(range(40)), list(range(40, 80))
    for case, (K, seed, step) in enumerate(g["cases"]):
        random.seed(int(seed))
        np.random.seed(int(seed))
        pos, neg = ref_batch.attribute_triple_batch(a1, a2, set(a1), set(a2), ents1, ents2, 64, int(step), None,
                                                    None, int(K))
        assert np.array_equal(np.array(pos, dtype=np.float64).reshape(-1, 4), g["case%d_pos" % case])
        assert np.array_equal(np.array(neg, dtype=np.float64).reshape(-1, 4), g["case%d_neg" % case])


def test_batch_split_matches_survey_numbers():
    # DBP-WD-100K: 463 294 + 448 774 local relation triples
    assert ref_batch.batch_sizes(463294, 448774, 5000) == (2539, 2461)
    assert ref_batch.batch_sizes(463294, 448774, 20000) == (10159, 9841)
