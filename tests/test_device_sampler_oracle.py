"""CPU checks of oracle/device_sampler.py (the bit-exact restatement of the on-device sampler):
it must keep the semantics of the reference sampler (base/batch.py:86-116)."""
import random

import numpy as np

from oracle import device_sampler as ds
from oracle import ref_batch


def _kg(golden):
    g = golden("ref_batch_relation.npz")
    t1 = [tuple(int(x) for x in r) for r in g["triples1"]]
    sup1 = [tuple(int(x) for x in r) for r in g["sup1"]]
    return t1, sup1, int(g["n_ent"])


def test_mix64_known_values():
    # splitmix64 finaliser; values computed independently with Python integers
    assert ds.mix64(0) == 0
    assert ds.mix64(1) == 0x5692161D100B05E5
    assert ds.stream_key(0, 0) == ds.mix64(ds.GAMMA)
    assert ds.draw_index(0xFFFFFFFF00000000, 100000) == 99999 and ds.draw_index(0, 7) == 0


def test_semantics_exactly_k_filtered_single_side(golden):
    t1, sup1, n_ent = _kg(golden)
    kg = ds.KG(entity_base=0, n_entities=n_ent, triples=t1 + sup1)
    K = 10
    pos = np.array(t1[:200])
    neg = ds.sample_batch(pos, kg, np.zeros((0, 3)), kg, K, seed=5, step=2).reshape(-1, K, 3)
    known = set(t1) | set(sup1)
    n_filtered_violations = 0
    for (h, r, t), negs in zip(pos, neg):
        assert len(negs) == K
        for nh, nr, nt in negs:
            assert nr == r and (nh == h or nt == t)              # one side corrupted, relation kept
            assert 0 <= nh < n_ent and 0 <= nt < n_ent           # same-KG candidates
            n_filtered_violations += (int(nh), int(nr), int(nt)) in known
    # only a 10th-try (unfiltered) round may emit a known triple: vanishingly rare here
    assert n_filtered_violations == 0
    # deterministic in (seed, step, index) and sensitive to each of them
    again = ds.sample_batch(pos, kg, np.zeros((0, 3)), kg, K, seed=5, step=2).reshape(-1, K, 3)
    assert np.array_equal(neg, again)
    assert not np.array_equal(neg, ds.sample_batch(pos, kg, np.zeros((0, 3)), kg, K, 5, 3).reshape(-1, K, 3))
    assert not np.array_equal(neg, ds.sample_batch(pos, kg, np.zeros((0, 3)), kg, K, 6, 2).reshape(-1, K, 3))


def test_statistics_match_reference_sampler(golden):
    """Head/tail side frequency and the spread of corrupted entities agree with the reference
    sampler restatement (which is pinned to the reference itself)."""
    t1, sup1, n_ent = _kg(golden)
    known = set(t1) | set(sup1)
    ents = list(range(n_ent))
    K = 10
    pos = t1[:600]
    random.seed(3)
    np.random.seed(3)
    ref = np.array(ref_batch.neg_triples_fast(pos, known, ents, K)).reshape(-1, K, 3)
    kg = ds.KG(entity_base=0, n_entities=n_ent, triples=t1 + sup1)
    dev = ds.sample_batch(np.array(pos), kg, np.zeros((0, 3)), kg, K, seed=3, step=0).reshape(-1, K, 3)
    P = np.array(pos)

    def head_frac(neg):
        return float(np.mean(neg[:, :, 0] != P[:, None, 0]))

    assert abs(head_frac(ref) - head_frac(dev)) < 0.06 and abs(head_frac(dev) - 0.5) < 0.06

    def corrupted(neg):
        hs = neg[:, :, 0] != P[:, None, 0]
        return np.where(hs, neg[:, :, 0], neg[:, :, 2]).ravel()

    hr, hd = np.bincount(corrupted(ref), minlength=n_ent), np.bincount(corrupted(dev), minlength=n_ent)
    assert hr.sum() == hd.sum() == 600 * K
    # both are (filtered) uniform over 40 entities: compare with a chi-square-like bound
    expect = 600 * K / n_ent
    assert np.abs(hd - expect).max() < 6 * np.sqrt(expect) and np.abs(hr - expect).max() < 6 * np.sqrt(expect)
    # duplicates inside a positive can only come from a later round re-drawing an entity (this KG
    # is dense, so later rounds are common): both samplers must show the same rate
    dup_dev = np.mean([len({tuple(x) for x in row}) < K for row in dev])
    dup_ref = np.mean([len({tuple(x) for x in row}) < K for row in ref])
    assert abs(dup_dev - dup_ref) < 0.08 and dup_dev < 0.4


def test_neighbour_lists_restrict_candidates(golden):
    t1, sup1, n_ent = _kg(golden)
    nb = -np.ones((n_ent, 12), dtype=np.int32)
    rng = np.random.default_rng(0)
    for e in range(0, n_ent, 2):
        nb[e] = rng.choice(n_ent, 12, replace=False)
    kg = ds.KG(entity_base=0, n_entities=n_ent, triples=t1 + sup1, neighbours=nb)
    pos = np.array(t1[:150])
    neg = ds.sample_batch(pos, kg, np.zeros((0, 3)), kg, 5, seed=1, step=0).reshape(-1, 5, 3)
    for (h, r, t), negs in zip(pos, neg):
        for nh, _, nt in negs:
            if nh != h and h % 2 == 0:
                assert nh in nb[h]
            if nt != t and t % 2 == 0:
                assert nt in nb[t]


def test_attribute_head_sampler_semantics_match_reference(golden):
    """oracle restatement of the device attribute sampler keeps attr_batch.py:13-25: head-only,
    K independent draws from the KG's entities, never a known (h, a, v)."""
    g = golden("ref_batch_attribute.npz")
    a1 = g["a1"]
    trip = a1[:, :3].astype(np.int64)
    n_ent = 40
    kg = ds.KG(entity_base=0, n_entities=n_ent, triples=trip)
    K = 4
    neg = ds.sample_attribute_heads(trip[:120], kg, np.zeros((0, 3)), kg, K, seed=2, step=1)
    known = {tuple(int(x) for x in t) for t in trip}
    assert neg.shape == (120, K) and neg.min() >= 0 and neg.max() < n_ent
    for (h, a, v), row in zip(trip[:120], neg):
        for e in row:
            assert (int(e), int(a), int(v)) not in known
    # reference restatement under its own RNG: same support and a flat histogram for both
    random.seed(4)
    ents = list(range(n_ent))
    pos = [(int(h), int(a), int(v), 1.0) for h, a, v in trip[:120]]
    ref = ref_batch.neg_attribute_triples(pos, {(h, a, v, 1.0) for h, a, v in known}, ents, K)
    ref_heads = np.array([t[0] for t in ref])
    assert len(ref) == 120 * K and all(t[1:3] == p[1:3] for t, p in zip(ref, [p for p in pos for _ in range(K)]))
    hr, hd = np.bincount(ref_heads, minlength=n_ent), np.bincount(neg.ravel(), minlength=n_ent)
    expect = 120 * K / n_ent
    assert np.abs(hr - expect).max() < 6 * np.sqrt(expect) and np.abs(hd - expect).max() < 6 * np.sqrt(expect)
    # index_base shifts the RNG coordinates: a rank holding positions [60, 120) draws the same
    part = ds.sample_attribute_heads(trip[60:120], kg, np.zeros((0, 3)), kg, K, seed=2, step=1, index_base=60)
    assert np.array_equal(part, neg[60:])


def test_sample_distinct_restatement_is_a_permutation_prefix():
    """oracle restatement of mke_sample_distinct: distinct indices, a permutation of range(n) at count == n, a
    function of (seed, draw); the GPU test compares the kernel with it bit for bit"""
    from oracle import device_sampler as ds
    for n in (1, 2, 7, 64, 1000, 4097):
        full = ds.sample_distinct(n, n, seed=5, draw=3)
        assert sorted(full) == list(range(n))
        assert ds.sample_distinct(n, min(n, 17), seed=5, draw=3) == full[:min(n, 17)]
    a, b = ds.sample_distinct(5000, 300, 5, 1), ds.sample_distinct(5000, 300, 5, 2)
    assert a != b and len(set(a)) == 300
    # uniform: mean position of many draws
    import numpy as np
    draws = np.concatenate([ds.sample_distinct(10007, 50, 9, k) for k in range(60)])
    assert abs(draws.mean() / 10007 - 0.5) < 0.03
