"""Synthetic DWY100K-shaped relation triples (no dataset or network on the GPU box).

Two KGs with disjoint id ranges (kg1 entities [0, n/2), kg2 [n/2, n); relations likewise), as the
reference loader produces for DBP-WD-100K (200 000 entities, 550 relations,
463 294 + 448 774 local relation triples).  Endpoints follow a Zipf-Mandelbrot law
p(rank) ~ (rank + 10)^-0.65 and relations p(rank) ~ (rank + 2)^-1.25, fitted to the measured
degree distribution of the real dataset; every entity appears at least once as a head.
"""
import numpy as np

DWY100K = dict(n_ent=200_000, n_rel=550, n_rel1=330, n_triples1=463_294, n_triples2=448_774)
SYNTH_1M = dict(n_ent=1_000_000, n_rel=1_000, n_rel1=500, n_triples1=5_000_000, n_triples2=5_000_000)


def _zipf_mandelbrot(rng, n, size, shift, expo):
    p = (np.arange(n, dtype=np.float64) + shift) ** (-expo)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    return np.searchsorted(cdf, rng.random(size), side="right").astype(np.int64).clip(0, n - 1)


def _one_kg(rng, ent_lo, n_ent, rel_lo, n_rel, n_triples):
    perm_h = rng.permutation(n_ent)
    perm_t = rng.permutation(n_ent)
    perm_r = rng.permutation(n_rel)
    out = np.empty((0, 3), dtype=np.int64)
    # every entity once as a head, the rest from the power law; de-duplicate and top up
    heads = np.concatenate([np.arange(n_ent), perm_h[_zipf_mandelbrot(rng, n_ent, max(n_triples - n_ent, 0), 10, 0.65)]])
    heads = heads[:n_triples]
    need = n_triples
    while True:
        k = heads.shape[0] if out.shape[0] == 0 else int(need * 1.1) + 16
        h = heads if out.shape[0] == 0 else perm_h[_zipf_mandelbrot(rng, n_ent, k, 10, 0.65)]
        t = perm_t[_zipf_mandelbrot(rng, n_ent, h.shape[0], 10, 0.65)]
        r = perm_r[_zipf_mandelbrot(rng, n_rel, h.shape[0], 2, 1.25)]
        cand = np.stack([h + ent_lo, r + rel_lo, t + ent_lo], 1)
        out = np.concatenate([out, cand], 0)
        key = (out[:, 0] << 40) | (out[:, 1] << 24) | out[:, 2]
        _, first = np.unique(key, return_index=True)
        out = out[np.sort(first)]
        if out.shape[0] >= n_triples:
            out = out[:n_triples]
            break
        need = n_triples - out.shape[0]
    return out[rng.permutation(n_triples)].astype(np.int32)


def make_kgs(shape=None, seed=1234, **over):
    """Returns dict(triples1, triples2 [n,3] int32, n_ent, n_rel, ent_split, rel_split)."""
    cfg = dict(DWY100K if shape is None else shape)
    cfg.update(over)
    rng = np.random.default_rng(seed)
    half = cfg["n_ent"] // 2
    t1 = _one_kg(rng, 0, half, 0, cfg["n_rel1"], cfg["n_triples1"])
    t2 = _one_kg(rng, half, cfg["n_ent"] - half, cfg["n_rel1"], cfg["n_rel"] - cfg["n_rel1"], cfg["n_triples2"])
    return dict(triples1=t1, triples2=t2, n_ent=cfg["n_ent"], n_rel=cfg["n_rel"], ent_split=half,
                rel_split=cfg["n_rel1"])


def _mix64(x):
    """splitmix64 finaliser on uint64 arrays (the counter-based generator of csrc/mke_common.cuh)"""
    x = x.copy()
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


def literal_vectors(literal_ids, dim, seed=20190754):
    """Stand-in for the literal / name embeddings when the authors' word-vector file is not available
    (SURVEY.md section 8c "name view caveat"): literal i -> a unit vector of `dim` Gaussians that is a
    pure function of (seed, i), so identical literals get identical vectors on every machine and the
    tables never have to be stored.  float32 [len(literal_ids), dim]."""
    ids = np.asarray(literal_ids, dtype=np.uint64).reshape(-1, 1)
    cols = np.arange(dim, dtype=np.uint64).reshape(1, -1)
    with np.errstate(over="ignore"):
        key = (ids << np.uint64(20)) | cols
        a = _mix64(key * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed))
        b = _mix64(a + np.uint64(0x9E3779B97F4A7C15))
    u1 = ((a >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / 9007199254740992.0)   # (0, 1]
    u2 = (b >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    z /= np.linalg.norm(z, axis=1, keepdims=True)
    return z.astype(np.float32)
