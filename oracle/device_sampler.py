"""Bit-exact CPU restatement of the ON-DEVICE negative sampler (multike_b200/csrc/mke_common.cuh:
mix64 / stream_key / draw64 / draw_index / sample_negs_sequential).

The device sampler keeps the semantics of generate_neg_triples_fast (base/batch.py:86-116) --
one head/tail coin per round, draws without replacement inside a round, known triples filtered,
last round unfiltered, exactly K per positive, positive-major output -- but replaces CPython's
Mersenne Twister by a counter-based generator so that CPU and GPU agree bit for bit.  Order
inside a positive is draw order (the reference's is Python-set order, i.e. unspecified).
"""
import numpy as np

MASK = (1 << 64) - 1
GAMMA = 0x9E3779B97F4A7C15
SIDE_DRAW = 0xFFFF
MAX_TRY = 10


def mix64(x):
    x &= MASK
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & MASK
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & MASK
    x ^= x >> 31
    return x


def stream_key(seed, step):
    return mix64((seed + GAMMA * (step + 1)) & MASK)


def draw64(skey, i, tr, c):
    coord = (i << 20) | (tr << 16) | c
    return mix64((skey + (coord + 1) * GAMMA) & MASK)


def draw_index(r, n):
    return ((r >> 32) * n) >> 32


class KG:
    """Candidate pool + filter set of one KG (mirror of mke_kg_sampler_t)."""

    def __init__(self, entity_base=0, n_entities=0, entity_list=None, triples=None, neighbours=None):
        self.entity_list = None if entity_list is None else [int(e) for e in entity_list]
        self.entity_base = int(entity_base)
        self.n_entities = len(self.entity_list) if self.entity_list is not None else int(n_entities)
        self.set = None if triples is None else {tuple(int(x) for x in t) for t in np.asarray(triples).reshape(-1, 3)}
        self.neighbours = None if neighbours is None else np.asarray(neighbours)

    def pool(self, anchor):
        if self.neighbours is not None and self.neighbours[anchor, 0] >= 0:
            row = self.neighbours[anchor]
            return (lambda k: int(row[k])), len(row)
        if self.entity_list is not None:
            return (lambda k: self.entity_list[k]), self.n_entities
        return (lambda k: self.entity_base + k), self.n_entities

    def contains(self, h, r, t):
        return self.set is not None and (h, r, t) in self.set


def sample_one(kg, h, r, t, K, skey, i):
    """Negatives of positive number i of the batch: list of K (h', r, t') tuples."""
    picks, sides = [], []
    remaining = K
    for tr in range(MAX_TRY):
        head_side = (draw64(skey, i, tr, SIDE_DRAW) >> 63) != 0
        at, n = kg.pool(h if head_side else t)
        cur, c = [], 0
        while len(cur) < remaining:
            e = at(draw_index(draw64(skey, i, tr, c), n))
            c += 1
            if e in cur and c < SIDE_DRAW:
                continue
            cur.append(e)
        if tr != MAX_TRY - 1:
            cur = [e for e in cur if not (kg.contains(e, r, t) if head_side else kg.contains(h, r, e))]
        picks += cur
        sides += [head_side] * len(cur)
        if len(picks) >= K:
            break
        remaining = K - len(picks)
    return [((e, r, t) if hs else (h, r, e)) for e, hs in zip(picks, sides)]


def sample_batch(pos1, kg1, pos2, kg2, K, seed, step):
    """mke_sample_uniform / the sampler fused into mke_rel_step_sampled: int32 [(n1+n2)*K, 3]."""
    skey = stream_key(seed, step)
    pos1 = np.asarray(pos1, dtype=np.int64).reshape(-1, 3)
    pos2 = np.asarray(pos2, dtype=np.int64).reshape(-1, 3)
    out = []
    for i, (h, r, t) in enumerate(pos1):
        out += sample_one(kg1, int(h), int(r), int(t), K, skey, i)
    for k, (h, r, t) in enumerate(pos2):
        out += sample_one(kg2, int(h), int(r), int(t), K, skey, len(pos1) + k)
    return np.asarray(out, dtype=np.int32).reshape(-1, 3)


ATTR_MAX_TRY = 64


def sample_attribute_heads(pos1, kg1, pos2, kg2, K, seed, step, index_base=0):
    """mke_sample_attribute_heads: restatement of attr_batch.py:13-25 with the counter-based RNG
    (draw coordinate = j + 32 * try); returns int32 [(n1+n2), K] corrupted heads."""
    skey = stream_key(seed, step)
    pos1 = np.asarray(pos1, dtype=np.int64).reshape(-1, 3)
    pos2 = np.asarray(pos2, dtype=np.int64).reshape(-1, 3)
    out = []
    for i, (h, a, v) in enumerate(np.concatenate([pos1, pos2])):
        kg = kg1 if i < len(pos1) else kg2
        at, n = kg.pool(int(h))
        row = []
        for j in range(K):
            e = int(h)
            for tr in range(ATTR_MAX_TRY):
                e = at(draw_index(draw64(skey, index_base + i, 0, j + 32 * tr), n))
                if tr == ATTR_MAX_TRY - 1 or not kg.contains(e, int(a), int(v)):
                    break
            row.append(e)
        out.append(row)
    return np.asarray(out, dtype=np.int32).reshape(-1, K)


# ---- vectorised front end (numpy) for full-size CPU runs; same results as sample_batch -------------
def _mix64_np(x):
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


def _draw64_np(skey, i, tr, c):
    coord = (i.astype(np.uint64) << np.uint64(20)) | (np.uint64(tr) << np.uint64(16)) | c.astype(np.uint64)
    with np.errstate(over="ignore"):
        return _mix64_np(np.uint64(skey) + (coord + np.uint64(1)) * np.uint64(GAMMA))


def sample_batch_fast(pos1, kg1, pos2, kg2, K, seed, step, index_base=0):
    """sample_batch for uniform pools (no neighbour lists), vectorised: round 0 is evaluated for all
    positives at once; positives whose round-0 candidates repeat an entity or hit a known triple
    are replayed through sample_one (the sequential definition)."""
    skey = stream_key(seed, step)
    pos1 = np.asarray(pos1, dtype=np.int64).reshape(-1, 3)
    pos2 = np.asarray(pos2, dtype=np.int64).reshape(-1, 3)
    out = np.empty(((len(pos1) + len(pos2)) * K, 3), dtype=np.int32)
    off = 0
    for pos, kg in ((pos1, kg1), (pos2, kg2)):
        n = len(pos)
        if n == 0:
            continue
        assert kg.neighbours is None and kg.entity_list is None
        i = np.arange(off, off + n, dtype=np.uint64) + np.uint64(index_base)
        with np.errstate(over="ignore"):
            head_side = (_draw64_np(skey, i, 0, np.full(n, SIDE_DRAW, dtype=np.uint64)) >> np.uint64(63)) != 0
            r = _draw64_np(skey, np.repeat(i, K), 0, np.tile(np.arange(K, dtype=np.uint64), n))
            cand = (((r >> np.uint64(32)) * np.uint64(kg.n_entities)) >> np.uint64(32)).astype(np.int64) + kg.entity_base
        cand = cand.reshape(n, K)
        srt = np.sort(cand, axis=1)
        bad = (srt[:, 1:] == srt[:, :-1]).any(axis=1)
        hs = np.repeat(head_side, K).reshape(n, K)
        nh = np.where(hs, cand, pos[:, [0]])
        nt = np.where(hs, pos[:, [2]], cand)
        nr = np.broadcast_to(pos[:, [1]], (n, K))
        if kg.set is not None:
            if not hasattr(kg, "_keys"):
                arr = np.array(sorted(kg.set), dtype=np.int64)
                kg._keys = np.sort((arr[:, 0] << 40) | (arr[:, 1] << 24) | arr[:, 2])
            keys = (nh << 40) | (nr << 24) | nt
            idx = np.searchsorted(kg._keys, keys.ravel())
            idx[idx >= len(kg._keys)] = len(kg._keys) - 1
            bad |= (kg._keys[idx] == keys.ravel()).reshape(n, K).any(axis=1)
        block = np.stack([nh, nr, nt], axis=2).astype(np.int32)
        for k in np.nonzero(bad)[0]:
            h, rr, t = (int(x) for x in pos[k])
            block[k] = np.asarray(sample_one(kg, h, rr, t, K, skey, int(index_base + off + k)), dtype=np.int32)
        out[off * K:(off + n) * K] = block.reshape(n * K, 3)
        off += n
    return out


# ---- mke_sample_distinct (csrc/mke_sampler.cu): random.sample(range(n), count) as the first images of a keyed Feistel
# permutation with cycle walking -- restated bit-exactly.  (The reference draws these batches with random.sample,
# MultiKE_model.py:355-358, :377, :399, :443, :462; its CPython stream is not reproduced, the semantics -- distinct,
# uniform indices -- are what the tests check.)
def _feistel_permute(x, half_bits, key):
    mask = (1 << half_bits) - 1
    left, right = x >> half_bits, x & mask
    for rnd in range(6):
        f = (mix64((key + (rnd + 1) * GAMMA + right) & MASK) >> 20) & mask
        left, right = right, left ^ f
    return (left << half_bits) | right


def sample_distinct(n, count, seed, draw):
    """[count] distinct indices of [0, n): what mke_sample_distinct returns for (seed, draw)"""
    half_bits = 1
    while (1 << (2 * half_bits)) < n:
        half_bits += 1
    key = stream_key((seed ^ 0x5DEECE66D) & MASK, draw)
    out = []
    for i in range(count):
        x = i
        while True:
            x = _feistel_permute(x, half_bits, key)
            if x < n:
                break
        out.append(x)
    return out
