"""Device-resident embedding tables and the functional wrappers over the C-ABI.

Layout in HBM (DESIGN.md "Data layout"): every table is row-major fp32 ``[rows, stride]`` with
``stride = round_up(dim, 8)`` floats, i.e. rows start on 32-byte sector boundaries (dim=75 ->
320-byte rows = exactly ten sectors) and are 16-byte aligned for float4 / TMA bulk access.  Pad
columns are zero and stay zero.  Next to the variable live the gradient accumulator of the same
shape (zero between steps), one ``touched`` byte per row, and one Adagrad accumulator per
optimizer slot (the reference creates a fresh AdagradOptimizer per loss graph,
MultiKE_model.py:28-31).
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _cabi

ADAGRAD_INIT = 0.1  # tf.train.AdagradOptimizer initial_accumulator_value [TF semantics]


def padded_stride(dim):
    return (dim + 7) // 8 * 8


def xavier_truncated_normal(rows, dim, generator=None, device="cpu"):
    """tf.contrib.layers.xavier_initializer(uniform=False) (base/initializers.py:25):
    truncated normal, stddev sqrt(1.3 / ((fan_in + fan_out) / 2)), resampled beyond 2 sigma."""
    std = math.sqrt(1.3 / ((rows + dim) / 2.0))
    out = torch.empty(rows, dim, dtype=torch.float32, device=device)
    torch.nn.init.trunc_normal_(out, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=generator)
    return out


class EmbeddingTable:
    """One tf.get_variable (+ optional l2_normalize(var, 1) view) of MultiKE_model.py:86-107."""

    FLAGS_MIN_ROWS = 8192  # smaller tables are swept whole by phase 2 (mke_table_t.touched = NULL)
    # ... and get this many gradient copies (mke_table_t.grad_replicas)
    SMALL_TABLE_REPLICAS = int(os.environ.get("MKE_SMALL_REPLICAS", "7"))

    def __init__(self, rows, dim, normalised, device="cuda", init=None, trainable=True, name="", flags=None,
                 grad_replicas=None):
        self.rows, self.dim, self.normalised, self.name = int(rows), int(dim), bool(normalised), name
        self.stride = padded_stride(dim)
        self.device = torch.device(device)
        self.var = torch.zeros(self.rows, self.stride, dtype=torch.float32, device=self.device)
        if init is not None:
            src = torch.as_tensor(np.asarray(init, dtype=np.float32)) if not torch.is_tensor(init) else init
            assert tuple(src.shape) == (self.rows, self.dim), (src.shape, self.rows, self.dim)
            self.var[:, : self.dim] = src.to(self.device, torch.float32)
        self.trainable = bool(trainable)
        self.grad_replicas = 1
        if self.trainable:
            if grad_replicas is None:
                grad_replicas = self.SMALL_TABLE_REPLICAS if self.rows < self.FLAGS_MIN_ROWS else 1
            self.grad_replicas = max(1, int(grad_replicas))
            shape = (self.rows, self.stride) if self.grad_replicas == 1 else (self.grad_replicas, self.rows, self.stride)
            self.grad = torch.zeros(shape, dtype=torch.float32, device=self.device)
            if flags is None:
                flags = self.rows >= self.FLAGS_MIN_ROWS
            self.touched = torch.zeros(self.rows, dtype=torch.uint8, device=self.device) if flags else None
        else:
            self.grad = None
            self.touched = None
        self._slots = {}
        self._c = _cabi.MkeTable(
            var=self.var.data_ptr(),
            grad=_cabi.ptr(self.grad),
            touched=_cabi.ptr(self.touched),
            rows=self.rows, stride=self.stride, dim=self.dim, normalised=int(self.normalised),
            grad_replicas=self.grad_replicas)

    # -- C view ---------------------------------------------------------------------------
    @property
    def c(self):
        return ctypes.byref(self._c)

    def grad_sum(self):
        """[rows, stride] gradient accumulated so far (replicas summed)."""
        return self.grad if self.grad_replicas == 1 else self.grad.sum(0)

    # -- optimizer slots --------------------------------------------------------------------
    def adagrad_slot(self, slot):
        acc = self._slots.get(slot)
        if acc is None:
            acc = torch.full((self.rows, self.stride), ADAGRAD_INIT, dtype=torch.float32, device=self.device)
            self._slots[slot] = acc
        return acc

    def apply_adagrad(self, slot, lr):
        """Phase 2 for this table with the accumulators of optimizer `slot`."""
        apply_adagrad(self, self.adagrad_slot(slot), lr)

    # -- reads ------------------------------------------------------------------------------
    def export(self, idx=None):
        """Dense [n, dim] torch tensor of the view the model reads (normalised rows if flagged)."""
        lib = _cabi.load()
        if idx is None:
            n, idx_t = self.rows, None
        else:
            idx_t = torch.as_tensor(idx, dtype=torch.int32, device=self.device).contiguous()
            n = idx_t.numel()
        out = torch.empty(n, self.dim, dtype=torch.float32, device=self.device)
        _cabi.check(lib.mke_table_export(self.c, _cabi.ptr(idx_t), n, out.data_ptr(), _cabi.current_stream()))
        return out

    def eval(self, session=None, idx=None):
        """tensor.eval(session=...) of the reference drivers (MultiKE_Late.py:16-26)."""
        return self.export(idx).cpu().numpy()

    def raw(self):
        return self.var[:, : self.dim].detach().cpu().numpy()


class TripleSet:
    """all_triples_set of base/batch.py:86 as an open-addressing table in HBM."""

    def __init__(self, triples, device="cuda"):
        lib = _cabi.load()
        t = torch.as_tensor(np.ascontiguousarray(triples, dtype=np.int32)).reshape(-1, 3).to(device)
        n = t.shape[0]
        if n:
            assert int(t[:, [0, 2]].max()) < (1 << 24) - 1 and int(t[:, 1].max()) < (1 << 16), "id range"
        cap = 1 << max(4, int(math.ceil(math.log2(max(2 * n, 2)))))
        self.slots = torch.full((cap,), -1, dtype=torch.int64, device=device)  # all 0xFF..FF
        self._c = _cabi.MkeTripleSet(slots=self.slots.data_ptr(), capacity=cap)
        self.size = n
        _cabi.check(lib.mke_tripleset_build(ctypes.byref(self._c), t.data_ptr(), n, _cabi.current_stream()))
        torch.cuda.current_stream().synchronize()  # `t` may be freed after return

    def contains(self, triples):
        lib = _cabi.load()
        t = torch.as_tensor(np.ascontiguousarray(triples, dtype=np.int32)).reshape(-1, 3).to(self.slots.device)
        out = torch.empty(t.shape[0], dtype=torch.uint8, device=self.slots.device)
        _cabi.check(lib.mke_tripleset_contains(ctypes.byref(self._c), t.data_ptr(), t.shape[0], out.data_ptr(),
                                               _cabi.current_stream()))
        return out.cpu().numpy().astype(bool)


class KGSampler:
    """Candidate pool + filter set of one KG (arguments of generate_neg_triples_fast)."""

    def __init__(self, entity_base=0, n_entities=0, entity_list=None, triple_set=None, neighbours=None,
                 device="cuda"):
        self.entity_list = None
        if entity_list is not None:
            self.entity_list = torch.as_tensor(np.asarray(entity_list, dtype=np.int32)).to(device).contiguous()
            n_entities = self.entity_list.numel()
        self.triple_set = triple_set
        self.neighbours = None
        self.entity_base, self.n_entities = int(entity_base), int(n_entities)
        self._c = _cabi.MkeKgSampler()
        self._c.entity_list = _cabi.ptr(self.entity_list)
        self._c.entity_base = self.entity_base
        self._c.n_entities = self.n_entities
        if triple_set is not None:
            self._c.set = triple_set._c
        self.set_neighbours(neighbours, device)

    def set_neighbours(self, neighbours, device="cuda"):
        """neighbours: int32 [rows_of_entity_table, k] (row[0] == -1 => no list) or None."""
        if neighbours is None:
            self.neighbours = None
            self._c.neighbours = None
            self._c.n_neighbours = 0
        else:
            nb = neighbours if torch.is_tensor(neighbours) else torch.as_tensor(np.asarray(neighbours, np.int32))
            self.neighbours = nb.to(device=device, dtype=torch.int32).contiguous()
            self._c.neighbours = self.neighbours.data_ptr()
            self._c.n_neighbours = self.neighbours.shape[1]

    @property
    def c(self):
        return ctypes.byref(self._c)


# ---------------------------------------------------------------------------------------------
# functional wrappers (one per C entry point)
# ---------------------------------------------------------------------------------------------
def _i32(t, device):
    if t is None:
        return None
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.ascontiguousarray(t, dtype=np.int32))
    return t.to(device=device, dtype=torch.int32).contiguous()


def _f32(t, device):
    if t is None:
        return None
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.ascontiguousarray(t, dtype=np.float32))
    return t.to(device=device, dtype=torch.float32).contiguous()


def new_loss_accumulator(device="cuda"):
    return torch.zeros(1, dtype=torch.float64, device=device)


def triple_fwd_bwd(head, mid, tail, ih, im, it, loss_accum, w=None, negative=False, scale=1.0, score_out=None):
    """mke_triple_fwd_bwd: one losses.py term (+ its backward) over index vectors."""
    lib = _cabi.load()
    dev = head.device
    ih, im, it, w = _i32(ih, dev), _i32(im, dev), _i32(it, dev), _f32(w, dev)
    n = ih.numel()
    assert im.numel() == n and it.numel() == n and (w is None or w.numel() == n)
    _cabi.check(lib.mke_triple_fwd_bwd(head.c, mid.c, tail.c, _cabi.ptr(ih), _cabi.ptr(im), _cabi.ptr(it), n,
                                       _cabi.ptr(w), int(bool(negative)), float(scale), _cabi.ptr(loss_accum),
                                       _cabi.ptr(score_out), _cabi.current_stream()))
    return ih, im, it, w  # keep-alive handles for the caller


def rel_step_sampled(ent, rel, pos1, kg1, pos2, kg2, K, seed, step, loss_accum, w=None, pos_scale=1.0,
                     neg_out=None, variant=0):
    """mke_rel_step_sampled: fused phase 1 with on-device negative sampling."""
    lib = _cabi.load()
    dev = ent.device
    pos1, pos2, w = _i32(pos1, dev), _i32(pos2, dev), _f32(w, dev)
    len1 = 0 if pos1 is None else pos1.numel() // 3
    len2 = 0 if pos2 is None else pos2.numel() // 3
    _cabi.check(lib.mke_rel_step_sampled(
        ent.c, rel.c, _cabi.ptr(pos1), len1, kg1.c if kg1 is not None else None,
        _cabi.ptr(pos2), len2, kg2.c if kg2 is not None else None,
        int(K), int(seed) & (2 ** 64 - 1), int(step) & (2 ** 64 - 1), _cabi.ptr(w), float(pos_scale),
        _cabi.ptr(loss_accum), _cabi.ptr(neg_out), int(variant), _cabi.current_stream()))
    return pos1, pos2, w


def rel_step_structured(ent, rel, pos, neg_ent, neg_side, K, loss_accum, w=None, pos_scale=1.0, variant=0):
    """mke_rel_step_structured: fused phase 1 with caller-supplied negatives."""
    lib = _cabi.load()
    dev = ent.device
    pos, neg_ent, w = _i32(pos, dev), _i32(neg_ent, dev), _f32(w, dev)
    if neg_side is not None and not torch.is_tensor(neg_side):
        neg_side = torch.as_tensor(np.ascontiguousarray(neg_side, dtype=np.uint32).view(np.int32))
    if neg_side is not None:
        neg_side = neg_side.to(device=dev).contiguous()
    n = pos.numel() // 3
    _cabi.check(lib.mke_rel_step_structured(ent.c, rel.c, _cabi.ptr(pos), n, int(K), _cabi.ptr(neg_ent),
                                            _cabi.ptr(neg_side), _cabi.ptr(w), float(pos_scale),
                                            _cabi.ptr(loss_accum), int(variant), _cabi.current_stream()))
    return pos, neg_ent, neg_side, w


def neg_keep_owned(neg_ent, K, n_shards, shard_split, my_shard, dummy_id):
    """mke_neg_keep_owned: in place; returns the ownership masks [n] (int32 bit masks)."""
    lib = _cabi.load()
    n = neg_ent.numel() // K
    valid = torch.empty(n, dtype=torch.int32, device=neg_ent.device)
    _cabi.check(lib.mke_neg_keep_owned(neg_ent.data_ptr(), n, int(K), int(n_shards), int(shard_split), int(my_shard),
                                       int(dummy_id), valid.data_ptr(), _cabi.current_stream()))
    return valid


def neg_keep_owned_compact(neg_ent, neg_side, K, n_shards, shard_split, my_shard, dummy_id):
    """mke_neg_keep_owned2: in place (ids AND side words); this rank's negatives first; returns the masks
    low_ones(count) [n]."""
    lib = _cabi.load()
    n = neg_ent.numel() // K
    valid = torch.empty(n, dtype=torch.int32, device=neg_ent.device)
    _cabi.check(lib.mke_neg_keep_owned2(neg_ent.data_ptr(), neg_side.data_ptr(), n, int(K), int(n_shards),
                                        int(shard_split), int(my_shard), int(dummy_id), valid.data_ptr(),
                                        _cabi.current_stream()))
    return valid


def rel_step_owned(ent, rel, pos, neg_ent, neg_side, neg_valid, own_lo, own_hi, K, loss_accum, variant=0, compact=False):
    """mke_rel_step_structured4: one rank's share of a step under "negatives where they live" (compact: the
    batch went through neg_keep_owned_compact)."""
    lib = _cabi.load()
    dev = ent.device
    pos, neg_ent = _i32(pos, dev), _i32(neg_ent, dev)
    n = pos.numel() // 3
    _cabi.check(lib.mke_rel_step_structured4(ent.c, rel.c, _cabi.ptr(pos), n, None, 0, int(K), _cabi.ptr(neg_ent),
                                             _cabi.ptr(neg_side), _cabi.ptr(neg_valid), int(bool(compact)), int(own_lo),
                                             int(own_hi), None, 1.0, _cabi.ptr(loss_accum), int(variant),
                                             _cabi.current_stream()))


def apply_adagrad(table, acc, lr):
    lib = _cabi.load()
    _cabi.check(lib.mke_rows_apply_adagrad(table.c, acc.data_ptr(), float(lr), _cabi.current_stream()))


def apply_adagrad_pair(table_a, acc_a, lr_a, table_b, acc_b, lr_b):
    """mke_rows_apply_adagrad_pair: phase 2 of two tables in one launch."""
    lib = _cabi.load()
    _cabi.check(lib.mke_rows_apply_adagrad_pair(table_a.c, acc_a.data_ptr(), float(lr_a), table_b.c, acc_b.data_ptr(),
                                                float(lr_b), _cabi.current_stream()))


def sample_distinct(n, count, seed, draw, device="cuda"):
    """mke_sample_distinct: random.sample(range(n), count) as int64 indices on the device (no sort)"""
    lib = _cabi.load()
    out = torch.empty(int(count), dtype=torch.int32, device=device)
    _cabi.check(lib.mke_sample_distinct(int(n), int(count), int(seed) & (2 ** 64 - 1), int(draw) & (2 ** 64 - 1),
                                        out.data_ptr(), _cabi.current_stream()))
    return out.long()


def sample_uniform(pos1, kg1, pos2, kg2, K, seed, step, device="cuda"):
    """mke_sample_uniform: the negatives the fused kernel would draw, as [(len1+len2)*K, 3]."""
    lib = _cabi.load()
    pos1, pos2 = _i32(pos1, device), _i32(pos2, device)
    len1 = 0 if pos1 is None else pos1.numel() // 3
    len2 = 0 if pos2 is None else pos2.numel() // 3
    out = torch.empty((len1 + len2) * K, 3, dtype=torch.int32, device=device)
    _cabi.check(lib.mke_sample_uniform(_cabi.ptr(pos1), len1, kg1.c if kg1 is not None else None,
                                       _cabi.ptr(pos2), len2, kg2.c if kg2 is not None else None,
                                       int(K), int(seed) & (2 ** 64 - 1), int(step) & (2 ** 64 - 1),
                                       out.data_ptr(), _cabi.current_stream()))
    return out


def sample_structured(pos1, kg1, pos2, kg2, K, seed, step, device="cuda"):
    """mke_sample_structured: (neg_ent [n,K] int32, neg_side [n] int32 bit masks)."""
    lib = _cabi.load()
    pos1, pos2 = _i32(pos1, device), _i32(pos2, device)
    len1 = 0 if pos1 is None else pos1.numel() // 3
    len2 = 0 if pos2 is None else pos2.numel() // 3
    ne = torch.empty(len1 + len2, K, dtype=torch.int32, device=device)
    ns = torch.empty(len1 + len2, dtype=torch.int32, device=device)
    _cabi.check(lib.mke_sample_structured(_cabi.ptr(pos1), len1, kg1.c if kg1 is not None else None,
                                          _cabi.ptr(pos2), len2, kg2.c if kg2 is not None else None,
                                          int(K), int(seed) & (2 ** 64 - 1), int(step) & (2 ** 64 - 1),
                                          ne.data_ptr(), ns.data_ptr(), _cabi.current_stream()))
    return ne, ns


def sample_attribute_heads(pos1, kg1, pos2, kg2, K, seed, step, index_base=0, device="cuda"):
    """mke_sample_attribute_heads: corrupted heads [(len1+len2), K] for (h, a, v) positives."""
    lib = _cabi.load()
    pos1, pos2 = _i32(pos1, device), _i32(pos2, device)
    len1 = 0 if pos1 is None else pos1.numel() // 3
    len2 = 0 if pos2 is None else pos2.numel() // 3
    out = torch.empty(len1 + len2, K, dtype=torch.int32, device=device)
    _cabi.check(lib.mke_sample_attribute_heads(_cabi.ptr(pos1), len1, kg1.c if kg1 is not None else None,
                                               _cabi.ptr(pos2), len2, kg2.c if kg2 is not None else None, int(K),
                                               int(seed) & (2 ** 64 - 1), int(step) & (2 ** 64 - 1), int(index_base),
                                               out.data_ptr(), _cabi.current_stream()))
    return out


def align_fwd_bwd(shared, name, rv, av, idx, loss_accum, name_weight=1.0, scale=1.0):
    """mke_align_fwd_bwd: ITC cross-view alignment term + backward for a batch of entity ids."""
    lib = _cabi.load()
    idx = _i32(idx, shared.device)
    _cabi.check(lib.mke_align_fwd_bwd(shared.c, name.c, rv.c, av.c, idx.data_ptr(), idx.numel(), float(name_weight),
                                      float(scale), _cabi.ptr(loss_accum), _cabi.current_stream()))
    return idx
