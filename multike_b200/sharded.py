"""Multi-GPU relation view: one process per GPU of one NVLink/NVSwitch box.

  * the entity table (variable, gradient accumulator, touched flags, Adagrad slots) is ROW-SHARDED:
    row id lives on rank id % G at local row id // G.  Every rank maps every shard into its own
    address space (CUDA IPC, mke_ipc_*), and phase 1 gathers rows / reduces gradient rows straight
    through those peer pointers over NVLink -- the exchange is fused into the kernel, there is no
    all-to-all;
  * the relation table is small and dense: replicated; its gradient bucket (176 KB) is summed between
    phase 1 and phase 2 by a kernel that reads every rank's copy through its peer mapping, and the points
    at which ranks wait for each other are flag barriers in peer memory (csrc/mke_sharded.cu) -- the step
    loop is issued from C without a collective library or a host round trip; torch.distributed only
    carries the IPC handles at set-up and the epoch's loss sum;
  * a global step of G * batch_size positives is split by position: rank k trains positions
    [k n / G, (k+1) n / G) of the concatenated (kg1 slice ++ kg2 slice) batch and draws its
    negatives at RNG coordinate index_base + i, so G ranks draw what one GPU would draw for the
    whole batch: the result equals a single-GPU run with batch_size G * B up to fp32 summation order.

The pure index arithmetic lives in module-level functions so that it is testable on CPU (gloo).
"""
import ctypes
import os
import math

import numpy as np
import torch

from . import _cabi
from . import tables as T
from .relation_view import clipped_slice, split_batch

# ---------------------------------------------------------------------------------------------
# index arithmetic (CPU-testable)
# ---------------------------------------------------------------------------------------------


def shard_owner(ids, world, split=0):
    """(owning rank, local row) of global row ids; split > 0 selects the KG-block placement
    (mke_table_t.shard_split): KG1 = ids [0, split) on ranks [0, world/2), KG2 on the others."""
    ids = np.asarray(ids)
    if split <= 0:
        return ids % world, ids // world
    half = world // 2
    second = ids >= split
    x = np.where(second, ids - split, ids)
    return np.where(second, half, 0) + x % half, x // half


def local_rows(rows, rank, world, split=0):
    def part(n, r, g):
        return (n - r + g - 1) // g if n > r else 0
    if split <= 0:
        return part(rows, rank, world)
    half = world // 2
    return part(split, rank, half) if rank < half else part(rows - split, rank - half, half)


def rank_range(n, rank, world):
    """positions of a batch of n positives trained by `rank`"""
    return rank * n // world, (rank + 1) * n // world


def rank_parts(n1, n2, global_batch, step, rank, world, by_kg=False):
    """((start1, len1), (start2, len2), index_base): the pieces of the kg1 / kg2 triple lists that
    `rank` trains in global step `step`, and the position of its first positive in the global
    batch (= its RNG coordinate base).  by_kg: ranks [0, world/2) share the kg1 slice and the others
    the kg2 slice (goes with the KG-block placement: positives are trained where their rows live)."""
    b1, b2 = split_batch(n1, n2, global_batch)
    a1, e1 = clipped_slice(n1, b1, step)
    a2, e2 = clipped_slice(n2, b2, step)
    len1, len2 = e1 - a1, e2 - a2
    if by_kg:
        half = world // 2
        if rank < half:
            lo, hi = rank_range(len1, rank, half)
            return (a1 + lo, hi - lo), (a2, 0), lo
        lo, hi = rank_range(len2, rank - half, half)
        return (a1 + len1, 0), (a2 + lo, hi - lo), len1 + lo
    lo, hi = rank_range(len1 + len2, rank, world)
    s1, t1 = min(lo, len1), min(hi, len1)
    s2, t2 = max(lo, len1) - len1, max(hi, len1) - len1
    return (a1 + s1, t1 - s1), (a2 + s2, t2 - s2), lo


def group_parts(n1, n2, global_batch, step, rank, world):
    """"Negatives where they live" (KG-block placement, world >= 4): every rank of a KG's half of the
    ranks walks the WHOLE slice of that KG.  Returns (kg (1 or 2), (start, length) of the slice in the
    KG's triple list, (own_lo, own_hi) = the positions inside the slice whose positive terms this
    rank computes, index_base = position of the slice's first positive in the global batch)."""
    b1, b2 = split_batch(n1, n2, global_batch)
    a1, e1 = clipped_slice(n1, b1, step)
    a2, e2 = clipped_slice(n2, b2, step)
    half = world // 2
    if rank < half:
        return 1, (a1, e1 - a1), rank_range(e1 - a1, rank, half), 0
    return 2, (a2, e2 - a2), rank_range(e2 - a2, rank - half, half), e1 - a1


# ---------------------------------------------------------------------------------------------
# peer memory
# ---------------------------------------------------------------------------------------------


class _DevView:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


class PeerBuffer:
    """A cudaMalloc'ed block that every rank of the box can address (CUDA IPC)."""

    def __init__(self, shape, dtype, group):
        import torch.distributed as dist
        lib = _cabi.load()
        self.shape, self.dtype = tuple(shape), dtype
        itemsize = torch.empty((), dtype=dtype).element_size()
        nbytes = max(int(np.prod(shape)) * itemsize, 256)
        p = ctypes.c_void_p()
        _cabi.check(lib.mke_peer_alloc(nbytes, ctypes.byref(p)))
        self.ptr = p.value
        typestr = {torch.float32: "<f4", torch.uint8: "|u1", torch.int32: "<i4"}[dtype]
        self._view = _DevView(self.ptr, self.shape, typestr)
        self.tensor = torch.as_tensor(self._view, device="cuda")
        handle = ctypes.create_string_buffer(64)
        _cabi.check(lib.mke_ipc_export(self.ptr, handle))
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.peers = []
        for k in range(world):
            if k == rank:
                self.peers.append(self.ptr)
            else:
                q = ctypes.c_void_p()
                _cabi.check(lib.mke_ipc_open(handles[k], ctypes.byref(q)))
                self.peers.append(q.value)
        self._rank = rank

    def close(self):
        lib = _cabi.load()
        for k, q in enumerate(self.peers):
            if k != self._rank and q:
                lib.mke_ipc_close(q)
        self.peers = []


class ShardedEmbeddingTable:
    """mke_table_t with n_shards = world: this rank's rows of a row-sharded normalised table."""

    def __init__(self, rows, dim, normalised, group, init=None, name="", split=0, flags=None):
        import torch.distributed as dist
        self.rows, self.dim, self.normalised, self.name = int(rows), int(dim), bool(normalised), name
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        assert self.world in (2, 4, 8), "row sharding supports 2, 4 or 8 ranks"
        self.stride = T.padded_stride(dim)
        self.split = int(split)
        self.local_rows = local_rows(self.rows, self.rank, self.world, self.split)
        # every shard gets the same (maximal) allocation so that peer offsets never overrun
        alloc_rows = max(local_rows(self.rows, r, self.world, self.split) for r in range(self.world))
        # touched flags of remote rows would be single-byte stores over NVLink (measured: 17 us of a
        # 166 us phase 1 at G = 4), while a shard is small enough to be swept whole by phase 2:
        # flags only where (almost) all traffic is local
        self.flags = (self.world == 2 and self.split > 0) if flags is None else bool(flags)
        self._bufs = [PeerBuffer((alloc_rows, self.stride), torch.float32, group),
                      PeerBuffer((alloc_rows, self.stride), torch.float32, group)]
        if self.flags:
            self._bufs.append(PeerBuffer((alloc_rows,), torch.uint8, group))
        self.var, self.grad = self._bufs[0].tensor, self._bufs[1].tensor
        self.touched = self._bufs[2].tensor if self.flags else None
        self.device = self.var.device
        if init is not None:  # init is the GLOBAL [rows, dim] table; keep rows rank, rank + world, ...
            src = torch.as_tensor(np.asarray(init, dtype=np.float32)) if not torch.is_tensor(init) else init
            mine = src[torch.as_tensor(self.owned_ids())].to(self.device, torch.float32)
            self.var[: mine.shape[0], : self.dim] = mine
        self._slots = {}
        c = _cabi.MkeTable(var=self.var.data_ptr(), grad=self.grad.data_ptr(), touched=_cabi.ptr(self.touched),
                           rows=self.rows, stride=self.stride, dim=self.dim, normalised=int(self.normalised),
                           grad_replicas=1, n_shards=self.world, shard_rank=self.rank, shard_split=self.split)
        for k in range(self.world):
            c.peer_var[k], c.peer_grad[k] = self._bufs[0].peers[k], self._bufs[1].peers[k]
            c.peer_touched[k] = self._bufs[2].peers[k] if self.flags else None
        self._c = c
        self.grad_replicas = 1
        torch.cuda.synchronize()
        dist.barrier(group)

    def owned_ids(self):
        """global ids of this rank's rows, in local-row order"""
        ids = np.arange(self.rows)
        owner, local = shard_owner(ids, self.world, self.split)
        mine = ids[owner == self.rank]
        assert np.array_equal(local[owner == self.rank], np.arange(mine.size))
        return mine

    @property
    def c(self):
        return ctypes.byref(self._c)

    def adagrad_slot(self, slot):
        acc = self._slots.get(slot)
        if acc is None:
            acc = torch.full((max(self.var.shape[0], 1), self.stride), T.ADAGRAD_INIT, dtype=torch.float32,
                             device=self.device)
            self._slots[slot] = acc
        return acc

    def grad_sum(self):
        return self.grad

    def export(self, idx=None):
        """normalised rows by GLOBAL id (peer reads)"""
        lib = _cabi.load()
        if idx is None:
            idx = np.arange(self.rows, dtype=np.int32)
        idx_t = torch.as_tensor(np.asarray(idx, dtype=np.int32)).to(self.device).contiguous()
        out = torch.empty(idx_t.numel(), self.dim, dtype=torch.float32, device=self.device)
        _cabi.check(lib.mke_table_export(self.c, idx_t.data_ptr(), idx_t.numel(), out.data_ptr(),
                                         _cabi.current_stream()))
        return out

    def eval(self, session=None, idx=None):
        return self.export(idx).cpu().numpy()

    def raw_local(self):
        return self.var[: self.local_rows, : self.dim].detach().cpu().numpy()

    def close(self):
        for b in self._bufs:
            b.close()


class ShardedRelationView:
    """The relation view on G GPUs; every rank constructs it with the same arguments."""

    SLOT = "relation"

    def __init__(self, n_ent, n_rel, dim, triples1, triples2, ent_split, batch_size, neg_num, lr, seed, group,
                 ent_init=None, rel_init=None, filter1=None, filter2=None, rel_replicas=1, by_kg=True,
                 variant=None):
        import torch.distributed as dist
        self._lib = _cabi.load()
        # phase-1 schedule (include/multike_b200.h `variant`); MKE_SHARDED_VARIANT is an experiment knob
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        # measured on 4 x B200: with rows behind NVLink the one-wave kernel (all of a positive's remote
        # rows requested up front) beats the in-order row stream; without remote rows (2 ranks, one KG
        # each) the row stream is the faster one
        default_variant = "3" if (self.world == 2 and by_kg) else "0"
        self.variant = int(os.environ.get("MKE_SHARDED_VARIANT", default_variant)) if variant is None else int(variant)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.dim, self.K, self.lr, self.seed = int(dim), int(neg_num), float(lr), int(seed)
        self.batch_size = int(batch_size)              # per rank
        self.global_batch = self.batch_size * self.world
        # by_kg: KG-block placement + positives trained on the ranks of their own KG (kg1 ids are
        # [0, ent_split), kg2 ids [ent_split, n_ent)); otherwise plain id % G placement
        self.by_kg = bool(by_kg)
        self.ent = ShardedEmbeddingTable(n_ent, dim, True, group, init=ent_init, name="rv_ent_embeds",
                                         split=ent_split if self.by_kg else 0)
        self.rel = T.EmbeddingTable(n_rel, dim, True, self.device, init=rel_init, name="rel_embeds", flags=False,
                                    grad_replicas=rel_replicas)
        t1 = np.ascontiguousarray(triples1, dtype=np.int32).reshape(-1, 3)
        t2 = np.ascontiguousarray(triples2, dtype=np.int32).reshape(-1, 3)
        self.triples1, self.triples2 = torch.from_numpy(t1).to(self.device), torch.from_numpy(t2).to(self.device)
        self.n1, self.n2 = t1.shape[0], t2.shape[0]
        self.set1 = T.TripleSet(t1 if filter1 is None else filter1, self.device)
        self.set2 = T.TripleSet(t2 if filter2 is None else filter2, self.device)
        self.kg1 = T.KGSampler(entity_base=0, n_entities=ent_split, triple_set=self.set1, device=self.device)
        self.kg2 = T.KGSampler(entity_base=ent_split, n_entities=n_ent - ent_split, triple_set=self.set2,
                               device=self.device)
        # most positives one rank can get in a step (by_kg: a KG's share of the global batch over
        # half of the ranks, which exceeds batch_size for the larger KG)
        b1, b2 = split_batch(self.n1, self.n2, self.global_batch)
        cap = (max(b1, b2) // max(self.world // 2, 1) + 2) if self.by_kg else self.batch_size + 2
        # negatives where they live (include/multike_b200.h, mke_neg_keep_owned): with more than one
        # rank per KG every rank walks its KG's whole slice and scores the negatives it owns
        self.owner_negs = self.by_kg and self.world >= 4 and self.K > 0 and \
            os.environ.get("MKE_OWNER_NEGS", "1") == "1"
        if self.owner_negs:
            cap = max(b1, b2) + 2
            half = self.world // 2
            self._dummy_row = self.rank if self.rank < half else ent_split + (self.rank - half)
        # two buffer sets: the negatives of step s + 1 are drawn while step s is exchanged and applied
        self._neg_ent = torch.empty(2, cap * max(self.K, 1), dtype=torch.int32, device=self.device)
        self._neg_side = torch.empty(2, cap, dtype=torch.int32, device=self.device)
        self._neg_valid = torch.empty(2, cap, dtype=torch.int32, device=self.device)
        self.ent_split = int(ent_split)
        # the next step's negatives are drawn on a second stream under the exchange and phase 2 (C-side events)
        self.draw_ahead = os.environ.get("MKE_DRAW_AHEAD", "1") == "1"
        self._side = torch.cuda.Stream(device=self.device)
        self.loss_acc = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._step_loss = torch.zeros(max(self.triple_steps, 1), dtype=torch.float64, device=self.device)
        self.global_step = 0
        # peer-mapped exchange buffer of the relation gradient bucket and barrier words (csrc/mke_sharded.cu)
        self._xchg = PeerBuffer((2, self.rel.rows, self.rel.stride), torch.float32, group)
        self._sync = PeerBuffer((16,), torch.int32, group)
        self._barrier_seq = ctypes.c_uint32(0)
        self._host = self._host_loss = None
        v = self._view = _cabi.MkeRelShardedView()
        v.ent, v.rel = ctypes.pointer(self.ent._c), ctypes.pointer(self.rel._c)
        v.ent_acc = self.ent.adagrad_slot(self.SLOT).data_ptr()
        v.rel_acc = self.rel.adagrad_slot(self.SLOT).data_ptr()
        v.lr = self.lr
        v.triples1, v.triples2, v.n1, v.n2 = self.triples1.data_ptr(), self.triples2.data_ptr(), self.n1, self.n2
        v.kg1, v.kg2 = ctypes.pointer(self.kg1._c), ctypes.pointer(self.kg2._c)
        v.global_batch, v.K, v.seed = self.global_batch, self.K, self.seed & (2 ** 64 - 1)
        v.world, v.rank, v.by_kg, v.owner_negs = self.world, self.rank, int(self.by_kg), int(self.owner_negs)
        v.dummy_row = self._dummy_row if self.owner_negs else 0
        v.variant = self.variant
        for k in range(2):
            v.neg_ent[k], v.neg_side[k] = self._neg_ent[k].data_ptr(), self._neg_side[k].data_ptr()
            v.neg_valid[k] = self._neg_valid[k].data_ptr()
        for k in range(self.world):
            v.xchg[k], v.sync[k] = self._xchg.peers[k], self._sync.peers[k]
        torch.cuda.synchronize()
        dist.barrier(group)

    @property
    def triple_steps(self):
        return int(math.ceil((self.n1 + self.n2) / self.global_batch))

    def set_neighbours(self, nb1, nb2):
        """truncated-eps candidate lists (base/batch.py:119-150, MultiKE_CSL.py:89-99): the same table on every rank"""
        self.kg1.set_neighbours(nb1, self.device)
        self.kg2.set_neighbours(nb2, self.device)

    def plan(self, step_in_epoch):
        """(positives of kg1, of kg2, index_base, own_lo, own_hi, positives this rank answers for) of one global
        step for this rank -- what csrc/mke_sharded.cu::make_plan computes; kept for the CPU tests"""
        if self.owner_negs:
            kg_no, (a, ln), (lo, hi), base = group_parts(self.n1, self.n2, self.global_batch, step_in_epoch,
                                                         self.rank, self.world)
            return ((a, ln), (0, 0), base, lo, hi, hi - lo) if kg_no == 1 else ((0, 0), (a, ln), base, lo, hi, hi - lo)
        (a1, l1), (a2, l2), base = rank_parts(self.n1, self.n2, self.global_batch, step_in_epoch, self.rank, self.world,
                                               by_kg=self.by_kg)
        return (a1, l1), (a2, l2), base, 0, 0x7fffffff, l1 + l2

    def train_steps(self, first_step, n_steps, host_fed=False):
        """n_steps consecutive GLOBAL steps starting at step `first_step` of the epoch (wrapping), issued by ONE
        library call on every rank; returns the number of positives this rank answers for.  The per-step
        losses of this rank are added to loss_acc.  host_fed: every step this rank's positives are copied in
        from pinned HOST memory and its share of the step loss is copied back (host_losses)."""
        if n_steps <= 0:
            return 0
        if self._step_loss.numel() < n_steps:
            self._step_loss = torch.zeros(n_steps, dtype=torch.float64, device=self.device)
        self._step_loss[:n_steps].zero_()
        v = self._view
        v.step_loss = self._step_loss.data_ptr()
        if host_fed:
            if self._host is None:
                cap = self._neg_side.shape[1]
                self._host = (self.triples1.cpu().pin_memory(), self.triples2.cpu().pin_memory(),
                              [torch.empty(cap * 3, dtype=torch.int32, device=self.device) for _ in range(4)])
            if self._host_loss is None or self._host_loss.numel() < n_steps:
                self._host_loss = torch.zeros(n_steps, dtype=torch.float64).pin_memory()
            v.host_triples1, v.host_triples2 = self._host[0].data_ptr(), self._host[1].data_ptr()
            for k in range(2):
                v.stage1[k], v.stage2[k] = self._host[2][k].data_ptr(), self._host[2][2 + k].data_ptr()
            v.host_step_loss = self._host_loss.data_ptr()
        else:
            v.host_triples1 = v.host_triples2 = v.host_step_loss = None
        main = torch.cuda.current_stream()
        mine = ctypes.c_int64(0)
        side = self._side.cuda_stream if self.draw_ahead else None
        _cabi.check(self._lib.mke_rel_sharded_train_steps(
            ctypes.byref(self._view), int(first_step), int(n_steps), self.global_step, ctypes.byref(self._barrier_seq),
            ctypes.byref(mine), main.cuda_stream, side))
        self.loss_acc += self._step_loss[:n_steps].sum()
        self.global_step += n_steps
        return int(mine.value)

    def positives_walked(self, first_step, n_steps):
        """positives whose ids this rank reads in steps first_step .. (what a host-fed call copies in)"""
        tot = 0
        for k in range(n_steps):
            (a1, l1), (a2, l2) = self.plan((first_step + k) % self.triple_steps)[:2]
            tot += l1 + l2
        return tot

    @property
    def host_losses(self):
        return self._host_loss

    def step(self, step_in_epoch):
        """one global step; returns the number of positives this rank answers for"""
        return self.train_steps(step_in_epoch, 1)

    def shuffle(self, seed):
        """MultiKE_model.py:314-315 random.shuffle of both triple lists: the SAME permutation on every rank
        (every rank holds the whole lists), drawn on the host from `seed`"""
        gen = torch.Generator().manual_seed(int(seed))
        p1 = torch.randperm(self.n1, generator=gen).to(self.device)
        p2 = torch.randperm(self.n2, generator=gen).to(self.device)
        self.triples1.copy_(self.triples1[p1])
        self.triples2.copy_(self.triples2[p2])
        if self._host is not None:
            self._host[0].copy_(self.triples1)
            self._host[1].copy_(self.triples2)

    def train_epoch(self, shuffle_seed=None):
        """train_relation_view_1epo on G GPUs: (average loss per positive, positives) over all ranks; the lists
        are shuffled afterwards when a seed is given (the same one on every rank)"""
        import torch.distributed as dist
        self.loss_acc.zero_()
        trained = self.train_steps(0, self.triple_steps)
        tot = torch.cat([self.loss_acc, torch.tensor([float(trained)], dtype=torch.float64, device=self.device)]).cpu()
        if dist.get_backend(self.group) != "gloo":
            tot = tot.to(self.device)
        dist.all_reduce(tot, group=self.group)
        if shuffle_seed is not None:
            self.shuffle(shuffle_seed)
        return float(tot[0]) / max(float(tot[1]), 1.0), int(tot[1])

    def close(self):
        torch.cuda.synchronize()
        self._xchg.close()
        self._sync.close()
        self.ent.close()
