"""Relation-view training on the device: the state and step/epoch drivers behind
``MultiKE.train_relation_view_1epo`` (MultiKE_model.py:291-317).

What the reference does per step and where it lives here:
  base/batch.py:33-54    batch = kg1 slice ++ kg2 slice, sizes by KG share   -> mke_rel_train_steps
  base/batch.py:86-116   K negatives per positive, filtered                  -> mke_sample_structured
                                                                               (or inside phase 1)
  MultiKE_model.py:123-131 + losses.py:4-12  gathers, score, loss, backward  -> phase 1 kernel
  MultiKE_model.py:15-31  Adagrad (own accumulators per graph)               -> phase 2 kernel
  MultiKE_model.py:302-315 step loop, loss bookkeeping, list shuffle         -> mke_rel_train_steps,
                                                                               train_epoch below
No TF, no CPU fallback: every call goes through the C-ABI library (multike_b200/_cabi.py).
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _cabi
from . import tables as T


def split_batch(n1, n2, batch_size):
    """base/batch.py:36-37 -- kg1's share is floored, kg2 takes the rest."""
    b1 = int(n1 / (n1 + n2) * batch_size)
    return b1, batch_size - b1


def clipped_slice(n, bs, step):
    """base/batch.py:45-54 (is_fixed_size=False): [step*bs, (step+1)*bs) clipped at the list end."""
    start = min(step * bs, n)
    return start, min(start + bs, n)


class RelationView:
    """rv_ent_embeds + rel_embeds, their Adagrad slots, the triple lists of both KGs (device
    resident, and optionally a pinned host copy for host-fed steps) and the negative samplers."""

    SLOT = "relation"  # one Adagrad accumulator set per loss graph (MultiKE_model.py:28-31)

    def __init__(self, n_ent, n_rel, dim, triples1, triples2, ent_split, batch_size=5000, neg_num=10,
                 lr=0.001, seed=0, device="cuda", variant=4, ent_init=None, rel_init=None,
                 filter1=None, filter2=None, generator=None, pipelined=True, entities1=None, entities2=None,
                 persist_chunk=None):
        self._lib = _cabi.load()
        self.device = torch.device(device)
        self.dim, self.batch_size, self.K, self.lr = int(dim), int(batch_size), int(neg_num), float(lr)
        self.seed, self.variant = int(seed), int(variant)
        if ent_init is None:
            ent_init = T.xavier_truncated_normal(n_ent, dim, generator)
        if rel_init is None:
            rel_init = T.xavier_truncated_normal(n_rel, dim, generator)
        # base/initializers.py:22-26 with is_l2_norm=True (MultiKE_model.py:92-95)
        self.ent = T.EmbeddingTable(n_ent, dim, True, device, init=ent_init, name="rv_ent_embeds", grad_replicas=1)
        self.rel = T.EmbeddingTable(n_rel, dim, True, device, init=rel_init, name="rel_embeds")
        t1 = np.ascontiguousarray(triples1, dtype=np.int32).reshape(-1, 3)
        t2 = np.ascontiguousarray(triples2, dtype=np.int32).reshape(-1, 3)
        self.triples1 = torch.from_numpy(t1).to(self.device)
        self.triples2 = torch.from_numpy(t2).to(self.device)
        self.n1, self.n2 = t1.shape[0], t2.shape[0]
        # filter set = relation_triples_set incl. swapped sup triples (base/kg.py:59,134)
        self.set1 = T.TripleSet(t1 if filter1 is None else filter1, device)
        self.set2 = T.TripleSet(t2 if filter2 is None else filter2, device)
        # candidate pools = kg.entities_list (base/batch.py:40-41); a contiguous id range needs no list
        self.kg1 = self._pool(entities1, 0, ent_split, self.set1)
        self.kg2 = self._pool(entities2, ent_split, n_ent - ent_split, self.set2)
        self.global_step = 0
        # Negatives of step s+1 are drawn on a second stream while step s trains: sampling reads
        # no embedding table (only triples, the filter set and the counter-based RNG), so it has
        # no dependence on the Adagrad update in flight.  Two buffers, used alternately.
        self.pipelined = bool(pipelined) and self.K > 0
        self._side = torch.cuda.Stream(device=self.device)
        self._neg = None
        if self.pipelined:
            self._neg = [(torch.empty(self.batch_size * self.K, dtype=torch.int32, device=self.device),
                          torch.empty(self.batch_size, dtype=torch.int32, device=self.device)) for _ in range(2)]
        self._step_loss = torch.zeros(max(self.triple_steps, 1), dtype=torch.float64, device=self.device)
        self._last_steps = 0
        self._host_loss = None
        self._host_triples = None
        self._stage = None
        # variant 4: persistent step kernel -- one cooperative launch per `persist_chunk` steps
        self._persist_ws = self._flag_src = None
        self.persist_chunk = 0
        if self.variant == 4 and self.pipelined:
            self.persist_chunk = min(int(persist_chunk or os.environ.get('MKE_PERSIST_STEPS', 0)) or 128, 128)
            nbytes = int(self._lib.mke_rel_persist_workspace_bytes(self.n1, self.n2, self.batch_size, self.persist_chunk))
            assert nbytes > 0
            self._persist_ws = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            self._flag_src = torch.arange(256, dtype=torch.int32).pin_memory()
        self._view = _cabi.MkeRelView()
        self._fill_view()

    def _pool(self, entities, base, count, triple_set):
        if entities is not None:
            e = np.asarray(entities, dtype=np.int64)
            if e.size and np.array_equal(e, np.arange(e[0], e[0] + e.size)):
                base, count, entities = int(e[0]), int(e.size), None
        if entities is None:
            return T.KGSampler(entity_base=base, n_entities=count, triple_set=triple_set, device=self.device)
        return T.KGSampler(entity_list=entities, triple_set=triple_set, device=self.device)

    # -- C view -----------------------------------------------------------------------------
    def _fill_view(self):
        v = self._view
        v.ent, v.rel = ctypes.pointer(self.ent._c), ctypes.pointer(self.rel._c)
        v.ent_acc = self.ent.adagrad_slot(self.SLOT).data_ptr()
        v.rel_acc = self.rel.adagrad_slot(self.SLOT).data_ptr()
        v.lr = self.lr
        v.triples1, v.triples2 = self.triples1.data_ptr(), self.triples2.data_ptr()
        v.n1, v.n2 = self.n1, self.n2
        v.kg1, v.kg2 = ctypes.pointer(self.kg1._c), ctypes.pointer(self.kg2._c)
        v.batch_size, v.K, v.seed, v.variant = self.batch_size, self.K, self.seed & (2 ** 64 - 1), self.variant
        for k in range(2):
            v.neg_ent[k] = self._neg[k][0].data_ptr() if self._neg else None
            v.neg_side[k] = self._neg[k][1].data_ptr() if self._neg else None
        v.step_loss = self._step_loss.data_ptr()
        if self._persist_ws is not None:
            v.persist_ws, v.persist_ws_bytes = self._persist_ws.data_ptr(), self._persist_ws.numel()
            v.persist_chunk, v.persist_flag_src = self.persist_chunk, self._flag_src.data_ptr()

    # -- bookkeeping of the reference drivers -------------------------------------------------
    @property
    def triple_steps(self):
        """MultiKE_CSL.py:40: ceil(#triples / batch_size)."""
        return int(math.ceil((self.n1 + self.n2) / self.batch_size))

    def step_slices(self, step):
        b1, b2 = split_batch(self.n1, self.n2, self.batch_size)
        return clipped_slice(self.n1, b1, step), clipped_slice(self.n2, b2, step)

    def set_neighbours(self, nb1, nb2):
        """truncated-eps candidate lists (base/batch.py:119-150, MultiKE_CSL.py:89-99)."""
        self.kg1.set_neighbours(nb1, self.device)
        self.kg2.set_neighbours(nb2, self.device)

    def use_host_triples(self, pinned1=None, pinned2=None):
        """Host-fed mode: every step's positives are copied from pinned HOST memory (what the
        reference's queue.get() + feed_dict hands to session.run) and every step's loss is copied
        back to a pinned host buffer.  Default source: pinned copies of the current lists."""
        b1, b2 = split_batch(self.n1, self.n2, self.batch_size)
        p1 = self.triples1.cpu().pin_memory() if pinned1 is None else pinned1
        p2 = self.triples2.cpu().pin_memory() if pinned2 is None else pinned2
        assert p1.is_pinned() and p2.is_pinned() and p1.dtype == torch.int32 and p2.dtype == torch.int32
        self._host_triples = (p1, p2)
        if self._stage is None:
            self._stage = [torch.empty(max(b, 1) * 3, dtype=torch.int32, device=self.device)
                           for b in (b1, b1, b2, b2)]

    # -- steps --------------------------------------------------------------------------------
    def train_steps(self, first_step, n_steps, host_fed=False):
        """n_steps consecutive steps starting at step `first_step` of the epoch (wrapping at the
        epoch end), issued by ONE library call.  Per-step losses are left in self.step_losses
        (device) and, host-fed, in self.host_losses (pinned).  Returns #positives trained."""
        if n_steps <= 0:
            return 0
        if self._step_loss.numel() < n_steps:
            self._step_loss = torch.zeros(n_steps, dtype=torch.float64, device=self.device)
            self._view.step_loss = self._step_loss.data_ptr()
        v = self._view   # (step_loss[:n_steps] is zeroed by the library call)
        if host_fed:
            if self._host_triples is None:
                self.use_host_triples()
            if self._host_loss is None or self._host_loss.numel() < n_steps:
                self._host_loss = torch.zeros(n_steps, dtype=torch.float64).pin_memory()
            v.host_triples1, v.host_triples2 = self._host_triples[0].data_ptr(), self._host_triples[1].data_ptr()
            v.stage1[0], v.stage1[1] = self._stage[0].data_ptr(), self._stage[1].data_ptr()
            v.stage2[0], v.stage2[1] = self._stage[2].data_ptr(), self._stage[3].data_ptr()
            v.host_step_loss = self._host_loss.data_ptr()
        else:
            v.host_triples1 = v.host_triples2 = None
            v.host_step_loss = None
        main = torch.cuda.current_stream()
        positives = ctypes.c_int64(0)
        _cabi.check(self._lib.mke_rel_train_steps(ctypes.byref(v), int(first_step), int(n_steps),
                                                  self.global_step, ctypes.byref(positives), main.cuda_stream,
                                                  self._side.cuda_stream))
        self.global_step += n_steps
        self._last_steps = n_steps
        return int(positives.value)

    def persist_trace(self, n_steps=None):
        """Device-side phase stamps of the LAST persistent launch (variant 4), in nanoseconds of
        %globaltimer: tensor [2 * n + 2] = launch start, after the negatives of the first step, then
        after phase 1 and after phase 2 of each of the launch's n steps."""
        assert self._persist_ws is not None
        n = self.persist_chunk if n_steps is None else int(n_steps)
        return self._persist_ws[2048: 2048 + 8 * (2 * n + 2)].view(torch.int64).clone()

    @property
    def step_losses(self):
        """Device tensor [n_steps] of the per-step batch losses of the last train_steps call."""
        return self._step_loss[: self._last_steps]

    @property
    def host_losses(self):
        """Pinned host tensor of the same (host-fed calls); valid after a stream synchronise."""
        return self._host_loss[: self._last_steps]

    def step_resident(self, step_in_epoch):
        """One training step on positives already in HBM; returns the number of positives."""
        return self.train_steps(step_in_epoch, 1)

    # -- one epoch ------------------------------------------------------------------------------
    def train_epoch(self, shuffle=True, generator=None, host_fed=False):
        """train_relation_view_1epo: all steps of one epoch; the losses stay on the device until
        the end of the epoch.  Returns (avg loss per positive, #positives)."""
        trained = self.train_steps(0, self.triple_steps, host_fed=host_fed)
        loss = float(self.step_losses.sum().item())
        if shuffle:  # MultiKE_model.py:314-315 random.shuffle of both lists (in place: same buffers)
            self.triples1.copy_(self.triples1[torch.randperm(self.n1, device=self.device, generator=generator)])
            self.triples2.copy_(self.triples2[torch.randperm(self.n2, device=self.device, generator=generator)])
            if self._host_triples is not None and host_fed:
                self._host_triples[0].copy_(self.triples1)
                self._host_triples[1].copy_(self.triples2)
        return loss / max(trained, 1), trained
