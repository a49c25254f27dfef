"""Relation-view training on the device: the state and step/epoch drivers behind
``MultiKE.train_relation_view_1epo`` (MultiKE_model.py:291-317).

What the reference does per step and where it lives here:
  base/batch.py:33-54    batch = kg1 slice ++ kg2 slice, sizes by KG share   -> step_slices()
  base/batch.py:86-116   K negatives per positive, filtered                  -> on device, inside
                                                                               mke_rel_step_sampled
  MultiKE_model.py:123-131 + losses.py:4-12  gathers, score, loss, backward  -> mke_rel_step_sampled
  MultiKE_model.py:15-31  Adagrad (own accumulators per graph)               -> mke_rows_apply_adagrad
  MultiKE_model.py:311-315 loss bookkeeping, list shuffle                    -> epoch drivers below
No TF, no CPU fallback: every call goes through the C-ABI library (multike_b200/_cabi.py).
"""
import math

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
    """rv_ent_embeds + rel_embeds, their Adagrad slots, the device-resident triple lists of both
    KGs and the negative samplers."""

    def __init__(self, n_ent, n_rel, dim, triples1, triples2, ent_split, batch_size=5000, neg_num=10,
                 lr=0.001, seed=0, device="cuda", variant=0, ent_init=None, rel_init=None,
                 filter1=None, filter2=None, generator=None, pipelined=True):
        _cabi.load()
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
        self.kg1 = T.KGSampler(entity_base=0, n_entities=ent_split, triple_set=self.set1, device=device)
        self.kg2 = T.KGSampler(entity_base=ent_split, n_entities=n_ent - ent_split, triple_set=self.set2,
                               device=device)
        self.loss_acc = T.new_loss_accumulator(device)
        self.global_step = 0
        self._lib = _cabi.load()
        self.phase1_events = None  # optional list of (start, end) CUDA events around phase 1
        # Negatives of step s+1 are drawn on a second stream while step s trains: sampling reads
        # no embedding table (only triples, the filter set and the counter-based RNG), so it has
        # no dependence on the Adagrad update in flight.  Two buffers, used alternately.
        self.pipelined = bool(pipelined) and self.K > 0
        if self.pipelined:
            self._side = torch.cuda.Stream(device=self.device)
            self._neg = [(torch.empty(self.batch_size, self.K, dtype=torch.int32, device=self.device),
                          torch.empty(self.batch_size, dtype=torch.int32, device=self.device)) for _ in range(2)]
            self._ready = {}  # global step -> (buffer index, event) of negatives already drawn

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

    # -- one step -------------------------------------------------------------------------------
    def _phase1(self, pos1, len1, pos2, len2):
        stream = _cabi.current_stream()
        ev = None
        if self.phase1_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        _cabi.check(self._lib.mke_rel_step_sampled(
            self.ent.c, self.rel.c, pos1, len1, self.kg1.c, pos2, len2, self.kg2.c, self.K,
            self.seed & (2 ** 64 - 1), self.global_step, None, 1.0, self.loss_acc.data_ptr(), None,
            self.variant, stream))
        if ev is not None:
            ev[1].record()
            self.phase1_events.append(ev)

    def _sample_into(self, buf, pos1, len1, pos2, len2, step, stream):
        ne, ns = self._neg[buf]
        _cabi.check(self._lib.mke_sample_structured(
            pos1, len1, self.kg1.c, pos2, len2, self.kg2.c, self.K, self.seed & (2 ** 64 - 1), step,
            ne.data_ptr(), ns.data_ptr(), stream))

    def _phase1_presampled(self, buf, pos1, len1, pos2, len2):
        ne, ns = self._neg[buf]
        ev = None
        if self.phase1_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        _cabi.check(self._lib.mke_rel_step_structured2(
            self.ent.c, self.rel.c, pos1, len1, pos2, len2, self.K, ne.data_ptr(), ns.data_ptr(), None, 1.0,
            self.loss_acc.data_ptr(), self.variant, _cabi.current_stream()))
        if ev is not None:
            ev[1].record()
            self.phase1_events.append(ev)

    def _slice_ptrs(self, step_in_epoch):
        (a1, b1), (a2, b2) = self.step_slices(step_in_epoch)
        return self.triples1.data_ptr() + 12 * a1, b1 - a1, self.triples2.data_ptr() + 12 * a2, b2 - a2

    def _phase2(self):
        T.apply_adagrad_pair(self.ent, self.ent.adagrad_slot("relation"), self.lr,
                             self.rel, self.rel.adagrad_slot("relation"), self.lr)

    def step_resident(self, step_in_epoch, next_step_in_epoch=None):
        """One training step on positives already in HBM; returns the number of positives.
        `next_step_in_epoch` (pipelined mode): the step whose negatives are drawn meanwhile."""
        p1, len1, p2, len2 = self._slice_ptrs(step_in_epoch)
        if len1 + len2 == 0:
            return 0
        if not self.pipelined:
            self._phase1(p1, len1, p2, len2)
        else:
            main = torch.cuda.current_stream()
            ready = self._ready.pop(self.global_step, None)
            if ready is None:  # nothing drawn ahead (first step, epoch boundary): draw in line
                buf = self.global_step & 1
                self._sample_into(buf, p1, len1, p2, len2, self.global_step, main.cuda_stream)
            else:
                buf, ev = ready
                main.wait_event(ev)
            self._phase1_presampled(buf, p1, len1, p2, len2)
            if next_step_in_epoch is not None:
                q1, m1, q2, m2 = self._slice_ptrs(next_step_in_epoch)
                if m1 + m2 > 0:
                    # starts when phase 1 of this step has finished, i.e. overlaps the (HBM-bound)
                    # apply kernel; the other buffer was last read by the previous step's phase 1
                    fence = main.record_event()
                    self._side.wait_event(fence)
                    self._sample_into(buf ^ 1, q1, m1, q2, m2, self.global_step + 1, self._side.cuda_stream)
                    self._ready[self.global_step + 1] = (buf ^ 1, self._side.record_event())
        self._phase2()
        self.global_step += 1
        return len1 + len2

    def step_host(self, pos1_pinned, pos2_pinned, staging):
        """One step whose positives arrive in (pinned) HOST memory, loss read back to the host:
        what one queue.get() + session.run([loss, optimizer]) of MultiKE_model.py:302-310 costs a
        caller.  Returns (batch loss, number of positives)."""
        len1, len2 = pos1_pinned.shape[0], pos2_pinned.shape[0]
        if len1 + len2 == 0:
            return 0.0, 0
        d1, d2 = staging[0][:len1], staging[1][:len2]
        d1.copy_(pos1_pinned, non_blocking=True)
        d2.copy_(pos2_pinned, non_blocking=True)
        self.loss_acc.zero_()
        self._phase1(d1.data_ptr(), len1, d2.data_ptr(), len2)
        self._phase2()
        self.global_step += 1
        return float(self.loss_acc.item()), len1 + len2

    def make_staging(self):
        b1, b2 = split_batch(self.n1, self.n2, self.batch_size)
        return (torch.empty(b1, 3, dtype=torch.int32, device=self.device),
                torch.empty(b2, 3, dtype=torch.int32, device=self.device))

    # -- one epoch ------------------------------------------------------------------------------
    def train_epoch(self, shuffle=True, generator=None):
        """train_relation_view_1epo: all steps of one epoch from device-resident triples; the
        loss stays on the device until the end of the epoch.  Returns (avg loss, #positives)."""
        self.loss_acc.zero_()
        trained = 0
        steps = self.triple_steps
        for s in range(steps):
            trained += self.step_resident(s, s + 1 if s + 1 < steps else None)
        loss = float(self.loss_acc.item())
        if shuffle:  # MultiKE_model.py:314-315 random.shuffle of both lists
            self.triples1 = self.triples1[torch.randperm(self.n1, device=self.device, generator=generator)]
            self.triples2 = self.triples2[torch.randperm(self.n2, device=self.device, generator=generator)]
        return loss / max(trained, 1), trained
