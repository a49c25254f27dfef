"""Full multi-view training on row-sharded entity tables (BASELINE configs[3]: DBP-YG-100K, SSL mode, entity tables
row-sharded over the GPUs of one box).

Every rank of the group constructs the same driver with the same arguments and calls the same methods in the same
order (one process per GPU; the torch / Python RNGs are seeded alike, so every rank draws the same batches):

  * the three trainable entity tables (rv_ent_embeds, av_ent_embeds, ent_embeds) are ShardedEmbeddingTables: every
    rank holds 1/G of the rows, of their gradient rows and of their Adagrad slots, mapped into every other rank's
    address space (CUDA IPC);
  * the relation view -- the path with the 20 000-positive batches -- is ShardedRelationView: data parallel, rows
    gathered and gradient rows reduced through peer pointers inside phase 1 (csrc/mke_sharded.cu);
  * the small-batch graphs (attribute CNN, cross-KG inference, ITC, SSL mapping: 5 000 rows per step, coupled through
    batch-wide l2-norms) run on STAGED rows (csrc/mke_stage.cu): every rank reads the batch's rows through the
    owners' mappings into a plain local table, runs the unchanged single-GPU kernels on it, adds the gradient rows
    of the ids it OWNS to its shard and applies Adagrad there.  The compute of these ~100 us steps is replicated;
    what is sharded is the state, and nothing but row reads crosses NVLink.  Two flag barriers per step order
    "everyone has staged" before "anyone updates" and back;
  * small dense parameters (rel_embeds, attr_embeds, the three CNN weight sets, the mappings) are replicated and see
    the same gradients on every rank; they are re-broadcast from rank 0 once per epoch so that float-atomic
    rounding cannot let the replicas drift.

The result equals the single-GPU drivers (refapi.drivers) up to fp32 summation order; tests/multi_gpu_model_check.py
compares the per-epoch losses and the evaluation of both.
"""
import ctypes
import random
import time

import torch

from . import _cabi
from . import tables as T
from .refapi import drivers as D
from .refapi.MultiKE_model import _neighbour_matrix, _triples
from .sharded import PeerBuffer, ShardedEmbeddingTable, ShardedRelationView


class PeerBarrier:
    """mke_peer_barrier: a flag barrier of the group's GPUs as a kernel launch on the current stream"""

    def __init__(self, group):
        import torch.distributed as dist
        self._lib = _cabi.load()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self._buf = PeerBuffer((16,), torch.int32, group)
        self._ptrs = (ctypes.c_void_p * 8)(*([ctypes.c_void_p(p) for p in self._buf.peers] + [None] * (8 - self.world)))
        self.seq = 0

    def wait(self):
        self.seq += 1
        _cabi.check(self._lib.mke_peer_barrier(self._ptrs, self.world, self.rank, self.seq, _cabi.current_stream()))

    def close(self):
        self._buf.close()


class StagedTable:
    """A row-sharded table behind the table interface the refapi model uses: reads (export / eval) go through the
    peer mappings; training steps stage() the rows of their batch and commit() the gradient rows they produced."""

    def __init__(self, sharded, cap):
        self.sh = sharded
        self.rows, self.dim, self.stride, self.normalised = sharded.rows, sharded.dim, sharded.stride, sharded.normalised
        self.device = sharded.device
        self._lib = _cabi.load()
        self.staged = T.EmbeddingTable(cap, self.dim, self.normalised, self.device, flags=False, grad_replicas=1)
        self._arange = torch.arange(cap, dtype=torch.int32, device=self.device)

    def stage(self, ids):
        """rows `ids` (global, repeats allowed) -> rows 0 .. n-1 of the staged table; returns (table, local index)"""
        n = ids.numel()
        _cabi.check(self._lib.mke_table_stage_rows(self.sh.c, ids.data_ptr(), n, self.staged.c, _cabi.current_stream()))
        return self.staged, self._arange[:n]

    def commit(self, ids, slot, lr):
        """gradient rows of the staged table -> this rank's rows, then phase 2 on the shard"""
        _cabi.check(self._lib.mke_table_commit_grads(self.sh.c, ids.data_ptr(), ids.numel(), self.staged.c,
                                                     _cabi.current_stream()))
        _cabi.check(self._lib.mke_rows_apply_adagrad(self.sh.c, self.sh.adagrad_slot(slot).data_ptr(), float(lr),
                                                     _cabi.current_stream()))

    def export(self, idx=None):
        return self.sh.export(idx)

    def eval(self, session=None, idx=None):
        return self.sh.eval(session, idx)


class StagedConstant:
    """A replicated constant table (name / literal vectors) staged by the same index vector as the sharded ones"""

    def __init__(self, table, cap):
        self.table = table
        self._lib = _cabi.load()
        self.staged = T.EmbeddingTable(cap, table.dim, table.normalised, table.device, trainable=False)

    def stage(self, ids):
        _cabi.check(self._lib.mke_table_stage_rows(self.table.c, ids.data_ptr(), ids.numel(), self.staged.c,
                                                   _cabi.current_stream()))
        return self.staged


class _ShardedModel:
    """Mixin in front of refapi.drivers.MultiKE_CV / MultiKE_Late: same schedule, same print lines, sharded state."""

    def __init__(self, data, args, predicate_align_model, group=None):
        import torch.distributed as dist
        self.group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        seed = int(getattr(args, "seed", 0))
        torch.manual_seed(seed)           # every rank draws the same batches (torch.randperm on the device ...
        torch.cuda.manual_seed(seed)
        random.seed(seed)                 # ... and random.shuffle of the attribute triple lists)
        self._barrier = PeerBarrier(self.group)
        super().__init__(data, args, predicate_align_model)

    # --- variables -------------------------------------------------------------------------------
    def _define_variables(self):
        a = self.args
        self._cap = max(2 * a.batch_size, a.attribute_batch_size, a.entity_batch_size) + 8
        kg1, kg2 = self.kgs.kg1, self.kgs.kg2
        self._split = len(kg1.entities_list)
        # KG-block placement: kg1 ids are [0, split), kg2 ids [split, entities_num) (base/kgs.py numbers them so)
        assert sorted(kg1.entities_list) == list(range(self._split)) and \
            sorted(kg2.entities_list) == list(range(self._split, self.kgs.entities_num))
        super()._define_variables()   # replicated: literal / name vectors, attr_embeds; entity tables via _entity_table
        self._name_staged = None if self.name_embeds is None else StagedConstant(self.name_embeds, self._cap)

    def _entity_table(self, name):
        sh = ShardedEmbeddingTable(self.kgs.entities_num, self.args.dim, True, self.group, init=self._init[name], name=name,
                                   split=self._split, flags=True)
        return StagedTable(sh, self._cap)

    def _define_relation_view_graph(self):
        kg1, kg2 = self.kgs.kg1, self.kgs.kg2
        t1, _ = _triples(kg1.local_relation_triples_list)
        t2, _ = _triples(kg2.local_relation_triples_list)
        f1, _ = _triples(list(kg1.local_relation_triples_set))
        f2, _ = _triples(list(kg2.local_relation_triples_set))
        assert self.args.batch_size % self.world == 0, "the global batch is split evenly over the ranks"
        # the GLOBAL batch is args.batch_size: G ranks train what one GPU trains per step
        self._rv = ShardedRelationView(
            self.kgs.entities_num, self.kgs.relations_num, self.args.dim, t1, t2, ent_split=self._split,
            batch_size=self.args.batch_size // self.world, neg_num=self.args.neg_triple_num, lr=self.args.learning_rate,
            seed=self.seed, group=self.group, ent_init=self._init["rv_ent_embeds"], rel_init=self._init["rel_embeds"],
            filter1=f1, filter2=f2)
        self.rv_ent_embeds = StagedTable(self._rv.ent, self._cap)
        self.rel_embeds = self._rv.rel

    # --- steps -----------------------------------------------------------------------------------
    def _attr_step(self, cnn, slot, cols, acc, weighted, scale):
        ih, ia, iv, w = cols
        ih = ih.contiguous()
        if not weighted:
            w = None
        staged, loc = self.av_ent_embeds.stage(ih)
        self._barrier.wait()   # every rank holds its copy of the batch's rows: updates may begin
        cnn.fwd_bwd(staged, self.attr_embeds, self.literal_embeds, loc, ia, iv, acc, w=w, scale=scale)
        lr = self.args.learning_rate
        self.av_ent_embeds.commit(ih, slot, lr)
        self.attr_embeds.apply_adagrad(slot, lr)
        cnn.apply_adagrad(slot, lr)
        self._barrier.wait()   # every shard is updated: the next step may stage

    def _positives_only_step(self, pos, w, acc, slot):
        rv = self._rv
        m = pos.shape[0]
        ids = torch.cat([pos[:, 0], pos[:, 2]]).contiguous()
        staged, loc = self.rv_ent_embeds.stage(ids)
        local = torch.stack([loc[:m], pos[:, 1], loc[m:2 * m]], 1).contiguous()
        self._barrier.wait()
        T.rel_step_structured(staged, rv.rel, local, None, None, 0, acc, w=w, pos_scale=2.0, variant=0)
        self.rv_ent_embeds.commit(ids, slot, rv.lr)
        T.apply_adagrad(rv.rel, rv.rel.adagrad_slot(slot), rv.lr)
        self._barrier.wait()

    def _align_step(self, pick, acc, lr, cvw):
        tabs = (self.ent_embeds, self.rv_ent_embeds, self.av_ent_embeds)
        staged = [t.stage(pick) for t in tabs]
        name = self._name_staged.stage(pick)
        loc = staged[0][1]
        self._barrier.wait()
        T.align_fwd_bwd(staged[0][0], name, staged[1][0], staged[2][0], loc, acc, name_weight=self.args.cv_name_weight,
                        scale=cvw)
        for t in tabs:
            t.commit(pick, self._cn_slot, lr)
        self._barrier.wait()

    def _space_step(self, idx, ws, total, lr, ow):
        lib = _cabi.load()
        F, loc = self.ent_embeds.stage(idx)
        rv, _ = self.rv_ent_embeds.stage(idx)
        av, _ = self.av_ent_embeds.stage(idx)
        name = self._name_staged.stage(idx)
        self._barrier.wait()
        _cabi.check(lib.mke_space_mapping_fwd_bwd(
            F.c, name.c, rv.c, av.c, loc.data_ptr(), idx.numel(), self._maps.data_ptr(), self._maps_grad.data_ptr(),
            float(ow), 0.0001, ws.data_ptr(), total.data_ptr(), _cabi.current_stream()))
        self.ent_embeds.commit(idx, self._sm_slot, lr)
        _cabi.check(lib.mke_dense_apply_adagrad(self._maps.data_ptr(), self._maps_grad.data_ptr(),
                                                self._maps_acc.data_ptr(), self._maps.numel(), float(lr),
                                                _cabi.current_stream()))
        self._barrier.wait()

    # --- epochs ----------------------------------------------------------------------------------
    def train_relation_view_1epo(self, epoch, triple_steps, steps_tasks, batch_queue, neighbors1, neighbors2):
        start = time.time()
        rv = self._rv
        rv.set_neighbours(_neighbour_matrix(neighbors1, rv.ent.rows), _neighbour_matrix(neighbors2, rv.ent.rows))
        epoch_loss, _ = rv.train_epoch()
        # random.shuffle of both lists (:314-315) as the single-GPU model does it: the same device permutation on
        # every rank (every rank holds the whole lists and the generators are in step)
        rv.triples1.copy_(rv.triples1[torch.randperm(rv.n1, device=rv.device)])
        rv.triples2.copy_(rv.triples2[torch.randperm(rv.n2, device=rv.device)])
        print('epoch {} of rel. view, avg. loss: {:.4f}, time: {:.4f}s'.format(epoch, epoch_loss, time.time() - start))
        return epoch_loss

    def _end_of_epoch_sync(self):
        """replicated dense parameters: rank 0's copy everywhere (identical up to float-atomic rounding before)"""
        import torch.distributed as dist
        tensors = [self.rel_embeds.var, self.attr_embeds.var]
        for tab in (self.rel_embeds, self.attr_embeds):
            tensors += list(tab._slots.values())
        for cnn in getattr(self, "_cnns", []):
            tensors += [cnn.theta] + list(cnn._slots.values())
        if getattr(self, "_maps", None) is not None:
            tensors += [self._maps, self._maps_acc]
        src = dist.get_global_rank(self.group, 0) if self.group is not dist.group.WORLD else 0
        on_host = dist.get_backend(self.group) == "gloo"   # (the shared-GPU test mode)
        for t in tensors:
            if on_host:
                h = t.cpu()
                dist.broadcast(h, src=src, group=self.group)
                t.copy_(h)
            else:
                dist.broadcast(t, src=src, group=self.group)

    def close(self):
        torch.cuda.synchronize()
        self._rv.close()
        for t in (self.av_ent_embeds, self.ent_embeds):
            t.sh.close()
        self._barrier.close()


class ShardedMultiKE_CV(_ShardedModel, D.MultiKE_CV):
    pass


class ShardedMultiKE_Late(_ShardedModel, D.MultiKE_Late):
    pass
