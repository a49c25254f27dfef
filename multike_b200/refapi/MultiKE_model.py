"""Mirror of code/MultiKE_model.py for the relation view (SURVEY.md section 8 rows a-2 .. a-7, a-11).

Same class name, constructor and method signatures as the reference, so MultiKE_CSL.MultiKE_CV /
MultiKE_Late.MultiKE_Late style drivers can call it; the TF graph + session.run are replaced by
device tables and the kernels behind the C-ABI (multike_b200/relation_view.py).  What differs on
purpose: batches and negatives never leave the device (``steps_tasks`` / ``batch_queue`` are
accepted and ignored), and the per-epoch print lines are kept verbatim.

Also mirrored: the three attribute-view CNN graphs (:134-151, :172-185, :203-221) on the conv()
kernels of csrc/mke_cnn.cu, and ITC common-space learning (:225-239, :458-473) on the fused
alignment kernel, and SSL space mapping (:241-261, :439-454; its 75x75 products go through
cuBLAS, gradient rows and Adagrad through the kernels).
"""
import sys as _sys

if __name__ == "MultiKE_model":  # imported under the reference's top-level name (refapi first on sys.path):
    import multike_b200.refapi.MultiKE_model as _canonical  # one module object, whichever name imported it first
    _sys.modules[__name__] = _canonical
import math
import os
import time

import numpy as np
import torch

from multike_b200 import _cabi
from multike_b200 import tables as T
from multike_b200.attr_view import AttrCNN
from multike_b200.relation_view import RelationView, clipped_slice, split_batch


def generate_out_folder(out_folder, training_data_path, div_path, method_name):
    """utils.py:52-57"""
    path = training_data_path.strip('/').split('/')[-1]
    folder = out_folder + method_name + '/' + path + "/" + div_path + str(time.strftime("%Y%m%d%H%M%S")) + "/"
    print("results output folder:", folder)
    return folder


def write_id_dict(path, dic):
    """utils.dict2file (utils.py:60-67): one "key<TAB>id" line per entry; None writes nothing"""
    if dic is None:
        return
    with open(path, 'w', encoding='utf8') as f:
        for key, idx in dic.items():
            f.write(str(key) + '\t' + str(idx) + '\n')
    print(path, "saved.")


def _triples(lst, with_weight=False):
    a = np.asarray(lst, dtype=np.float64 if with_weight else np.int64)
    if a.size == 0:
        return np.zeros((0, 3), np.int32), np.zeros(0, np.float32)
    return np.ascontiguousarray(a[:, :3], dtype=np.int32), (a[:, 3].astype(np.float32) if with_weight else None)


def _neighbour_matrix(neighbors, rows):
    """dict entity -> candidate list (base/batch.py:119-150) as an int32 [rows, k] matrix; entities
    without an entry get -1 in column 0 (fall back to the whole KG, neighbor.get(e, entities_list))."""
    if neighbors is None or len(neighbors) == 0:
        return None
    if hasattr(neighbors, "matrix"):  # refapi.base.batch.NeighbourTable: already on the device
        return neighbors.matrix
    k = min(len(v) for v in neighbors.values())
    m = -np.ones((rows, k), dtype=np.int32)
    for e, cand in neighbors.items():
        m[e] = np.asarray(cand[:k], dtype=np.int32)
    return m


class MultiKE:

    def __check_args(self):
        assert self.args.alignment_module == 'swapping'  # for cross-KG inference

    def __init__(self, data, args, attr_align_model):
        self.predicate_align_model = attr_align_model
        self.args = args
        self.__check_args()
        self.data = data
        self.kgs = kgs = data.kgs
        self.kg1 = kgs.kg1
        self.kg2 = kgs.kg2
        self.out_folder = generate_out_folder(self.args.output, self.args.training_data, '', self.__class__.__name__)
        self.session = None  # kept for the drivers' `.eval(session=self.session)` idiom
        self.device = torch.device(getattr(args, "device", "cuda"))
        self.seed = int(getattr(args, "seed", 0))
        _cabi.load()  # no library, no model: there is no CPU path

    # --- variables (MultiKE_model.py:86-107) ---------------------------------------------------
    def _define_variables(self):
        n_ent, n_rel, n_attr, dim = self.kgs.entities_num, self.kgs.relations_num, self.kgs.attributes_num, self.args.dim
        gen = torch.Generator().manual_seed(self.seed)
        dev = self.device
        self._init = {name: T.xavier_truncated_normal(rows, dim, gen) for name, rows in
                      (("rv_ent_embeds", n_ent), ("rel_embeds", n_rel), ("av_ent_embeds", n_ent),
                       ("attr_embeds", max(n_attr, 1)), ("ent_embeds", n_ent))}
        value_vectors = getattr(self.data, "value_vectors", None)
        name_vectors = getattr(self.data, "local_name_vectors", None)
        self.literal_embeds = None if value_vectors is None else T.EmbeddingTable(
            len(value_vectors), dim, False, dev, init=np.asarray(value_vectors, np.float32), trainable=False)
        self.name_embeds = None if name_vectors is None else T.EmbeddingTable(
            len(name_vectors), dim, False, dev, init=np.asarray(name_vectors, np.float32), trainable=False)
        self.av_ent_embeds = self._entity_table("av_ent_embeds")
        # False important! (MultiKE_model.py:96-97)
        self.attr_embeds = T.EmbeddingTable(max(n_attr, 1), dim, False, dev, init=self._init["attr_embeds"])
        self.ent_embeds = self._entity_table("ent_embeds")
        self.rv_ent_embeds = None  # created with the relation-view graph (they own the triple lists)
        self.rel_embeds = None

    def _entity_table(self, name):
        """a trainable [entities, dim] table read through l2_normalize(., 1) (the multi-GPU model shards these)"""
        return T.EmbeddingTable(self.kgs.entities_num, self.args.dim, True, self.device, init=self._init[name], flags=True,
                                grad_replicas=1)

    # --- view-specific graphs ---------------------------------------------------------------
    def _define_name_view_graph(self):
        pass

    def _define_relation_view_graph(self):
        """MultiKE_model.py:114-132"""
        kg1, kg2 = self.kgs.kg1, self.kgs.kg2
        t1, _ = _triples(kg1.local_relation_triples_list)
        t2, _ = _triples(kg2.local_relation_triples_list)
        # the filter set aliases relation_triples_set and so also holds the swapped sup triples
        f1, _ = _triples(list(kg1.local_relation_triples_set))
        f2, _ = _triples(list(kg2.local_relation_triples_set))
        self._rv = RelationView(
            self.kgs.entities_num, self.kgs.relations_num, self.args.dim, t1, t2, ent_split=len(kg1.entities_list),
            batch_size=self.args.batch_size, neg_num=self.args.neg_triple_num, lr=self.args.learning_rate,
            seed=self.seed, device=self.device, ent_init=self._init["rv_ent_embeds"], rel_init=self._init["rel_embeds"],
            filter1=f1, filter2=f2, entities1=kg1.entities_list, entities2=kg2.entities_list)
        self.rv_ent_embeds, self.rel_embeds = self._rv.ent, self._rv.rel

    def _define_cross_kg_name_view_graph(self):
        pass

    def _define_cross_kg_entity_reference_relation_view_graph(self):
        """MultiKE_model.py:158-170: 2 * relation_logistic_loss_wo_negs, its own Adagrad slots"""
        self._ckge_slot = "ckge_relation"

    def _define_cross_kg_relation_reference_graph(self):
        """MultiKE_model.py:187-201: 2 * logistic_loss_wo_negs (weighted), its own Adagrad slots"""
        self._ckgp_slot = "ckgp_relation"

    def _define_common_space_learning_graph(self):
        """MultiKE_model.py:225-239: cv_weight * (cv_name_weight |F-N|^2 + |F-R|^2 + |F-A|^2), Adagrad
        with args.ITC_learning_rate and its own accumulator slots"""
        assert self.name_embeds is not None, "the common-space graph needs data.local_name_vectors"
        self._cn_slot = "cross_name"

    # --- attribute view: conv() score (MultiKE_model.py:34-63), three independent weight sets ------
    def _new_cnn(self):
        gen = torch.Generator().manual_seed(self.seed + 1000 + len(getattr(self, "_cnns", [])))
        cnn = AttrCNN(self.args.dim, self.device, generator=gen)
        self._cnns = getattr(self, "_cnns", []) + [cnn]
        return cnn

    def _define_attribute_view_graph(self):
        """MultiKE_model.py:134-151 (weighted, no negatives: neg_triples_num is the literal 0, :331)"""
        assert self.literal_embeds is not None, "the attribute view needs data.value_vectors"
        self._attr_cnn, self._attr_slot = self._new_cnn(), "attribute"

    def _define_cross_kg_entity_reference_attribute_view_graph(self):
        """MultiKE_model.py:172-185: 2 * sum log(1 + exp(-score)), own conv() weights"""
        self._ckge_attr_cnn, self._ckge_attr_slot = self._new_cnn(), "ckge_attribute"

    def _define_cross_kg_attribute_reference_graph(self):
        """MultiKE_model.py:203-221: weighted, own conv() weights"""
        self._ckga_attr_cnn, self._ckga_attr_slot = self._new_cnn(), "ckga_attribute"

    # --- device-resident triple lists ------------------------------------------------
    # The reference keeps every triple list as a Python list of tuples, re-reads it every epoch and shuffles it in
    # place (MultiKE_model.py:314-315, :343-344).  Here a list is converted ONCE into int32 / fp32 column vectors in HBM
    # (1.6 M attribute triples: 0.9 s of np.asarray per epoch otherwise, plus 1.3 s of random.shuffle); the copy is
    # found again by the list object's identity and a fingerprint of its content, and the epoch's shuffle is a device
    # permutation of the copy -- the Python list keeps its order, which nothing else reads.
    def _pick(self, n, count):
        """random.sample(range(n), count) on the device (mke_sample_distinct): the reference's per-step batch draw"""
        self._draws = getattr(self, "_draws", 0) + 1
        return T.sample_distinct(n, count, self.seed, self._draws, self.device)

    def _device_columns(self, lst):
        """(h, p, t int32 [n], w float32 [n]) of a list of (h, p, t[, w]) tuples, cached"""
        cache = self.__dict__.setdefault("_list_cache", {})
        n = len(lst)
        mark = (n, tuple(lst[0]), tuple(lst[n // 2]), tuple(lst[-1])) if n else (0,)
        hit = cache.get(id(lst))
        if hit is not None and hit[0] is lst and hit[1] == mark:
            return hit[2]
        a = np.asarray(lst, dtype=np.float64)
        if n == 0:
            a = np.zeros((0, 4))
        if a.shape[1] == 3:
            a = np.concatenate([a, np.ones((n, 1))], 1)
        cols = tuple(torch.from_numpy(np.ascontiguousarray(a[:, k], dtype=np.int32)).to(self.device) for k in range(3)) + \
            (torch.from_numpy(np.ascontiguousarray(a[:, 3], dtype=np.float32)).to(self.device),)
        if len(cache) > 16:
            cache.clear()
        cache[id(lst)] = (lst, mark, cols)
        return cols

    def _store_columns(self, lst, cols):
        """the epoch's shuffle: the device copy is replaced by its permutation"""
        hit = self._list_cache[id(lst)]
        self._list_cache[id(lst)] = (hit[0], hit[1], cols)   # (a cached [n, 3] copy, if any, is dropped with the old order)

    def _attr_step(self, cnn, slot, cols, acc, weighted, scale):
        """one session.run([loss, optimizer]) of an attribute graph on device columns (h, a, v int32, w fp32)"""
        ih, ia, iv, w = cols
        if not weighted:
            w = None
        cnn.fwd_bwd(self.av_ent_embeds, self.attr_embeds, self.literal_embeds, ih, ia, iv, acc, w=w, scale=scale)
        lr = self.args.learning_rate
        self.av_ent_embeds.apply_adagrad(slot, lr)
        self.attr_embeds.apply_adagrad(slot, lr)
        cnn.apply_adagrad(slot, lr)

    def train_attribute_view_1epo(self, epoch, triple_steps, steps_tasks, batch_queue, neighbors1, neighbors2):
        """MultiKE_model.py:319-345; batches as attr_batch.py:39-50 with neg_triples_num = 0"""
        start = time.time()
        pam = self.predicate_align_model
        l1, l2 = pam.attribute_triples_w_weights1, pam.attribute_triples_w_weights2
        c1, c2 = self._device_columns(l1), self._device_columns(l2)
        n1, n2 = len(l1), len(l2)
        b1, b2 = split_batch(n1, n2, self.args.attribute_batch_size)
        # the epoch's batches, attr_batch.py:39-50: step k = kg1 slice k ++ kg2 slice k; gathered ONCE into batch order,
        # so that a step's batch is a contiguous view
        bounds, order, pos = [], [], 0
        for step in range(triple_steps):
            (s1, e1), (s2, e2) = clipped_slice(n1, b1, step), clipped_slice(n2, b2, step)
            bounds.append((pos, pos + (e1 - s1) + (e2 - s2)))
            pos += (e1 - s1) + (e2 - s2)
            order += [(0, s1, e1), (1, s2, e2)]
        if not order:   # no attribute triples (or zero steps): nothing to train
            order = [(0, 0, 0)]
        epoch_cols = tuple(torch.cat([(c1, c2)[which][k][s:e] for which, s, e in order]) for k in range(4))
        acc = T.new_loss_accumulator(self.device)
        trained = 0
        for lo, hi in bounds:
            if hi == lo:
                continue
            self._attr_step(self._attr_cnn, self._attr_slot, tuple(c[lo:hi] for c in epoch_cols), acc, weighted=True, scale=1.0)
            trained += hi - lo
        epoch_loss = float(acc.item()) / max(trained, 1)
        # random.shuffle of both lists (:343-344), on the device copies
        p1, p2 = torch.randperm(n1, device=self.device), torch.randperm(n2, device=self.device)
        self._store_columns(l1, tuple(c[p1] for c in c1))
        self._store_columns(l2, tuple(c[p2] for c in c2))
        print('epoch {} of att. view, avg. loss: {:.4f}, time: {:.4f}s'.format(epoch, epoch_loss, time.time() - start))
        return epoch_loss

    def _attr_sampled_epoch(self, sup_triples, cnn, slot, weighted, scale):
        cols = self._device_columns(sup_triples)
        n = len(sup_triples)
        steps = int(math.ceil(n / self.args.attribute_batch_size))
        batch_size = self.args.attribute_batch_size if steps > 1 else n
        acc = T.new_loss_accumulator(self.device)
        for _ in range(steps):
            pick = self._pick(n, batch_size)  # random.sample
            self._attr_step(cnn, slot, tuple(c[pick] for c in cols), acc, weighted=weighted, scale=scale)
        return float(acc.item()) / max(steps * batch_size, 1)

    def train_cross_kg_entity_inference_attribute_view_1epo(self, epoch, sup_triples):
        """MultiKE_model.py:371-391"""
        if len(sup_triples) == 0:
            return
        start = time.time()
        epoch_loss = self._attr_sampled_epoch(sup_triples, self._ckge_attr_cnn, self._ckge_attr_slot, False, 2.0)
        print('epoch {} of cross-kg entity inference in attr. view, avg. loss: {:.4f}, time: {:.4f}s'.format(
            epoch, epoch_loss, time.time() - start))
        return epoch_loss

    def train_cross_kg_attribute_inference_1epo(self, epoch, sup_triples):
        """MultiKE_model.py:416-437"""
        if len(sup_triples) == 0:
            return
        start = time.time()
        epoch_loss = self._attr_sampled_epoch(sup_triples, self._ckga_attr_cnn, self._ckga_attr_slot, True, 1.0)
        print('epoch {} of cross-kg attribute inference in attr. view, avg. loss: {:.4f}, time: {:.4f}s'.format(
            epoch, epoch_loss, time.time() - start))
        return epoch_loss

    # --- SSL late combination (MultiKE_model.py:241-261, :439-454) --------------------------------
    def _define_space_mapping_graph(self):
        """Only variables whose name starts with "shared" train (:257): ent_embeds and the three
        dim x dim mappings (tf.initializers.orthogonal(), gain 1).  The whole step -- gathers, the three
        [batch, dim] x [dim, dim] products, the batch-wide norms, M M^T - I and every gradient -- is
        mke_space_mapping_fwd_bwd (csrc/mke_space.cu); both Adagrad updates use the apply kernels."""
        assert self.name_embeds is not None, "the space-mapping graph needs data.local_name_vectors"
        from multike_b200.refapi import losses as L
        self._sm_losses = L
        dim = self.args.dim
        gen = torch.Generator().manual_seed(self.seed + 77)
        mats = []
        for _ in range(3):  # orthogonal initializer: QR of a normal matrix, signs fixed by diag(R)
            q, r = torch.linalg.qr(torch.randn(dim, dim, generator=gen))
            mats.append(q * torch.sign(torch.diagonal(r)))
        self._maps = torch.stack(mats).to(self.device, torch.float32).contiguous()   # nv, rv, av
        self.nv_mapping, self.rv_mapping, self.av_mapping = self._maps[0], self._maps[1], self._maps[2]
        self._maps_grad = torch.zeros_like(self._maps)
        self._maps_acc = torch.full_like(self._maps, T.ADAGRAD_INIT)
        self.eye_mat = torch.eye(dim, device=self.device)
        self._sm_slot = "shared_comb"

    def _space_step(self, idx, ws, total, lr, ow):
        """one session.run of the space-mapping graph: forward + backward of the three mapping losses in three
        launches (csrc/mke_space.cu): gradient rows of the shared table and the three mapping gradients; only
        `shared*` variables train (:257)"""
        lib = _cabi.load()
        F_tab = self.ent_embeds
        _cabi.check(lib.mke_space_mapping_fwd_bwd(
            F_tab.c, self.name_embeds.c, self._rv.ent.c, self.av_ent_embeds.c, idx.data_ptr(), idx.numel(),
            self._maps.data_ptr(), self._maps_grad.data_ptr(), float(ow), 0.0001, ws.data_ptr(), total.data_ptr(),
            _cabi.current_stream()))
        F_tab.apply_adagrad(self._sm_slot, lr)
        _cabi.check(lib.mke_dense_apply_adagrad(self._maps.data_ptr(), self._maps_grad.data_ptr(),
                                                self._maps_acc.data_ptr(), self._maps.numel(), float(lr),
                                                _cabi.current_stream()))

    def train_shared_space_mapping_1epo(self, epoch, entities):
        start = time.time()
        lib = _cabi.load()
        ents = torch.as_tensor(np.asarray(entities, dtype=np.int32)).to(self.device)
        n = ents.numel()
        steps = int(math.ceil(n / self.args.entity_batch_size))
        batch_size = self.args.entity_batch_size if steps > 1 else n
        lr, ow = self.args.learning_rate, self.args.orthogonal_weight
        total = torch.zeros(1, dtype=torch.float64, device=self.device)
        dim = self.ent_embeds.dim
        ws = torch.empty(int(lib.mke_space_mapping_workspace_floats(batch_size, dim)), dtype=torch.float32,
                         device=self.device)
        for _ in range(steps):
            idx = ents[self._pick(n, batch_size)].contiguous()  # random.sample: distinct ids
            self._space_step(idx, ws, total, lr, ow)
        epoch_loss = float(total) / max(steps * batch_size, 1)
        print('epoch {} of shared space learning, avg. loss: {:.4f}, time: {:.4f}s'.format(epoch, epoch_loss,
                                                                                           time.time() - start))
        return epoch_loss

    def _align_step(self, pick, acc, lr, cvw):
        """one session.run of the common-space graph (:225-239)"""
        T.align_fwd_bwd(self.ent_embeds, self.name_embeds, self._rv.ent, self.av_ent_embeds, pick, acc,
                        name_weight=self.args.cv_name_weight, scale=cvw)
        for t in (self.ent_embeds, self._rv.ent, self.av_ent_embeds):
            t.apply_adagrad(self._cn_slot, lr)

    def train_common_space_learning_1epo(self, epoch, entities):
        """MultiKE_model.py:458-473"""
        start = time.time()
        ents = torch.as_tensor(np.asarray(entities, dtype=np.int32)).to(self.device)
        n = ents.numel()
        steps = int(math.ceil(n / self.args.entity_batch_size))
        batch_size = self.args.entity_batch_size if steps > 1 else n
        acc = T.new_loss_accumulator(self.device)
        lr, cvw = self.args.ITC_learning_rate, float(self.args.cv_weight)
        trained = 0
        for _ in range(steps):
            pick = ents[self._pick(n, batch_size)].contiguous()  # random.sample
            self._align_step(pick, acc, lr, cvw)
            trained += batch_size
        # the fetched cross_name_loss is the un-weighted sum (the optimizer minimises cv_weight * loss)
        epoch_loss = float(acc.item()) / (cvw if cvw != 0 else 1.0) / max(trained, 1)
        print('epoch {} of common space learning, avg. loss: {:.4f}, time: {:.4f}s'.format(epoch, epoch_loss,
                                                                                           time.time() - start))
        return epoch_loss

    # --- reads (MultiKE_model.py:263-287) ------------------------------------------------------
    def eval_kg1_ent_embeddings(self):
        return self.rv_ent_embeds.eval(idx=self.kgs.kg1.entities_list)

    def eval_kg2_ent_embeddings(self):
        return self.rv_ent_embeds.eval(idx=self.kgs.kg2.entities_list)

    def eval_kg1_useful_ent_embeddings(self):
        return self.rv_ent_embeds.eval(idx=self.kgs.useful_entities_list1)

    def eval_kg2_useful_ent_embeddings(self):
        return self.rv_ent_embeds.eval(idx=self.kgs.useful_entities_list2)

    def save(self):
        """six .npy files as utils.save_embeddings writes them (utils.py:70-91)"""
        folder = self.out_folder
        os.makedirs(folder, exist_ok=True)
        for fname, tab in (("ent_embeds", self.ent_embeds), ("nv_ent_embeds", self.name_embeds),
                           ("rv_ent_embeds", self.rv_ent_embeds), ("av_ent_embeds", self.av_ent_embeds),
                           ("rel_embeds", self.rel_embeds), ("attr_embeds", self.attr_embeds)):
            if tab is not None:
                np.save(folder + fname + '.npy', tab.eval())
        # the id dictionaries next to them: "<key>\t<id>" per line (utils.py:60-67, :84-89)
        for kg_name, kg in (("kg1", self.kgs.kg1), ("kg2", self.kgs.kg2)):
            for what, attr in (("ent", "entities_id_dict"), ("rel", "relations_id_dict"), ("attr", "attributes_id_dict")):
                write_id_dict(folder + kg_name + '_' + what + '_ids', getattr(kg, attr, None))
        print("Embeddings saved!")

    # --- training (MultiKE_model.py:291-317) ---------------------------------------------------
    def train_relation_view_1epo(self, epoch, triple_steps, steps_tasks, batch_queue, neighbors1, neighbors2):
        start = time.time()
        rv = self._rv
        rv.set_neighbours(_neighbour_matrix(neighbors1, rv.ent.rows), _neighbour_matrix(neighbors2, rv.ent.rows))
        trained_samples_num = rv.train_steps(0, triple_steps)
        epoch_loss = float(rv.step_losses.sum().item()) / max(trained_samples_num, 1)
        # random.shuffle of both lists (:314-315), on the device
        rv.triples1.copy_(rv.triples1[torch.randperm(rv.n1, device=rv.device)])
        rv.triples2.copy_(rv.triples2[torch.randperm(rv.n2, device=rv.device)])
        end = time.time()
        print('epoch {} of rel. view, avg. loss: {:.4f}, time: {:.4f}s'.format(epoch, epoch_loss, end - start))
        return epoch_loss

    def _positives_only_step(self, pos, w, acc, slot):
        rv = self._rv
        T.rel_step_structured(rv.ent, rv.rel, pos, None, None, 0, acc, w=w, pos_scale=2.0, variant=rv.variant)
        T.apply_adagrad_pair(rv.ent, rv.ent.adagrad_slot(slot), rv.lr, rv.rel, rv.rel.adagrad_slot(slot), rv.lr)

    def _end_of_epoch_sync(self):
        """hook of the multi-GPU model (multike_b200/sharded_model.py); nothing to do on one GPU"""

    def _positives_only_epoch(self, sup_triples, slot, weighted):
        """MultiKE_model.py:349-369 / 393-414: `steps` batches of random.sample(sup_triples, B),
        loss = 2 * [weighted] logistic loss without negatives, Adagrad slots of this graph."""
        rv = self._rv
        h, r, t, w_d = self._device_columns(sup_triples)
        entry = self._list_cache[id(sup_triples)]
        if len(entry) == 3:   # the [n, 3] rows next to the columns, same lifetime
            entry = self._list_cache[id(sup_triples)] = entry + (torch.stack([h, r, t], 1).contiguous(),)
        pos_d = entry[3]
        n = pos_d.shape[0]
        if not weighted:
            w_d = None
        steps = int(math.ceil(n / self.args.batch_size))
        batch_size = self.args.batch_size if steps > 1 else n
        acc = T.new_loss_accumulator(rv.device)
        trained = 0
        for _ in range(steps):
            pick = self._pick(n, batch_size)  # random.sample: without replacement
            self._positives_only_step(pos_d[pick].contiguous(), None if w_d is None else w_d[pick].contiguous(), acc, slot)
            trained += batch_size
        return float(acc.item()) / max(trained, 1)

    def train_cross_kg_entity_inference_relation_view_1epo(self, epoch, sup_triples):
        if len(sup_triples) == 0:
            return
        start = time.time()
        epoch_loss = self._positives_only_epoch(sup_triples, self._ckge_slot, weighted=False)
        print('epoch {} of cross-kg entity inference in rel. view, avg. loss: {:.4f}, time: {:.4f}s'.format(
            epoch, epoch_loss, time.time() - start))
        return epoch_loss

    def train_cross_kg_relation_inference_1epo(self, epoch, sup_triples):
        if len(sup_triples) == 0:
            return
        start = time.time()
        epoch_loss = self._positives_only_epoch(sup_triples, self._ckgp_slot, weighted=True)
        print('epoch {} of cross-kg relation inference in rel. view, avg. loss: {:.4f}, time: {:.4f}s'.format(
            epoch, epoch_loss, time.time() - start))
        return epoch_loss
