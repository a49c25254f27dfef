"""Epoch drivers on top of refapi.MultiKE_model.MultiKE: what `run_ITC.py` / `run_SSL.py` call.

  MultiKE_CV.run     cross-view training, ITC        code/MultiKE_CSL.py:13-112
  MultiKE_Late.run   late combination, SSL mapping   code/MultiKE_Late.py:184-300
  valid / test       MultiKE_Late.py:14-61           (ranking by the fused similarity kernel)
  valid_WVA/test_WVA MultiKE_Late.py:62-181          (weighted view averaging)

Both run() loops are the same schedule with different tails, so they share one plan object
(`_Schedule`) instead of two copies of the loop.  Everything a step needs already lives on the
device: the `steps_tasks` / `batch_queue` arguments of the trainers are passed as None, the
neighbour lists are a device table (refapi.base.batch.generate_neighbours), and evaluation gathers
and ranks rows on the device (no .eval() round trip through numpy).
"""
import math
import time

import torch

from multike_b200.refapi.MultiKE_model import MultiKE
from multike_b200.refapi.base import batch as bat
from multike_b200.refapi.base import evaluation as eva

_VIEW_TABLES = {"nv": "name_embeds", "rv": "rv_ent_embeds", "av": "av_ent_embeds", "final": "ent_embeds"}


def view_rows(model, embed_choice, entities, w=(1, 1, 1)):
    """Rows `entities` of the chosen view as a device [n, dim] tensor (what the reference gets
    from table.eval(session)[entities], MultiKE_Late.py:15-30); 'avg' is the w-weighted sum of the
    name, relation and attribute views."""
    if embed_choice == "avg":
        parts = [getattr(model, _VIEW_TABLES[c]).export(entities) for c in ("nv", "rv", "av")]
        return w[0] * parts[0] + w[1] * parts[1] + w[2] * parts[2]
    return getattr(model, _VIEW_TABLES.get(embed_choice, "ent_embeds")).export(entities)


def _rank(model, embeds1, embeds2):
    hits1_12, mrr_12 = eva.valid(embeds1, embeds2, None, model.args.top_k, model.args.test_threads_num, normalize=True)
    return mrr_12


def valid(model, embed_choice='avg', w=(1, 1, 1)):
    """validation links against the valid + test candidates of KG2 (MultiKE_Late.py:14-36)"""
    print(embed_choice, 'valid results:')
    kgs = model.kgs
    return _rank(model, view_rows(model, embed_choice, kgs.valid_entities1, w),
                 view_rows(model, embed_choice, list(kgs.valid_entities2) + list(kgs.test_entities2), w))


def test(model, embed_choice='avg', w=(1, 1, 1)):
    """test links (MultiKE_Late.py:39-61; the reference also ranks these through eva.valid)"""
    print(embed_choice, 'test results:')
    kgs = model.kgs
    return _rank(model, view_rows(model, embed_choice, kgs.test_entities1, w),
                 view_rows(model, embed_choice, kgs.test_entities2, w))


def _unit(x):
    n = x.norm(dim=1, keepdim=True)
    return x / torch.where(n == 0, torch.ones_like(n), n)


def _agreement(view, others):
    """mean cosine between a view's rows and the mean of all three views (_compute_weight,
    MultiKE_Late.py:62-80): diag(normalize(v) normalize(mean)^T) is a row-wise dot product"""
    mean = (view + others[0] + others[1]) / 3
    wts = (_unit(view) * _unit(mean)).sum(1)
    print(tuple(wts.shape), float(wts.mean()))
    return float(wts.mean())


def wva(embeds1, embeds2, embeds3):
    return (_agreement(embeds1, (embeds2, embeds3)), _agreement(embeds2, (embeds1, embeds3)),
            _agreement(embeds3, (embeds1, embeds2)))


def _wva_rank(model, ents1, ents2, label):
    views1 = [getattr(model, _VIEW_TABLES[c]).export(ents1) for c in ("nv", "rv", "av")]
    views2 = [getattr(model, _VIEW_TABLES[c]).export(ents2) for c in ("nv", "rv", "av")]
    w = [a + b for a, b in zip(wva(*views1), wva(*views2))]
    total = sum(w)
    w = [x / total for x in w]
    print('weights', *w)
    print(label)
    return _rank(model, sum(x * v for x, v in zip(w, views1)), sum(x * v for x, v in zip(w, views2)))


def valid_WVA(model):
    kgs = model.kgs
    return _wva_rank(model, kgs.valid_entities1, list(kgs.valid_entities2) + list(kgs.test_entities2), 'wvag valid results:')


def test_WVA(model):
    return _wva_rank(model, model.kgs.test_entities1, model.kgs.test_entities2, 'wvag test results:')


class _Schedule:
    """Step counts and cross-KG triple lists of one run (first lines of both run() methods)."""

    def __init__(self, m):
        kg1, kg2, a = m.kgs.kg1, m.kgs.kg2, m.args
        self.relation_steps = int(math.ceil((kg1.local_relation_triples_num + kg2.local_relation_triples_num) / a.batch_size))
        self.attribute_steps = int(math.ceil((kg1.local_attribute_triples_num + kg2.local_attribute_triples_num) / a.batch_size))
        self.ckge_relation = kg1.sup_relation_triples_list + kg2.sup_relation_triples_list
        self.ckge_attribute = kg1.sup_attribute_triples_list + kg2.sup_attribute_triples_list
        self.entity_list = kg1.entities_list + kg2.entities_list
        self.neighbors1 = self.neighbors2 = None
        self.refresh_predicates(m, update=False)

    def refresh_predicates(self, m, update=True):
        """soft predicate alignment from the current rel/attr embeddings (MultiKE_CSL.py:80-87)"""
        pam = m.predicate_align_model
        if update:
            pam.update_predicate_alignment(m.rel_embeds.eval(session=m.session))
            pam.update_predicate_alignment(m.attr_embeds.eval(session=m.session), predicate_type='attribute')
        self.ckgp_relation = pam.sup_relation_alignment_triples1 + pam.sup_relation_alignment_triples2
        self.ckgp_attribute = pam.sup_attribute_alignment_triples1 + pam.sup_attribute_alignment_triples2

    def train_views(self, m, i):
        """one epoch of the two views with their cross-KG inference steps"""
        soft = i > m.args.start_predicate_soft_alignment
        m.train_relation_view_1epo(i, self.relation_steps, None, None, self.neighbors1, self.neighbors2)
        m.train_cross_kg_entity_inference_relation_view_1epo(i, self.ckge_relation)
        if soft:
            m.train_cross_kg_relation_inference_1epo(i, self.ckgp_relation)
        m.train_attribute_view_1epo(i, self.attribute_steps, None, None, self.neighbors1, self.neighbors2)
        m.train_cross_kg_entity_inference_attribute_view_1epo(i, self.ckge_attribute)
        if soft:
            m.train_cross_kg_attribute_inference_1epo(i, self.ckgp_attribute)
        m._end_of_epoch_sync()

    def refresh_neighbours(self, m, i):
        """truncated-epsilon candidates every truncated_freq epochs (MultiKE_CSL.py:89-103)"""
        a = m.args
        if a.neg_sampling != 'truncated' or i % a.truncated_freq != 0:
            return
        t1 = time.time()
        assert 0.0 < a.truncated_epsilon < 1.0
        rows = m.rv_ent_embeds.rows
        lists = []
        for kg, useful in ((m.kgs.kg1, m.kgs.useful_entities_list1), (m.kgs.kg2, m.kgs.useful_entities_list2)):
            k = int((1 - a.truncated_epsilon) * kg.entities_num)
            lists.append(bat.generate_neighbours(m.rv_ent_embeds.export(useful), useful, k, a.batch_threads_num,
                                                 table_rows=rows))
        self.neighbors1, self.neighbors2 = lists
        print('neighbor dict:', len(self.neighbors1), type(self.neighbors2))
        print("generating neighbors of {} entities costs {:.3f} s.".format(len(self.entity_list), time.time() - t1))


class _Driver(MultiKE):
    def __init__(self, data, args, predicate_align_model):
        super().__init__(data, args, predicate_align_model)
        self.flag1 = self.flag2 = -1
        self.early_stop = False
        self._define_variables()
        for view in ("name", "relation", "attribute"):
            getattr(self, "_define_%s_view_graph" % view)()
        self._define_cross_kg_entity_reference_relation_view_graph()
        self._define_cross_kg_entity_reference_attribute_view_graph()
        self._define_cross_kg_relation_reference_graph()
        self._define_cross_kg_attribute_reference_graph()
        self._define_common_space_learning_graph()

    def _due(self, i):
        return i >= self.args.start_valid and i % self.args.eval_freq == 0


class MultiKE_CV(_Driver):
    """ITC: every epoch ends with common-space learning over all entities (MultiKE_CSL.py:36-108)."""

    def run(self):
        t = time.time()
        plan = _Schedule(self)
        test(self, embed_choice='nv')
        for i in range(1, self.args.max_epoch + 1):
            print('epoch {}:'.format(i))
            plan.train_views(self, i)
            self.train_common_space_learning_1epo(i, plan.entity_list)
            if self._due(i):
                for choice in ('rv', 'av', 'final'):
                    valid(self, embed_choice=choice)
                if self.early_stop or i == self.args.max_epoch:
                    break
            if i >= self.args.start_predicate_soft_alignment and i % 10 == 0:
                plan.refresh_predicates(self)
            plan.refresh_neighbours(self, i)
        self.save()
        for choice in ('nv', 'rv', 'av', 'final'):
            test(self, embed_choice=choice)
        print("Training ends. Total time = {:.3f} s.".format(time.time() - t))


class MultiKE_Late(_Driver):
    """SSL: views trained separately, then the shared space is learned by orthogonal mappings
    (MultiKE_Late.py:184-290)."""

    def __init__(self, data, args, attr_align_model):
        super().__init__(data, args, attr_align_model)
        self._define_space_mapping_graph()

    def run(self):
        t = time.time()
        plan = _Schedule(self)
        valid(self, embed_choice='nv')
        valid(self, embed_choice='avg')
        for i in range(1, self.args.max_epoch + 1):
            print('epoch {}:'.format(i))
            plan.train_views(self, i)
            if self._due(i):
                for choice in ('rv', 'av', 'avg'):
                    valid(self, embed_choice=choice)
                valid_WVA(self)
                if i >= self.args.start_predicate_soft_alignment:
                    plan.refresh_predicates(self)
            if self.early_stop or i == self.args.max_epoch:
                break
            plan.refresh_neighbours(self, i)
        for i in range(1, self.args.shared_learning_max_epoch + 1):
            self.train_shared_space_mapping_1epo(i, plan.entity_list)
            self._end_of_epoch_sync()
            if self._due(i):
                valid(self, embed_choice='final')
        self.save()
        for choice in ('nv', 'rv', 'av', 'avg'):
            test(self, embed_choice=choice)
        test_WVA(self)
        test(self, embed_choice='final')
        print("Training ends. Total time = {:.3f} s.".format(time.time() - t))
