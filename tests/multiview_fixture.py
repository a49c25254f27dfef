"""A small synthetic two-KG, three-view dataset shaped like what DataModel / PredicateAlignModel
hand to the drivers (attribute names follow code/base/kgs.py, code/base/kg.py, code/data_model.py,
code/predicate_alignment.py).  KG2 is a re-labelled noisy copy of KG1, so alignment is learnable."""
import types

import numpy as np


class PredicateAlignStub:
    """the fields and the one method the drivers touch (predicate_alignment.py:60-126)"""

    def __init__(self, attr1, attr2, sup_rel1, sup_rel2, sup_attr1, sup_attr2):
        self.attribute_triples_w_weights1, self.attribute_triples_w_weights2 = attr1, attr2
        self.attribute_triples_w_weights_set1, self.attribute_triples_w_weights_set2 = set(attr1), set(attr2)
        self.sup_relation_alignment_triples1, self.sup_relation_alignment_triples2 = sup_rel1, sup_rel2
        self.sup_attribute_alignment_triples1, self.sup_attribute_alignment_triples2 = sup_attr1, sup_attr2
        self.updates = []

    def update_predicate_alignment(self, embeds, predicate_type='relation', w=0.7):
        self.updates.append((predicate_type, np.asarray(embeds).shape))


def make(n=400, n_rel=6, n_attr=5, n_val=60, dim=75, seed=0, train_frac=0.3, valid_frac=0.1):
    rng = np.random.default_rng(seed)
    ents1, ents2 = list(range(n)), list(range(n, 2 * n))
    perm = rng.permutation(n)                      # entity i of KG1 <-> n + perm[i] of KG2
    t1 = sorted({(int(rng.integers(n)), int(rng.integers(n_rel)), int(rng.integers(n))) for _ in range(8 * n)})
    keep = rng.random(len(t1)) < 0.85
    t2 = sorted({(n + int(perm[h]), n_rel + r, n + int(perm[t])) for (h, r, t), k in zip(t1, keep) if k})
    links = [(i, n + int(perm[i])) for i in range(n)]
    order = rng.permutation(n)
    n_tr, n_va = int(train_frac * n), int(valid_frac * n)
    train = [links[i] for i in order[:n_tr]]
    valid = [links[i] for i in order[n_tr:n_tr + n_va]]
    test = [links[i] for i in order[n_tr + n_va:]]
    to2 = dict(train)
    to1 = {b: a for a, b in train}
    # swapping (base/read.py:130-145): training links generate triples with the counterpart entity
    sup1 = sorted({(to2.get(h, h), r, to2.get(t, t)) for h, r, t in t1 if h in to2 or t in to2})
    sup2 = sorted({(to1.get(h, h), r, to1.get(t, t)) for h, r, t in t2 if h in to1 or t in to1})
    # attribute triples (entity, attribute, literal id, weight)
    val_of = rng.integers(0, n_val, (n, 3))
    a1 = [(i, int(a), int(val_of[i, a % 3]), 1.0) for i in range(n) for a in rng.choice(n_attr, 2, replace=False)]
    a2 = [(n + int(perm[i]), n_attr + int(a), int(v), float(rng.choice([1.0, 0.9])))
          for (i, a, v, _) in a1 if rng.random() < 0.8]
    sa1 = [(to2[h], a, v) for h, a, v, _ in a1 if h in to2]
    sa2 = [(to1[h], a, v) for h, a, v, _ in a2 if h in to1]

    def kg(trip, sup, attr, sup_attr, ents):
        k = types.SimpleNamespace()
        k.entities_list, k.entities_num = list(ents), len(ents)
        k.local_relation_triples_list, k.local_relation_triples_num = list(trip), len(trip)
        k.local_relation_triples_set = set(trip) | set(sup)
        k.sup_relation_triples_list = list(sup)
        k.local_attribute_triples_num = len(attr)
        k.sup_attribute_triples_list = list(sup_attr)
        return k

    kgs = types.SimpleNamespace(
        kg1=kg(t1, sup1, a1, sa1, ents1), kg2=kg(t2, sup2, a2, sa2, ents2),
        entities_num=2 * n, relations_num=2 * n_rel, attributes_num=2 * n_attr,
        useful_entities_list1=ents1, useful_entities_list2=ents2,
        train_links=train, valid_links=valid, test_links=test,
        valid_entities1=[a for a, _ in valid], valid_entities2=[b for _, b in valid],
        test_entities1=[a for a, _ in test], test_entities2=[b for _, b in test])
    # name view: 60 % of the aligned pairs share their name vector exactly, the rest are noisy copies
    names = rng.standard_normal((2 * n, dim)).astype(np.float32)
    for i in range(n):
        j = n + int(perm[i])
        names[j] = names[i] if rng.random() < 0.6 else names[i] + 3.0 * rng.standard_normal(dim).astype(np.float32)
    names /= np.linalg.norm(names, axis=1, keepdims=True)
    values = rng.standard_normal((n_val, dim)).astype(np.float32)
    values /= np.linalg.norm(values, axis=1, keepdims=True)
    data = types.SimpleNamespace(kgs=kgs, local_name_vectors=names, value_vectors=values)
    w = [(h, r, t, 0.9) for h, r, t in sup1[:50]]
    pam = PredicateAlignStub(a1, a2, w, [(h, r, t, 0.8) for h, r, t in sup2[:50]],
                             [(h, a, v, 0.9) for h, a, v in sa1[:40]], [(h, a, v, 0.7) for h, a, v in sa2[:40]])
    args = types.SimpleNamespace(
        alignment_module='swapping', output='/tmp/mke_out/', training_data='x/SYN_toy/', dim=dim, seed=seed,
        learning_rate=0.01, ITC_learning_rate=0.04, batch_size=500, entity_batch_size=200, attribute_batch_size=400,
        neg_triple_num=10, neg_sampling='truncated', truncated_epsilon=0.9, truncated_freq=2, batch_threads_num=4,
        test_threads_num=8, max_epoch=4, shared_learning_max_epoch=3, start_valid=2, eval_freq=2, top_k=[1, 5, 10, 50],
        orthogonal_weight=2, cv_name_weight=1, cv_weight=1, start_predicate_soft_alignment=1, is_save=False)
    return data, args, pam
