"""Restatement of the reference batcher / negative sampler, consuming Python's ``random`` and
``numpy.random`` in the same order as the reference so that, under equal seeds, outputs are
bit-identical to ``code/base/batch.py`` and ``code/attr_batch.py`` (pinned by
tests/golden/ref_batch_*.npz, generated from the reference itself by tests/golden/make_golden.py).
"""
import random

import numpy as np


def batch_sizes(n1, n2, batch_size):
    """base/batch.py:36-37: the kg1 share is floored, kg2 takes the rest."""
    b1 = int(n1 / (n1 + n2) * batch_size)
    return b1, batch_size - b1


def pos_slice(triples, batch_size, step):
    """base/batch.py:45-54 (is_fixed_size=False): contiguous slice clipped at the list end."""
    start = step * batch_size
    end = min(start + batch_size, len(triples))
    return triples[start:end]


def neg_triples_fast(pos_batch, all_triples_set, entities_list, neg_triples_num, neighbor=None, max_try=10):
    """base/batch.py:86-116.  Per positive: up to max_try rounds; each round flips ONE coin
    (head or tail side), draws the still-missing count without replacement from the candidate
    list of the replaced entity, drops candidates that are known triples (the last round keeps
    everything), and stops at neg_triples_num."""
    if neighbor is None:
        neighbor = {}
    out = []
    for head, relation, tail in pos_batch:
        got = []
        need = neg_triples_num
        head_cands = neighbor.get(head, entities_list)
        tail_cands = neighbor.get(tail, entities_list)
        for attempt in range(max_try):
            if np.random.binomial(1, 0.5):
                cand = {(h2, relation, tail) for h2 in random.sample(head_cands, need)}
            else:
                cand = {(head, relation, t2) for t2 in random.sample(tail_cands, need)}
            if attempt == max_try - 1:
                got += list(cand)
                break
            got += list(cand - all_triples_set)
            if len(got) == neg_triples_num:
                break
            need = neg_triples_num - len(got)
        assert len(got) == neg_triples_num
        out.extend(got)
    return out


def relation_triple_batch(triple_list1, triple_list2, triple_set1, triple_set2, entity_list1, entity_list2,
                          batch_size, step, neighbor1, neighbor2, neg_triples_num):
    """base/batch.py:33-42"""
    b1, b2 = batch_sizes(len(triple_list1), len(triple_list2), batch_size)
    pos1 = pos_slice(triple_list1, b1, step)
    pos2 = pos_slice(triple_list2, b2, step)
    neg1 = neg_triples_fast(pos1, triple_set1, entity_list1, neg_triples_num, neighbor=neighbor1)
    neg2 = neg_triples_fast(pos2, triple_set2, entity_list2, neg_triples_num, neighbor=neighbor2)
    return pos1 + pos2, neg1 + neg2


def neg_attribute_triples(pos_batch, all_triples_set, entity_list, neg_triples_num, neighbor=None):
    """attr_batch.py:13-25: head-only corruption, rejection until unseen, with replacement."""
    if neighbor is None:
        neighbor = {}
    out = []
    for head, attribute, value, w in pos_batch:
        for _ in range(neg_triples_num):
            while True:
                neg_head = random.choice(neighbor.get(head, entity_list))
                if (neg_head, attribute, value, w) not in all_triples_set:
                    break
            out.append((neg_head, attribute, value, w))
    return out


def attribute_triple_batch(triple_list1, triple_list2, triple_set1, triple_set2, entity_list1, entity_list2,
                           batch_size, step, neighbor1, neighbor2, neg_triples_num):
    """attr_batch.py:39-50"""
    b1, b2 = batch_sizes(len(triple_list1), len(triple_list2), batch_size)
    pos1 = pos_slice(triple_list1, b1, step)
    pos2 = pos_slice(triple_list2, b2, step)
    neg1 = neg_attribute_triples(pos1, triple_set1, entity_list1, neg_triples_num, neighbor=neighbor1)
    neg2 = neg_attribute_triples(pos2, triple_set2, entity_list2, neg_triples_num, neighbor=neighbor2)
    return pos1 + pos2, neg1 + neg2
