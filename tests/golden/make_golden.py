"""Generates the golden fixtures under tests/golden/.  Run ONCE in the build container
(`python tests/golden/make_golden.py`); the GPU box never sees /root/reference.

  ref_batch_relation.npz / ref_batch_attribute.npz
      outputs of the REFERENCE's own code/base/batch.py and code/attr_batch.py (imported
      unmodified from /root/reference/code with stub `tensorflow`/`gensim` modules) under fixed
      `random` / `numpy.random` seeds.  They pin oracle/ref_batch.py.
  relation_step_d75.npz / relation_step_d128.npz
      known-answer vectors of the relation-view step computed by the float64 dense oracle
      (oracle/relation_view.py): init tables, adversarial index batches, loss, dense gradients and
      tables after 1 and 3 Adagrad steps.  TF itself cannot run here ("parity unpinned" for the
      TF arithmetic, see oracle/__init__.py); these vectors freeze the oracle's answers.
"""
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/code"


def synthetic_kg(rng, ent_lo, n_ent, rel_lo, n_rel, n_triples):
    trip = set()
    while len(trip) < n_triples:
        h = ent_lo + int(rng.integers(n_ent))
        t = ent_lo + int(rng.integers(n_ent))
        r = rel_lo + int(rng.integers(n_rel))
        trip.add((h, r, t))
    return sorted(trip)


def import_reference():
    tf = types.ModuleType("tensorflow")
    g = types.ModuleType("gensim")
    gm = types.ModuleType("gensim.models")
    gw = types.ModuleType("gensim.models.word2vec")
    gw.Word2Vec = object
    sys.modules.update({"tensorflow": tf, "gensim": g, "gensim.models": gm, "gensim.models.word2vec": gw})
    sys.path.insert(0, REF)
    import attr_batch
    import base.batch as bat
    return bat, attr_batch


def make_ref_batch():
    bat, attr_batch = import_reference()
    rng = np.random.default_rng(7)
    # small, dense KGs so that filtering, retries and duplicate draws really happen
    n_ent = 40
    t1 = synthetic_kg(rng, 0, n_ent, 0, 3, 700)
    t2 = synthetic_kg(rng, n_ent, n_ent, 3, 2, 500)
    ents1, ents2 = list(range(0, n_ent)), list(range(n_ent, 2 * n_ent))
    # the filter set also holds "sup" triples that are not in the list (base/kg.py:59,134)
    sup1 = synthetic_kg(rng, 0, n_ent, 0, 3, 150)
    sup2 = synthetic_kg(rng, n_ent, n_ent, 3, 2, 150)
    set1, set2 = set(t1) | set(sup1), set(t2) | set(sup2)
    nb1 = {e: [int(x) for x in rng.choice(ents1, 12, replace=False)] for e in ents1[::2]}
    nb2 = {e: [int(x) for x in rng.choice(ents2, 12, replace=False)] for e in ents2[::3]}
    out = {"triples1": np.array(t1), "triples2": np.array(t2), "sup1": np.array(sup1), "sup2": np.array(sup2),
           "n_ent": n_ent,
           "nb1_keys": np.array(sorted(nb1)), "nb1_vals": np.array([nb1[k] for k in sorted(nb1)]),
           "nb2_keys": np.array(sorted(nb2)), "nb2_vals": np.array([nb2[k] for k in sorted(nb2)])}
    B, K = 100, 10
    cases = []
    for case, (use_nb, seed, step) in enumerate([(False, 11, 0), (False, 12, 3), (True, 13, 1), (False, 14, 11),
                                                 (True, 15, 12)]):
        random.seed(seed)
        np.random.seed(seed)
        pos, neg = bat.generate_relation_triple_batch(t1, t2, set1, set2, ents1, ents2, B, step,
                                                      nb1 if use_nb else None, nb2 if use_nb else None, K)
        out["case%d_pos" % case] = np.array(pos, dtype=np.int64).reshape(-1, 3)
        out["case%d_neg" % case] = np.array(neg, dtype=np.int64).reshape(-1, 3)
        cases.append((int(use_nb), seed, step))
    out["cases"] = np.array(cases)
    out["B"], out["K"] = B, K
    np.savez_compressed(os.path.join(HERE, "ref_batch_relation.npz"), **out)

    # attribute batcher (attr_batch.py): tuples carry a weight
    a1 = [(h, a, v, float(w)) for (h, a, v), w in zip(synthetic_kg(rng, 0, n_ent, 0, 4, 300),
                                                       rng.choice([1.0, 0.9, 0.85], 300))]
    a2 = [(h, a, v, float(w)) for (h, a, v), w in zip(synthetic_kg(rng, n_ent, n_ent, 4, 3, 260),
                                                       rng.choice([1.0, 0.95], 260))]
    aout = {"a1": np.array(a1), "a2": np.array(a2)}
    acases = []
    for case, (K2, seed, step) in enumerate([(0, 21, 0), (2, 22, 1), (3, 23, 4)]):
        random.seed(seed)
        np.random.seed(seed)
        pos, neg = attr_batch.generate_attribute_triple_batch(a1, a2, set(a1), set(a2), ents1, ents2, 64, step,
                                                              None, None, K2)
        aout["case%d_pos" % case] = np.array(pos, dtype=np.float64).reshape(-1, 4)
        aout["case%d_neg" % case] = np.array(neg, dtype=np.float64).reshape(-1, 4)
        acases.append((K2, seed, step))
    aout["cases"] = np.array(acases)
    np.savez_compressed(os.path.join(HERE, "ref_batch_attribute.npz"), **aout)


def make_relation_step(dim, K, fname):
    import torch

    from oracle import relation_view as orv
    from oracle.tf_semantics import xavier_truncated_normal
    gen = torch.Generator().manual_seed(20190754 + dim)
    n_ent, n_rel, B = 200, 7, 48
    ent0 = xavier_truncated_normal((n_ent, dim), gen, torch.float64)
    rel0 = xavier_truncated_normal((n_rel, dim), gen, torch.float64)
    ent0[5] = 1e-9 * ent0[5] / ent0[5].norm()  # |v|^2 = 1e-18 < 1e-12: hits the l2_normalize clamp
    rng = np.random.default_rng(dim)
    pos = np.stack([rng.integers(0, n_ent, B), rng.integers(0, n_rel, B), rng.integers(0, n_ent, B)], 1)
    pos[0] = (3, 1, 3)            # self loop h == t
    pos[1:9, 1] = 2               # one hot relation
    pos[9:17, 0] = 11             # duplicate-heavy head
    pos[17] = (5, 0, 6)           # clamped row as head
    neg_ent = rng.integers(0, n_ent, (B, K)).astype(np.int32)
    side = rng.integers(0, 2, B).astype(np.uint32) * np.uint32((1 << K) - 1)
    side[2] = np.uint32(0b0101010101 & ((1 << K) - 1))   # mixed sides inside one positive
    neg_ent[3, 0] = pos[3, 2]     # a negative identical to its positive (legal: batch.py:103-105)
    side[3] &= ~np.uint32(1)
    neg_ent[4, :] = 5             # clamped row as corrupted entity, repeated
    neg = orv.structured_to_negatives(pos, neg_ent, side, K)
    out = {"ent0": ent0.numpy(), "rel0": rel0.numpy(), "pos": pos, "neg_ent": neg_ent, "neg_side": side, "neg": neg,
           "K": K, "lr": 0.001}
    ent, rel = orv.DenseTable(ent0, True, torch.float64), orv.DenseTable(rel0, True, torch.float64)
    _, gE, gR = orv.view_gradients(ent, rel, pos[:, 0], pos[:, 1], pos[:, 2], neg[:, 0], neg[:, 1], neg[:, 2])
    out["view_grad_ent"], out["view_grad_rel"] = gE.numpy(), gR.numpy()
    out["pos_score"] = orv.triple_scores(ent, rel, ent, pos[:, 0], pos[:, 1], pos[:, 2]).numpy()
    out["neg_score"] = orv.triple_scores(ent, rel, ent, neg[:, 0], neg[:, 1], neg[:, 2]).numpy()
    for step in range(3):
        loss, ge, gr = orv.relation_view_step(ent, rel, pos[:, 0], pos[:, 1], pos[:, 2], neg[:, 0], neg[:, 1],
                                              neg[:, 2], 0.001)
        if step == 0:
            out["loss"], out["grad_ent"], out["grad_rel"] = loss, ge.numpy(), gr.numpy()
            out["ent1"], out["rel1"] = ent.var.numpy().copy(), rel.var.numpy().copy()
    out["ent3"], out["rel3"] = ent.var.numpy().copy(), rel.var.numpy().copy()
    out["loss3"] = loss
    # weighted positives-only variant (ckgp graph, MultiKE_model.py:187-201): scale 2, weights
    w = rng.choice([1.0, 0.9, 0.5], B)
    ent, rel = orv.DenseTable(ent0, True, torch.float64), orv.DenseTable(rel0, True, torch.float64)
    e = np.zeros(0, dtype=np.int64)
    loss, ge, gr = orv.relation_view_step(ent, rel, pos[:, 0], pos[:, 1], pos[:, 2], e, e, e, 0.001, pos_w=w,
                                          pos_scale=2.0)
    out.update({"w": w, "wo_loss": loss, "wo_ent1": ent.var.numpy().copy(), "wo_rel1": rel.var.numpy().copy()})
    np.savez_compressed(os.path.join(HERE, fname), **out)


if __name__ == "__main__":
    make_ref_batch()
    make_relation_step(75, 10, "relation_step_d75.npz")
    make_relation_step(128, 25, "relation_step_d128.npz")
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
