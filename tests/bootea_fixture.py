"""Writes a tiny two-KG dataset in the BootEA folder layout the reference's loader reads
(code/base/kgs.py:76-89, code/base/read.py:216-365, code/utils.py:94-137): rel_triples_{1,2},
attr_triples_{1,2}, entity_local_name_{1,2}, predicate_local_name_{1,2}, <division>/{train,valid,test}_links, plus a word-vector
text file in the .vec format of utils.read_word2vec and an args.json with the keys of code/args.json.
KG2 is a re-labelled noisy copy of KG1 (so alignment is learnable); every word of every literal is in
the word-vector file (so the character-embedding fallback, which needs gensim, is never asked)."""
import json
import os

import numpy as np

WORDS = ["alpha", "beta", "gamma", "delta", "river", "mount", "lake", "city", "north", "south", "east", "west",
         "red", "blue", "green", "old", "new", "great", "little", "saint", "port", "fort", "bridge", "field"]


def write(folder, n=240, n_rel=5, n_attr=4, seed=0, division="631/", word_dim=300):
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(folder, division), exist_ok=True)
    e1 = ["http://kg1.org/resource/E%d" % i for i in range(n)]
    e2 = ["http://kg2.org/entity/Q%d" % i for i in range(n)]
    perm = rng.permutation(n)  # E_i <-> Q_perm[i]
    r1 = ["http://kg1.org/ontology/rel%d" % r for r in range(n_rel)]
    r2 = ["http://kg2.org/prop/P%d" % r for r in range(n_rel)]
    a1 = ["http://kg1.org/ontology/attr%d" % a for a in range(n_attr)]
    a2 = ["http://kg2.org/prop/A%d" % a for a in range(n_attr)]
    t1 = set()
    for i in range(n):  # every entity occurs in a relation triple (entities come from them, base/kg.py:64)
        t1.add((i, int(rng.integers(n_rel)), int(rng.integers(n))))
    while len(t1) < 6 * n:
        t1.add((int(rng.integers(n)), int(rng.integers(n_rel)), int(rng.integers(n))))
    t1 = sorted(t1)
    t2 = [(int(perm[h]), r, int(perm[t])) for k, (h, r, t) in enumerate(t1) if k < n or rng.random() < 0.85]
    names = [" ".join(rng.choice(WORDS, size=2, replace=False)) for _ in range(n)]
    vals = [" ".join(rng.choice(WORDS, size=int(rng.integers(1, 4)))) for _ in range(3 * n)]

    def lines(path, rows):
        with open(os.path.join(folder, path), "w", encoding="utf8") as fh:
            for row in rows:
                fh.write("\t".join(row) + "\n")

    lines("rel_triples_1", [(e1[h], r1[r], e1[t]) for h, r, t in t1])
    lines("rel_triples_2", [(e2[h], r2[r], e2[t]) for h, r, t in t2])
    at1, at2 = [], []
    for i in range(n):
        for a in rng.choice(n_attr, size=2, replace=False):  # >= 10 triples per attribute survive the cleaning
            v = vals[int(rng.integers(len(vals)))]
            at1.append((e1[i], a1[int(a)], '"%s"@en' % v))
            if rng.random() < 0.8:
                at2.append((e2[int(perm[i])], a2[int(a)], v))
    lines("attr_triples_1", at1)
    lines("attr_triples_2", at2)
    # two thirds of the counterparts carry the same local name (cf. 64 934 of 100 000 in DBP-WD)
    lines("entity_local_name_1", [(e1[i], names[i].replace(" ", "_")) for i in range(n)])
    lines("entity_local_name_2", [(e2[int(perm[i])], (names[i] if i % 3 else names[(i + 1) % n]).replace(" ", "_"))
                                  for i in range(n)])
    # predicate_alignment.py:75-86, 138-141: "<predicate uri>\t<local name>" for relations and attributes alike;
    # counterpart predicates carry near-identical names, so the name-based initial alignment finds them
    lines("predicate_local_name_1", [(u, "relation%d" % k) for k, u in enumerate(r1)] +
          [(u, "attribute%d" % k) for k, u in enumerate(a1)])
    lines("predicate_local_name_2", [(u, "relation%d" % k) for k, u in enumerate(r2)] +
          [(u, "attribute%d" % k if k % 2 else "attributes%d" % k) for k, u in enumerate(a2)])
    order = rng.permutation(n)
    n_tr, n_va = int(0.3 * n), int(0.1 * n)
    link = lambda i: (e1[i], e2[int(perm[i])])
    lines(division + "train_links", [link(int(i)) for i in order[:n_tr]])
    lines(division + "valid_links", [link(int(i)) for i in order[n_tr:n_tr + n_va]])
    lines(division + "test_links", [link(int(i)) for i in order[n_tr + n_va:]])
    vec_path = os.path.join(folder, "words.vec")
    with open(vec_path, "w", encoding="utf-8") as fh:
        fh.write("%d %d\n" % (len(WORDS), word_dim))  # header line: skipped by read_word2vec (wrong field count)
        for w in WORDS:
            fh.write(w + " " + " ".join("%.4f" % x for x in rng.standard_normal(word_dim)) + "\n")
    return vec_path


def write_args(path, training_data, output, word2vec_path, **overrides):
    """the keys of code/args.json at sizes a test can afford"""
    args = {
        "training_data": training_data, "output": output, "word2vec_path": word2vec_path, "dataset_division": "631/",
        "alignment_module": "swapping",
        "encoder_epoch": 2, "encoder_active": "thah", "encoder_normalize": True, "retrain_literal_embeds": True,
        "literal_normalize": True,
        "dim": 75, "learning_rate": 0.001, "optimizer": "Adagrad", "max_epoch": 3, "shared_learning_max_epoch": 2,
        "batch_size": 500, "entity_batch_size": 100, "attribute_batch_size": 200,
        "neg_triple_num": 10, "neg_sampling": "truncated", "truncated_epsilon": 0.9, "truncated_freq": 2,
        "batch_threads_num": 2, "test_threads_num": 2,
        "start_valid": 1, "eval_freq": 1, "stop_metric": "mrr", "top_k": [1, 5, 10, 50], "is_save": True,
        "orthogonal_weight": 2, "cv_name_weight": 1, "cv_weight": 1,
        "start_predicate_soft_alignment": 1, "predicate_soft_sim": 0.85, "predicate_init_sim": 0.90,
        "relation_learning_rate": 0.005, "ITC_learning_rate": 0.004,
    }
    args.update(overrides)
    with open(path, "w") as fh:
        json.dump(args, fh, indent=1)
    return args
