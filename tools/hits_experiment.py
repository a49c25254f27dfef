"""Relation-view-only training on the REAL DBP-WD-100K relation triples, then Hits@k on the
validation links -- run once with the CPU oracle (`--impl oracle`, build container) and once with
the B200 path (`--impl b200`, GPU box) on IDENTICAL inputs: same init tables, same list shuffles,
same cross-KG batches, same negatives (the device sampler and its CPU restatement are bit-exact).

Per epoch, as MultiKE_CSL.run does for the relation view (MultiKE_CSL.py:57-63):
  train_relation_view_1epo                             (MultiKE_model.py:291-317)
  train_cross_kg_entity_inference_relation_view_1epo   (MultiKE_model.py:349-369, swapped sup triples)
Evaluation = base/evaluation.py valid(): cosine similarity of the normalised rv_ent_embeds rows of
the link entities, rank of the gold counterpart (greedy_alignment, base/alignment.py:8-79).

Inputs come from tests/golden/dbp_wd_100k_relation.npz (tools/digest_dbp_wd.py).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def hits(emb1, emb2, ks=(1, 5, 10, 50)):
    """emb1[i] should align with emb2[i]; rows are l2-normalised."""
    n = emb1.shape[0]
    ranks = np.empty(n, dtype=np.int64)
    e2 = torch.as_tensor(emb2)
    for a in range(0, n, 2000):
        sim = torch.as_tensor(emb1[a:a + 2000]) @ e2.T
        gold = sim[torch.arange(sim.shape[0]), torch.arange(a, a + sim.shape[0])]
        ranks[a:a + sim.shape[0]] = (sim > gold[:, None]).sum(1).numpy() + 1
    out = {"hits@%d" % k: float((ranks <= k).mean() * 100) for k in ks}
    out["mr"] = float(ranks.mean())
    out["mrr"] = float((1.0 / ranks).mean())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["oracle", "b200"], required=True)
    ap.add_argument("--epochs", type=int, default=20)
    ap.add_argument("--batch", type=int, default=5000)
    ap.add_argument("--neg", type=int, default=10)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    g = np.load(os.path.join(ROOT, "tests", "golden", "dbp_wd_100k_relation.npz"))
    t1, t2 = g["triples1"].copy(), g["triples2"].copy()
    sup = np.concatenate([g["sup1"], g["sup2"]])
    f1, f2 = np.concatenate([t1, g["sup1"]]), np.concatenate([t2, g["sup2"]])  # filter set incl. sup triples
    n_ent, n_rel, split = int(g["entities_num"]), int(g["relations_num"]), len(g["entities1"])
    valid = g["valid_links"]
    dim, B, K, lr, seed = 75, args.batch, args.neg, 0.001, 7
    gen = torch.Generator().manual_seed(20190754)
    from oracle.tf_semantics import xavier_truncated_normal
    ent0 = xavier_truncated_normal((n_ent, dim), gen)
    rel0 = xavier_truncated_normal((n_rel, dim), gen)
    rng = np.random.default_rng(99)  # list shuffles and cross-KG batches: identical in both runs
    steps = -(-(len(t1) + len(t2)) // B)
    ck_steps = -(-len(sup) // B)
    log = []
    t_start = time.time()

    if args.impl == "oracle":
        torch.set_num_threads(os.cpu_count() or 1)
        from oracle import device_sampler as ds
        from oracle import relation_view as orv
        from multike_b200.relation_view import clipped_slice, split_batch
        ent, rel = orv.DenseTable(ent0, True, torch.float32), orv.DenseTable(rel0, True, torch.float32)
        kg1 = ds.KG(entity_base=0, n_entities=split, triples=f1)
        kg2 = ds.KG(entity_base=split, n_entities=n_ent - split, triples=f2)
        gstep = 0
        for epoch in range(1, args.epochs + 1):
            tot, npos = 0.0, 0
            b1, b2 = split_batch(len(t1), len(t2), B)
            for s in range(steps):
                a1, e1 = clipped_slice(len(t1), b1, s)
                a2, e2 = clipped_slice(len(t2), b2, s)
                p1, p2 = t1[a1:e1], t2[a2:e2]
                neg = ds.sample_batch_fast(p1, kg1, p2, kg2, K, seed, gstep)
                pos = np.concatenate([p1, p2])
                loss, _, _ = orv.relation_view_step(ent, rel, pos[:, 0], pos[:, 1], pos[:, 2], neg[:, 0], neg[:, 1],
                                                    neg[:, 2], lr, slot="relation")
                tot += loss
                npos += len(pos)
                gstep += 1
            t1, t2 = t1[rng.permutation(len(t1))], t2[rng.permutation(len(t2))]
            e = np.zeros(0, np.int64)
            ck = 0.0
            for s in range(ck_steps):
                P = sup[rng.permutation(len(sup))[:B]]
                loss, _, _ = orv.relation_view_step(ent, rel, P[:, 0], P[:, 1], P[:, 2], e, e, e, lr, slot="ckge",
                                                    pos_scale=2.0)
                ck += loss
            log.append({"epoch": epoch, "rel_loss": tot / npos, "ckge_loss": ck / (ck_steps * B),
                        "elapsed_s": time.time() - t_start})
            print(json.dumps(log[-1]), flush=True)
        E = ent.view().numpy()
    else:
        from multike_b200 import tables as T
        from multike_b200.relation_view import RelationView
        rv = RelationView(n_ent, n_rel, dim, t1, t2, split, batch_size=B, neg_num=K, lr=lr, seed=seed, ent_init=ent0,
                          rel_init=rel0, filter1=f1, filter2=f2)
        sup_d = torch.from_numpy(sup).cuda()
        acc = T.new_loss_accumulator()
        for epoch in range(1, args.epochs + 1):
            trained = rv.train_steps(0, steps)
            rel_loss = float(rv.step_losses.sum().item()) / trained
            p1, p2 = rng.permutation(len(t1)), rng.permutation(len(t2))
            rv.triples1.copy_(rv.triples1[torch.from_numpy(p1).cuda()])
            rv.triples2.copy_(rv.triples2[torch.from_numpy(p2).cuda()])
            acc.zero_()
            for s in range(ck_steps):
                pick = torch.from_numpy(rng.permutation(len(sup))[:B]).cuda()
                T.rel_step_structured(rv.ent, rv.rel, sup_d[pick].contiguous(), None, None, 0, acc, pos_scale=2.0)
                T.apply_adagrad_pair(rv.ent, rv.ent.adagrad_slot("ckge"), lr, rv.rel, rv.rel.adagrad_slot("ckge"), lr)
            log.append({"epoch": epoch, "rel_loss": rel_loss, "ckge_loss": float(acc.item()) / (ck_steps * B),
                        "elapsed_s": time.time() - t_start})
            print(json.dumps(log[-1]), flush=True)
        E = rv.ent.eval()
        # the same metrics from the fused device evaluator (mke_sim_rank), candidates = the valid links'
        # KG2 entities (base/evaluation.py valid()): must agree with the numpy/torch ranking below
        from multike_b200 import similarity as S
        rank, _ = S.sim_rank(rv.ent.var, rv.ent.var, idx1=valid[:, 0].astype(np.int32),
                             idx2=valid[:, 1].astype(np.int32), normalize=True, dim=dim)
        r = rank.cpu().numpy().astype(np.int64) + 1
        device_eval = {"hits@%d" % k: float((r <= k).mean() * 100) for k in (1, 5, 10, 50)}
        device_eval.update(mr=float(r.mean()), mrr=float((1.0 / r).mean()))
        # ... and from the fp32 FMA tiles (the baseline implementation of the same evaluator): the tensor-core tiles
        # (3xTF32) must not move a single Hits@k on the real rows; against the valid + test candidates too (valid())
        from multike_b200 import _cabi
        lib = _cabi.load()
        cand = np.concatenate([valid[:, 1], g["test_links"][:, 1]]).astype(np.int32)
        both = {}
        for name, mode in (("tcgen05", 1), ("fma", 0)):
            prev = lib.mke_sim_use_tensor_cores(mode)
            rk, t1 = S.sim_rank(rv.ent.var, rv.ent.var, idx1=valid[:, 0].astype(np.int32), idx2=cand, normalize=True, dim=dim)
            lib.mke_sim_use_tensor_cores(prev)
            rr = rk.cpu().numpy().astype(np.int64) + 1
            both[name] = {"rank": rr, "top1": t1.cpu().numpy(),
                          "hits": {"hits@%d" % k: float((rr <= k).mean() * 100) for k in (1, 5, 10, 50)}}
        device_eval["tcgen05_vs_fma"] = {
            "candidates": int(len(cand)), "hits_tcgen05": both["tcgen05"]["hits"], "hits_fma": both["fma"]["hits"],
            "ranks_equal_fraction": float((both["tcgen05"]["rank"] == both["fma"]["rank"]).mean()),
            "top1_equal_fraction": float((both["tcgen05"]["top1"] == both["fma"]["top1"]).mean())}
    res = hits(E[valid[:, 0]], E[valid[:, 1]])
    summary = {"impl": args.impl, "epochs": args.epochs, "batch": B, "neg": K, "valid_links": int(len(valid)),
               "train_seconds": time.time() - t_start, **res, "log": log}
    if args.impl == "b200":
        summary["device_evaluator"] = device_eval
    print(json.dumps({k: v for k, v in summary.items() if k != "log"}))
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(summary, fh, indent=1)


if __name__ == "__main__":
    main()
