"""Full multi-view ITC training (BASELINE.json configs[2]) on the REAL DBP-WD-100K data, then Hits@k of
the relation, attribute, name and combined views -- once with the CPU oracle (`--impl oracle`, build
container) and once on the B200 (`--impl b200`) from IDENTICAL inputs: same initial tables and CNN
weights, same list shuffles, same cross-KG / entity batches (one numpy generator drives both) and the
same negatives (device sampler == its CPU restatement).

One epoch = MultiKE_CV.run (MultiKE_CSL.py:57-72) without the soft predicate alignment, which only
starts after epoch 10 (args.json: start_predicate_soft_alignment):
  train_relation_view_1epo                                 MultiKE_model.py:291-317
  train_cross_kg_entity_inference_relation_view_1epo       :349-369   (swapped relation triples)
  train_attribute_view_1epo                                :319-345   (weighted attribute triples, no negatives)
  train_cross_kg_entity_inference_attribute_view_1epo      :371-391   (swapped attribute triples)
  train_common_space_learning_1epo                         :458-473   (ITC, lr 0.004)
Evaluation = MultiKE_Late.valid (:14-36): validation links of KG1 against the valid + test entities
of KG2, views 'nv', 'rv', 'av', 'final'.

Inputs: tests/golden/dbp_wd_100k_relation.npz (tools/digest_dbp_wd.py) and
tests/golden/dbp_wd_100k_multiview.npz (tools/digest_dbp_wd_multiview.py); the name / value vector
tables are multike_b200.synthetic.literal_vectors of the recorded literal ids (the authors'
word-vector file is not available: SURVEY.md section 8c).
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
    """emb1[i] aligns with emb2[i]; emb2 may hold more candidates; rows get l2-normalised here"""
    e1 = torch.as_tensor(emb1, dtype=torch.float32)
    e2 = torch.as_tensor(emb2, dtype=torch.float32)
    e1 = e1 / e1.norm(dim=1, keepdim=True).clamp_min(1e-12)
    e2 = e2 / e2.norm(dim=1, keepdim=True).clamp_min(1e-12)
    n = e1.shape[0]
    ranks = np.empty(n, dtype=np.int64)
    for a in range(0, n, 2000):
        sim = e1[a:a + 2000] @ e2.T
        gold = sim[torch.arange(sim.shape[0]), torch.arange(a, a + sim.shape[0])]
        ranks[a:a + sim.shape[0]] = (sim > gold[:, None]).sum(1).numpy() + 1
    out = {"hits@%d" % k: float((ranks <= k).mean() * 100) for k in ks}
    out["mr"], out["mrr"] = float(ranks.mean()), float((1.0 / ranks).mean())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["oracle", "b200"], required=True)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--batch", type=int, default=5000)
    ap.add_argument("--neg", type=int, default=10)
    ap.add_argument("--out", default=None)
    ap.add_argument("--mode", choices=["itc", "ssl"], default="itc",
                    help="itc: MultiKE_CV epochs (common-space step every epoch); ssl: MultiKE_Late -- views only, "
                         "then --shared-epochs of orthogonal space mapping (MultiKE_model.py:241-261, :439-454)")
    ap.add_argument("--shared-epochs", type=int, default=10)
    args = ap.parse_args()
    itc = args.mode == "itc"
    g = np.load(os.path.join(ROOT, "tests", "golden", "dbp_wd_100k_relation.npz"))
    mvw = np.load(os.path.join(ROOT, "tests", "golden", "dbp_wd_100k_multiview.npz"))
    from multike_b200 import synthetic
    from multike_b200.relation_view import clipped_slice, split_batch
    from oracle import attr_cnn as oc
    from oracle.tf_semantics import xavier_truncated_normal

    t1, t2 = g["triples1"].copy(), g["triples2"].copy()
    sup = np.concatenate([g["sup1"], g["sup2"]])
    f1, f2 = np.concatenate([t1, g["sup1"]]), np.concatenate([t2, g["sup2"]])
    n_ent, n_rel, split = int(g["entities_num"]), int(g["relations_num"]), len(g["entities1"])
    n_attr = int(mvw["attributes_num"])
    a1 = np.concatenate([mvw["attr1"].astype(np.float64), mvw["attr1_w"].astype(np.float64)[:, None]], 1)
    a2 = np.concatenate([mvw["attr2"].astype(np.float64), mvw["attr2_w"].astype(np.float64)[:, None]], 1)
    sup_attr = np.concatenate([mvw["sup_attr1"], mvw["sup_attr2"]]).astype(np.int64)
    names = synthetic.literal_vectors(mvw["name_literal"], 75)       # local_name_vectors (normalised rows)
    values = synthetic.literal_vectors(mvw["value_literal"], 75)     # value_vectors
    valid, test = g["valid_links"], g["test_links"]
    cand2 = np.concatenate([valid[:, 1], test[:, 1]])
    dim, B, K, lr, itc_lr, seed = 75, args.batch, args.neg, 0.001, 0.004, 7
    gen = torch.Generator().manual_seed(20190754)
    init = {name: xavier_truncated_normal((rows, dim), gen) for name, rows in
            (("rv_ent", n_ent), ("rel", n_rel), ("av_ent", n_ent), ("attr", n_attr), ("ent", n_ent))}
    thetas = [oc.init_theta(dim, generator=torch.Generator().manual_seed(1000 + k)).float() for k in range(2)]
    maps0 = []
    mgen = torch.Generator().manual_seed(77)
    for _ in range(3):   # tf.initializers.orthogonal(): QR of a normal matrix, signs fixed by diag(R)
        qm, rm = torch.linalg.qr(torch.randn(dim, dim, generator=mgen))
        maps0.append(qm * torch.sign(torch.diagonal(rm)))
    maps0 = torch.stack(maps0).float().contiguous()   # name, relation, attribute view -> shared space
    rng = np.random.default_rng(99)   # list shuffles, cross-KG batches, entity batches: identical in both runs
    steps = -(-(len(t1) + len(t2)) // B)
    ck_steps = -(-len(sup) // B)
    attr_steps = -(-(len(a1) + len(a2)) // B)
    cka_steps = -(-len(sup_attr) // B)
    ent_steps = -(-n_ent // B)
    log, t_start = [], time.time()

    def evaluate(tables):
        res = {}
        for view, emb in tables.items():
            res[view] = hits(emb[valid[:, 0]], emb[cand2])
        return res

    if args.impl == "oracle":
        torch.set_num_threads(os.cpu_count() or 1)
        from oracle import device_sampler as ds
        from oracle import relation_view as orv
        from oracle.tf_semantics import adagrad_dense_, l2_normalize
        ent, rel = orv.DenseTable(init["rv_ent"], True), orv.DenseTable(init["rel"], True)
        av, attr, fin = orv.DenseTable(init["av_ent"], True), orv.DenseTable(init["attr"], False), orv.DenseTable(init["ent"], True)
        N, V = torch.as_tensor(names), torch.as_tensor(values)
        th = [t.clone() for t in thetas]
        th_acc = [torch.full_like(t, 0.1) for t in th]
        kg1 = ds.KG(entity_base=0, n_entities=split, triples=f1)
        kg2 = ds.KG(entity_base=split, n_entities=n_ent - split, triples=f2)

        def attr_step(rows, k, slot, weighted, scale):
            ih, ia, iv = (torch.as_tensor(rows[:, c].astype(np.int64)) for c in range(3))
            w = torch.as_tensor(rows[:, 3].astype(np.float32)) if weighted else None
            ve, va, t = av.var.clone().requires_grad_(True), attr.var.clone().requires_grad_(True), th[k].clone().requires_grad_(True)
            loss = oc.attribute_cnn_loss(l2_normalize(ve, 1)[ih], va[ia], V[iv], w, t, dim, scale=scale)
            ge, ga, gt = torch.autograd.grad(loss, [ve, va, t])
            with torch.no_grad():
                adagrad_dense_(av.var, av.acc(slot), ge, lr)
                adagrad_dense_(attr.var, attr.acc(slot), ga, lr)
                adagrad_dense_(th[k], th_acc[k], gt, lr)
            return float(loss)

        gstep = 0
        for epoch in range(1, args.epochs + 1):
            rec = {"epoch": epoch}
            tot, npos = 0.0, 0
            b1, b2 = split_batch(len(t1), len(t2), B)
            for s in range(steps):
                s1, e1 = clipped_slice(len(t1), b1, s)
                s2, e2 = clipped_slice(len(t2), b2, s)
                p1, p2 = t1[s1:e1], t2[s2:e2]
                neg = ds.sample_batch_fast(p1, kg1, p2, kg2, K, seed, gstep)
                pos = np.concatenate([p1, p2])
                loss, _, _ = orv.relation_view_step(ent, rel, pos[:, 0], pos[:, 1], pos[:, 2], neg[:, 0], neg[:, 1],
                                                    neg[:, 2], lr, slot="relation")
                tot, npos, gstep = tot + loss, npos + len(pos), gstep + 1
            rec["rel_loss"] = tot / npos
            t1, t2 = t1[rng.permutation(len(t1))], t2[rng.permutation(len(t2))]
            e, ck = np.zeros(0, np.int64), 0.0
            for s in range(ck_steps):
                P = sup[rng.permutation(len(sup))[:B]]
                loss, _, _ = orv.relation_view_step(ent, rel, P[:, 0], P[:, 1], P[:, 2], e, e, e, lr, slot="ckge", pos_scale=2.0)
                ck += loss
            rec["ckge_rel_loss"] = ck / (ck_steps * B)
            tot, cnt = 0.0, 0
            c1, c2 = split_batch(len(a1), len(a2), B)
            for s in range(attr_steps):
                s1, e1 = clipped_slice(len(a1), c1, s)
                s2, e2 = clipped_slice(len(a2), c2, s)
                rows = np.concatenate([a1[s1:e1], a2[s2:e2]])
                tot, cnt = tot + attr_step(rows, 0, "attribute", True, 1.0), cnt + len(rows)
            rec["attr_loss"] = tot / cnt
            a1, a2 = a1[rng.permutation(len(a1))], a2[rng.permutation(len(a2))]
            ck = 0.0
            for s in range(cka_steps):
                P = sup_attr[rng.permutation(len(sup_attr))[:B]]
                ck += attr_step(np.concatenate([P, np.ones((len(P), 1))], 1), 1, "ckge_attribute", False, 2.0)
            rec["ckge_attr_loss"] = ck / (cka_steps * B)
            cs = 0.0
            for s in range(ent_steps if itc else 0):
                idx = torch.as_tensor(rng.permutation(n_ent)[:B].astype(np.int64))
                vs = [t.var.clone().requires_grad_(True) for t in (fin, ent, av)]
                F, R, A = (l2_normalize(v, 1)[idx] for v in vs)
                loss = ((F - N[idx]) ** 2).sum() + ((F - R) ** 2).sum() + ((F - A) ** 2).sum()
                grads = torch.autograd.grad(loss, vs)
                with torch.no_grad():
                    for t, gr in zip((fin, ent, av), grads):
                        adagrad_dense_(t.var, t.acc("cross_name"), gr, itc_lr)
                cs += float(loss)
            rec["common_loss"] = cs / (ent_steps * B)
            rec["elapsed_s"] = time.time() - t_start
            log.append(rec)
            print(json.dumps(rec), flush=True)
        if not itc:
            from oracle import losses as ol
            M, M_acc, eye = maps0.clone(), torch.full_like(maps0, 0.1), torch.eye(dim)
            R_const, A_const = ent.view().detach(), av.view().detach()      # only "shared*" variables train (:257)
            for epoch in range(1, args.shared_epochs + 1):
                tot = 0.0
                for s in range(ent_steps):
                    idx = torch.as_tensor(rng.permutation(n_ent)[:B].astype(np.int64))
                    v, m = fin.var.clone().requires_grad_(True), M.clone().requires_grad_(True)
                    Fv = l2_normalize(v, 1)[idx]
                    loss = sum(ol.space_mapping_loss(x[idx], Fv, m[k], eye, 2) for k, x in enumerate((N, R_const, A_const)))
                    gv, gm = torch.autograd.grad(loss, [v, m])
                    with torch.no_grad():
                        adagrad_dense_(fin.var, fin.acc("shared_comb"), gv, lr)
                        adagrad_dense_(M, M_acc, gm, lr)
                    tot += float(loss)
                rec = {"shared_epoch": epoch, "mapping_loss": tot / (ent_steps * B), "elapsed_s": time.time() - t_start}
                log.append(rec)
                print(json.dumps(rec), flush=True)
        tables = {"nv": names, "rv": ent.view().numpy(), "av": av.view().numpy(), "final": fin.view().numpy()}
    else:
        from multike_b200 import tables as T
        from multike_b200.attr_view import AttrCNN
        from multike_b200.relation_view import RelationView
        rv = RelationView(n_ent, n_rel, dim, t1, t2, split, batch_size=B, neg_num=K, lr=lr, seed=seed,
                          ent_init=init["rv_ent"], rel_init=init["rel"], filter1=f1, filter2=f2)
        av = T.EmbeddingTable(n_ent, dim, True, "cuda", init=init["av_ent"], flags=True, grad_replicas=1)
        attr = T.EmbeddingTable(n_attr, dim, False, "cuda", init=init["attr"])
        fin = T.EmbeddingTable(n_ent, dim, True, "cuda", init=init["ent"], flags=True, grad_replicas=1)
        N = T.EmbeddingTable(n_ent, dim, False, "cuda", init=names, trainable=False)
        V = T.EmbeddingTable(len(values), dim, False, "cuda", init=values, trainable=False)
        cnns = [AttrCNN(dim, "cuda", theta=t) for t in thetas]
        sup_d, sup_attr_d = torch.from_numpy(sup).cuda(), torch.from_numpy(sup_attr).cuda()
        a1_d, a2_d = torch.from_numpy(a1).cuda(), torch.from_numpy(a2).cuda()
        acc = T.new_loss_accumulator()

        def attr_step(rows, k, slot, weighted, scale):
            ih, ia, iv = (rows[:, c].to(torch.int32).contiguous() for c in range(3))
            w = rows[:, 3].to(torch.float32).contiguous() if weighted else None
            cnns[k].fwd_bwd(av, attr, V, ih, ia, iv, acc, w=w, scale=scale)
            av.apply_adagrad(slot, lr)
            attr.apply_adagrad(slot, lr)
            cnns[k].apply_adagrad(slot, lr)

        for epoch in range(1, args.epochs + 1):
            rec = {"epoch": epoch}
            trained = rv.train_steps(0, steps)
            rec["rel_loss"] = float(rv.step_losses.sum().item()) / trained
            p1, p2 = rng.permutation(len(t1)), rng.permutation(len(t2))
            rv.triples1.copy_(rv.triples1[torch.from_numpy(p1).cuda()])
            rv.triples2.copy_(rv.triples2[torch.from_numpy(p2).cuda()])
            acc.zero_()
            for s in range(ck_steps):
                pick = torch.from_numpy(rng.permutation(len(sup))[:B]).cuda()
                T.rel_step_structured(rv.ent, rv.rel, sup_d[pick].contiguous(), None, None, 0, acc, pos_scale=2.0)
                T.apply_adagrad_pair(rv.ent, rv.ent.adagrad_slot("ckge"), lr, rv.rel, rv.rel.adagrad_slot("ckge"), lr)
            rec["ckge_rel_loss"] = float(acc.item()) / (ck_steps * B)
            acc.zero_()
            cnt = 0
            c1, c2 = split_batch(len(a1_d), len(a2_d), B)
            for s in range(attr_steps):
                s1, e1 = clipped_slice(len(a1_d), c1, s)
                s2, e2 = clipped_slice(len(a2_d), c2, s)
                rows = torch.cat([a1_d[s1:e1], a2_d[s2:e2]])
                attr_step(rows, 0, "attribute", True, 1.0)
                cnt += rows.shape[0]
            rec["attr_loss"] = float(acc.item()) / cnt
            a1_d = a1_d[torch.from_numpy(rng.permutation(len(a1_d))).cuda()]
            a2_d = a2_d[torch.from_numpy(rng.permutation(len(a2_d))).cuda()]
            acc.zero_()
            for s in range(cka_steps):
                P = sup_attr_d[torch.from_numpy(rng.permutation(len(sup_attr))[:B]).cuda()]
                attr_step(torch.cat([P.double(), torch.ones(len(P), 1, dtype=torch.float64, device="cuda")], 1), 1,
                          "ckge_attribute", False, 2.0)
            rec["ckge_attr_loss"] = float(acc.item()) / (cka_steps * B)
            acc.zero_()
            for s in range(ent_steps if itc else 0):
                pick = torch.from_numpy(rng.permutation(n_ent)[:B].astype(np.int32)).cuda()
                T.align_fwd_bwd(fin, N, rv.ent, av, pick, acc, name_weight=1.0, scale=1.0)
                for t in (fin, rv.ent, av):
                    t.apply_adagrad("cross_name", itc_lr)
            rec["common_loss"] = float(acc.item()) / (ent_steps * B)
            rec["elapsed_s"] = time.time() - t_start
            log.append(rec)
            print(json.dumps(rec), flush=True)
        if not itc:
            from multike_b200 import _cabi
            from multike_b200.refapi import losses as L
            lib = _cabi.load()
            M = maps0.cuda().contiguous()
            M_grad, M_acc, eye = torch.zeros_like(M), torch.full_like(M, T.ADAGRAD_INIT), torch.eye(dim, device="cuda")
            for epoch in range(1, args.shared_epochs + 1):
                tot = torch.zeros((), dtype=torch.float64, device="cuda")
                for s in range(ent_steps):
                    idx = torch.from_numpy(rng.permutation(n_ent)[:B].astype(np.int32)).cuda()
                    final = fin.export(idx).requires_grad_(True)
                    views = (N.export(idx), rv.ent.export(idx), av.export(idx))
                    m = M.detach().clone().requires_grad_(True)
                    loss = sum(L.space_mapping_loss(x, final, m[k], eye, 2) for k, x in enumerate(views))
                    g_final, g_m = torch.autograd.grad(loss, [final, m])
                    fin.grad[:, :dim].index_add_(0, idx.long(), g_final)     # ids are distinct (random.sample)
                    if fin.touched is not None:
                        fin.touched[idx.long()] = 1
                    fin.apply_adagrad("shared_comb", lr)
                    M_grad.copy_(g_m)
                    _cabi.check(lib.mke_dense_apply_adagrad(M.data_ptr(), M_grad.data_ptr(), M_acc.data_ptr(), M.numel(),
                                                            float(lr), _cabi.current_stream()))
                    tot += loss.detach().double()
                rec = {"shared_epoch": epoch, "mapping_loss": float(tot) / (ent_steps * B), "elapsed_s": time.time() - t_start}
                log.append(rec)
                print(json.dumps(rec), flush=True)
        tables = {"nv": names, "rv": rv.ent.eval(), "av": av.eval(), "final": fin.eval()}
    if not itc:   # MultiKE_Late.valid(embed_choice='avg'): the sum of the three (normalised) views
        tables["avg"] = tables["nv"] + tables["rv"] + tables["av"]
    res = evaluate(tables)
    summary = {"impl": args.impl, "mode": args.mode, "shared_epochs": 0 if itc else args.shared_epochs, "epochs": args.epochs, "batch": B, "neg": K, "valid_links": int(len(valid)),
               "candidates": int(len(cand2)), "train_seconds": time.time() - t_start, "views": res, "log": log}
    print(json.dumps({k: v for k, v in summary.items() if k != "log"}))
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(summary, fh, indent=1)


if __name__ == "__main__":
    main()
