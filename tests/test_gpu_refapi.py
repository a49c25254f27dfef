"""The host-side mirror of the reference's operator surface (multike_b200/refapi): losses.py as
free differentiable functions and the relation-view part of class MultiKE, against the oracle."""
import types

import numpy as np
import pytest
import torch

from oracle import device_sampler as ds
from oracle import losses as ol
from oracle import relation_view as orv

pytestmark = pytest.mark.gpu


def _rand(n, d, seed, scale=0.5):
    return torch.tensor(np.random.default_rng(seed).normal(0, scale, (n, d)), dtype=torch.float64)


def test_losses_mirror_values_and_gradients():
    from multike_b200.refapi import losses as ml
    n, d = 37, 75
    mats = [_rand(n, d, s) for s in range(6)]
    w = torch.tensor(np.random.default_rng(9).choice([1.0, 0.9, 0.5], n))
    cases = [
        (ml.relation_logistic_loss, ol.relation_logistic_loss, mats),
        (ml.attribute_logistic_loss, ol.attribute_logistic_loss, mats[:3] + [w] + mats[3:] + [w]),
        (ml.relation_logistic_loss_wo_negs, ol.relation_logistic_loss_wo_negs, mats[:3]),
        (ml.attribute_logistic_loss_wo_negs, ol.attribute_logistic_loss_wo_negs, mats[:3]),
        (ml.logistic_loss_wo_negs, ol.logistic_loss_wo_negs, mats[:3] + [w]),
        (ml.alignment_loss, ol.alignment_loss, mats[:2]),
    ]
    for mine, oracle, args in cases:
        ref_in = [a.clone().requires_grad_(a.dim() == 2) for a in args]
        ref = oracle(*ref_in)
        ref.backward()
        dev_in = [a.float().cuda().requires_grad_(a.dim() == 2) for a in args]
        got = mine(*dev_in)
        assert got.dim() == 0 and got.dtype == torch.float32
        (3.0 * got).backward()  # upstream gradient is honoured
        assert float(got) == pytest.approx(float(ref), rel=1e-5)
        for a, b in zip(dev_in, ref_in):
            if b.grad is not None:
                np.testing.assert_allclose(a.grad.cpu().numpy(), 3.0 * b.grad.numpy(), rtol=1e-4, atol=1e-5)
    # matmul-based terms
    eye = torch.eye(d, dtype=torch.float64)
    M = torch.linalg.qr(_rand(d, d, 20))[0]
    ref = ol.space_mapping_loss(mats[0], mats[1], M, eye, 2.0)
    got = ml.space_mapping_loss(mats[0].float().cuda(), mats[1].float().cuda(), M.float().cuda(), eye.float().cuda(), 2.0)
    assert float(got) == pytest.approx(float(ref), rel=1e-4)
    assert float(ml.orthogonal_loss(2 * eye.float().cuda(), eye.float().cuda())) == pytest.approx(9.0 * d)
    with pytest.raises(TypeError):
        ml.alignment_loss(mats[0].float(), mats[1].float())  # CPU tensors: no fallback


def _fake_model(golden, batch_size=200, K=10):
    g = golden("ref_batch_relation.npz")
    n_ent = int(g["n_ent"])
    t1 = [tuple(int(x) for x in r) for r in g["triples1"]]
    t2 = [tuple(int(x) for x in r) for r in g["triples2"]]
    sup1 = [tuple(int(x) for x in r) for r in g["sup1"]]
    sup2 = [tuple(int(x) for x in r) for r in g["sup2"]]

    def kg(trip, sup, lo):
        k = types.SimpleNamespace()
        k.local_relation_triples_list = list(trip)
        k.local_relation_triples_set = set(trip) | set(sup)  # aliasing quirk: sup triples are in the filter set
        k.entities_list = list(range(lo, lo + n_ent))
        k.sup_relation_triples_list = list(sup)
        return k

    kgs = types.SimpleNamespace(kg1=kg(t1, sup1, 0), kg2=kg(t2, sup2, n_ent), entities_num=2 * n_ent, relations_num=5,
                                attributes_num=3, useful_entities_list1=[0, 3, 5], useful_entities_list2=[41, 42])
    data = types.SimpleNamespace(kgs=kgs)
    args = types.SimpleNamespace(alignment_module='swapping', output='/tmp/mke_out/', training_data='x/DBP_toy/', dim=75,
                                 batch_size=batch_size, neg_triple_num=K, learning_rate=0.001, seed=3)
    from multike_b200.refapi.MultiKE_model import MultiKE
    m = MultiKE(data, args, None)
    m._define_variables()
    m._define_name_view_graph()
    m._define_relation_view_graph()
    m._define_cross_kg_entity_reference_relation_view_graph()
    m._define_cross_kg_relation_reference_graph()
    return m, (n_ent, t1, t2, sup1, sup2)


def test_multike_relation_view_epochs_match_oracle(golden, capsys):
    m, (n_ent, t1, t2, sup1, sup2) = _fake_model(golden)
    K, B = 10, 200
    ent0, rel0 = m.rv_ent_embeds.raw().astype(np.float64), m.rel_embeds.raw().astype(np.float64)
    oe, orl = orv.DenseTable(ent0, True, torch.float64), orv.DenseTable(rel0, True, torch.float64)
    all1, all2 = np.array(t1 + sup1), np.array(t2 + sup2)
    ok1 = ds.KG(entity_base=0, n_entities=n_ent, triples=all1)
    ok2 = ds.KG(entity_base=n_ent, n_entities=n_ent, triples=all2)
    steps = int(np.ceil((len(t1) + len(t2)) / B))
    # epoch 1 of the relation view: same batches, same draws as the CPU restatement
    a1, a2 = np.array(t1), np.array(t2)
    tot, npos = 0.0, 0
    for step in range(steps):
        (s1, e1), (s2, e2) = m._rv.step_slices(step)
        q1, q2 = a1[s1:e1], a2[s2:e2]
        neg = ds.sample_batch(q1, ok1, q2, ok2, K, 3, step)
        pos = np.concatenate([q1, q2])
        loss, _, _ = orv.relation_view_step(oe, orl, pos[:, 0], pos[:, 1], pos[:, 2], neg[:, 0], neg[:, 1], neg[:, 2], 0.001)
        tot += loss
        npos += len(pos)
    got = m.train_relation_view_1epo(1, steps, None, None, None, None)
    out = capsys.readouterr().out
    assert "epoch 1 of rel. view, avg. loss: {:.4f}".format(tot / npos) in out
    assert got == pytest.approx(tot / npos, rel=1e-5)
    np.testing.assert_allclose(m.rv_ent_embeds.raw(), oe.var.numpy(), rtol=0, atol=1e-5)
    # cross-KG entity inference (positives only, loss x2, its own Adagrad accumulators): one step
    # holds all sup triples, so the random order inside the batch does not matter
    sup = m.kgs.kg1.sup_relation_triples_list + m.kgs.kg2.sup_relation_triples_list
    m.args.batch_size = len(sup)
    e = np.zeros(0, np.int64)
    P = np.array(sup)
    ref_loss, _, _ = orv.relation_view_step(oe, orl, P[:, 0], P[:, 1], P[:, 2], e, e, e, 0.001, slot="ckge", pos_scale=2.0)
    got = m.train_cross_kg_entity_inference_relation_view_1epo(1, sup)
    assert got == pytest.approx(ref_loss / len(sup), rel=1e-5)
    np.testing.assert_allclose(m.rv_ent_embeds.raw(), oe.var.numpy(), rtol=0, atol=1e-5)
    # weighted cross-KG relation inference
    w = np.random.default_rng(0).choice([1.0, 0.8, 0.6], len(sup))
    supw = [(h, r, t, float(x)) for (h, r, t), x in zip(sup, w)]
    ref_loss, _, _ = orv.relation_view_step(oe, orl, P[:, 0], P[:, 1], P[:, 2], e, e, e, 0.001, slot="ckgp", pos_w=w,
                                            pos_scale=2.0)
    got = m.train_cross_kg_relation_inference_1epo(1, supw)
    assert got == pytest.approx(ref_loss / len(sup), rel=1e-5)
    np.testing.assert_allclose(m.rv_ent_embeds.raw(), oe.var.numpy(), rtol=0, atol=1e-5)
    np.testing.assert_allclose(m.rel_embeds.raw(), orl.var.numpy(), rtol=0, atol=1e-5)
    # reads used by the drivers
    e1 = m.eval_kg1_ent_embeddings()
    assert e1.shape == (n_ent, 75) and np.allclose(np.linalg.norm(e1, axis=1), 1.0, atol=1e-5)
    assert m.eval_kg2_useful_ent_embeddings().shape == (2, 75)
    assert m.rv_ent_embeds.eval(session=m.session).shape == (2 * n_ent, 75)


def test_multike_attribute_view_epochs_match_oracle(golden, capsys):
    """train_attribute_view_1epo / ckge-attr / ckga-attr (MultiKE_model.py:319-345, 371-391, 416-437)
    against the torch oracle of conv(): batches as attr_batch.py builds them (no negatives)."""
    from oracle import attr_cnn as oc
    from oracle.tf_semantics import l2_normalize
    from multike_b200 import tables as T
    from multike_b200.relation_view import clipped_slice, split_batch
    m, (n_ent, *_r) = _fake_model(golden)
    g = golden("ref_batch_attribute.npz")
    a1 = [(int(h), int(a) % 3, int(v), float(w)) for h, a, v, w in g["a1"]]
    a2 = [(int(h), int(a) % 3, int(v), float(w)) for h, a, v, w in g["a2"]]
    n_val = int(max(max(t[2] for t in a1), max(t[2] for t in a2))) + 1
    rng = np.random.default_rng(4)
    m.literal_embeds = T.EmbeddingTable(n_val, 75, False, "cuda", init=rng.normal(0, 0.3, (n_val, 75)), trainable=False)
    m.predicate_align_model = types.SimpleNamespace(attribute_triples_w_weights1=list(a1), attribute_triples_w_weights2=list(a2))
    m.args.attribute_batch_size = 128
    m._define_attribute_view_graph()
    m._define_cross_kg_entity_reference_attribute_view_graph()
    m._define_cross_kg_attribute_reference_graph()
    dim, lr = 75, 0.001
    Ve = torch.tensor(m.av_ent_embeds.raw().astype(np.float64))
    Va = torch.tensor(m.attr_embeds.raw().astype(np.float64))
    Vv = torch.tensor(m.literal_embeds.raw().astype(np.float64))
    th = m._attr_cnn.theta.double().cpu()
    accs = [torch.full_like(x, 0.1) for x in (Ve, Va, th)]
    steps = int(np.ceil((len(a1) + len(a2)) / 128))
    b1, b2 = split_batch(len(a1), len(a2), 128)
    A1, A2 = np.array(a1), np.array(a2)
    tot, cnt = 0.0, 0
    for step in range(steps):
        (s1, e1), (s2, e2) = clipped_slice(len(a1), b1, step), clipped_slice(len(a2), b2, step)
        rows = np.concatenate([A1[s1:e1], A2[s2:e2]])
        ih, ia, iv, w = rows[:, 0].astype(int), rows[:, 1].astype(int), rows[:, 2].astype(int), torch.tensor(rows[:, 3])
        ve, va, t = Ve.clone().requires_grad_(True), Va.clone().requires_grad_(True), th.clone().requires_grad_(True)
        loss = oc.attribute_cnn_loss(l2_normalize(ve, 1)[ih], va[ia], Vv[iv], w, t, dim)
        grads = torch.autograd.grad(loss, [ve, va, t])
        for x, a, gr in zip((Ve, Va, th), accs, grads):
            a += gr * gr
            x -= lr * gr * torch.rsqrt(a)
        tot += float(loss)
        cnt += len(rows)
    got = m.train_attribute_view_1epo(1, steps, None, None, None, None)
    assert "epoch 1 of att. view, avg. loss: {:.4f}".format(tot / cnt) in capsys.readouterr().out
    assert got == pytest.approx(tot / cnt, rel=2e-5)
    np.testing.assert_allclose(m.av_ent_embeds.raw(), Ve.numpy(), rtol=0, atol=1e-5)
    np.testing.assert_allclose(m.attr_embeds.raw(), Va.numpy(), rtol=0, atol=1e-5)
    np.testing.assert_allclose(m._attr_cnn.theta.cpu().numpy(), th.numpy(), rtol=0, atol=2e-5)
    # cross-KG variants: one batch holding every triple; own conv() weights and Adagrad slots
    sup = [(h, a, v) for h, a, v, _ in a1[:150]]
    m.args.attribute_batch_size = 10 ** 6
    th2 = m._ckge_attr_cnn.theta.double().cpu()
    P = np.array(sup)
    ve, t = Ve.clone().requires_grad_(True), th2.clone().requires_grad_(True)
    ref = oc.attribute_cnn_loss(l2_normalize(ve, 1)[P[:, 0]], Va[P[:, 1]], Vv[P[:, 2]], None, t, dim, scale=2.0)
    got = m.train_cross_kg_entity_inference_attribute_view_1epo(1, sup)
    assert got == pytest.approx(float(ref) / len(sup), rel=2e-5)
    assert not torch.equal(m._ckge_attr_cnn.theta, m._attr_cnn.theta)
    got = m.train_cross_kg_attribute_inference_1epo(1, a2[:100])
    assert np.isfinite(got) and got > 0


def test_multike_common_space_epoch_matches_oracle(golden, capsys):
    """train_common_space_learning_1epo (MultiKE_model.py:458-473) with one batch holding every
    entity (so the random batch order does not matter) against torch autograd."""
    from oracle.tf_semantics import l2_normalize
    m, (n_ent, *_rest) = _fake_model(golden)
    rng = np.random.default_rng(2)
    from multike_b200 import tables as T
    m.name_embeds = T.EmbeddingTable(2 * n_ent, 75, False, "cuda", init=rng.normal(0, 0.1, (2 * n_ent, 75)),
                                     trainable=False)
    m.args.entity_batch_size, m.args.ITC_learning_rate, m.args.cv_weight, m.args.cv_name_weight = 10 ** 6, 0.004, 1.0, 1.0
    m._define_common_space_learning_graph()
    ents = list(range(0, 2 * n_ent, 3))
    before = [t.raw().astype(np.float64) for t in (m.ent_embeds, m.rv_ent_embeds, m.av_ent_embeds)]
    N = torch.tensor(m.name_embeds.raw().astype(np.float64))
    V = [torch.tensor(b, requires_grad=True) for b in before]
    F, R, A = (l2_normalize(v, 1)[ents] for v in V)
    loss = ((F - N[ents]) ** 2).sum() + ((F - R) ** 2).sum() + ((F - A) ** 2).sum()
    grads = torch.autograd.grad(loss, V)
    got = m.train_common_space_learning_1epo(1, ents)
    assert got == pytest.approx(float(loss) / len(ents), rel=1e-5)
    assert "epoch 1 of common space learning, avg. loss: {:.4f}".format(float(loss) / len(ents)) in capsys.readouterr().out
    for t, b, g in zip((m.ent_embeds, m.rv_ent_embeds, m.av_ent_embeds), before, grads):
        want = b - 0.004 * g.numpy() / np.sqrt(0.1 + g.numpy() ** 2)
        np.testing.assert_allclose(t.raw(), want, rtol=0, atol=2e-6)


def test_multike_space_mapping_epoch_matches_oracle(golden):
    """train_shared_space_mapping_1epo (MultiKE_model.py:439-454): only ent_embeds and the three
    mappings move; one batch with every entity against torch autograd (float64)."""
    from oracle import losses as ol
    from oracle.tf_semantics import l2_normalize
    from multike_b200 import tables as T
    m, (n_ent, *_r) = _fake_model(golden)
    rng = np.random.default_rng(6)
    m.name_embeds = T.EmbeddingTable(2 * n_ent, 75, False, "cuda", init=rng.normal(0, 0.1, (2 * n_ent, 75)), trainable=False)
    m.args.entity_batch_size, m.args.orthogonal_weight = 10 ** 6, 2
    m._define_space_mapping_graph()
    ents = list(range(1, 2 * n_ent, 2))
    F0 = m.ent_embeds.raw().astype(np.float64)
    rv0, av0 = m.rv_ent_embeds.raw().copy(), m.av_ent_embeds.raw().copy()
    maps0 = m._maps.double().cpu()
    assert torch.allclose(maps0[0] @ maps0[0].T, torch.eye(75, dtype=torch.float64), atol=1e-5)   # orthogonal init
    V = torch.tensor(F0, requires_grad=True)
    M = maps0.clone().requires_grad_(True)
    Fv = l2_normalize(V, 1)[ents]
    X = [torch.tensor(m.name_embeds.raw().astype(np.float64))[ents],
         l2_normalize(torch.tensor(rv0.astype(np.float64)), 1)[ents], l2_normalize(torch.tensor(av0.astype(np.float64)), 1)[ents]]
    eye = torch.eye(75, dtype=torch.float64)
    loss = sum(ol.space_mapping_loss(x, Fv, M[k], eye, 2) for k, x in enumerate(X))
    gV, gM = torch.autograd.grad(loss, [V, M])
    got = m.train_shared_space_mapping_1epo(1, ents)
    assert got == pytest.approx(float(loss) / len(ents), rel=1e-4)
    np.testing.assert_allclose(m.ent_embeds.raw(), F0 - 0.001 * gV.numpy() / np.sqrt(0.1 + gV.numpy() ** 2), rtol=0, atol=5e-6)
    np.testing.assert_allclose(m._maps.cpu().numpy(), maps0.numpy() - 0.001 * gM.numpy() / np.sqrt(0.1 + gM.numpy() ** 2),
                               rtol=0, atol=5e-6)
    assert np.array_equal(m.rv_ent_embeds.raw(), rv0) and np.array_equal(m.av_ent_embeds.raw(), av0)  # not "shared*"


def test_multike_truncated_neighbours_are_used(golden):
    m, (n_ent, t1, t2, _, _) = _fake_model(golden, batch_size=100, K=5)
    rng = np.random.default_rng(1)
    nb1 = {e: [int(x) for x in rng.choice(np.arange(0, n_ent), 12, replace=False)] for e in range(0, n_ent)}
    nb2 = {e: [int(x) for x in rng.choice(np.arange(n_ent, 2 * n_ent), 12, replace=False)] for e in range(n_ent, 2 * n_ent)}
    before = m.rv_ent_embeds.raw().copy()
    loss = m.train_relation_view_1epo(1, m._rv.triple_steps, None, None, nb1, nb2)
    assert np.isfinite(loss) and loss > 0
    assert np.abs(m.rv_ent_embeds.raw() - before).max() > 0
    # the sampler now draws from the 12-entity lists only
    from multike_b200 import tables as T
    neg = T.sample_uniform(np.array(t1[:50], np.int32), m._rv.kg1, None, None, 5, 3, 0).cpu().numpy().reshape(-1, 5, 3)
    for (h, r, t), row in zip(t1[:50], neg):
        for nh, _, nt in row:
            assert (nh == h and nt in nb1[t]) or (nt == t and nh in nb1[h])
    with pytest.raises(AssertionError):
        from multike_b200.refapi.MultiKE_model import MultiKE
        MultiKE(m.data, types.SimpleNamespace(alignment_module='mapping', output='', training_data='x'), None)


@pytest.mark.parametrize("gemm,tol", [("tcgen05", 1e-5), ("cublas", 1e-5)])
def test_literal_autoencoder_matches_oracle(gemm, tol, capsys):
    """AutoEncoderModel (literal_encoder.py:19-144): all GEMMs on the hand-written tcgen05 kernel at
    fp32-equivalent precision (default) or on fp32 cuBLAS (the baseline), chain rule written out, Adagrad on the
    hand-written dense kernel; two epochs against the float64 oracle at ONE tolerance for both."""
    from multike_b200.refapi.literal_encoder import AutoEncoderModel
    from oracle import autoencoder as oa
    rng = np.random.default_rng(0)
    n, d_in, hidden = 230, 96, [64, 32, 16]
    data = rng.normal(0, 1, (n, d_in))
    init = [rng.normal(0, 1, s) for s in oa.shapes(d_in, hidden)]
    args = types.SimpleNamespace(dim=16, encoder_normalize=True, encoder_active="thah", learning_rate=0.01, batch_size=50,
                                 encoder_gemm=gemm)
    m = AutoEncoderModel(data, args, input_dimension=d_in, hidden_dimensions=list(hidden), init_params=init)
    o = oa.AutoEncoderOracle(init, 3, active="thah", normalize=True, lr=0.01)
    rows = data / np.linalg.norm(data, axis=1, keepdims=True)
    for epoch in (1, 2):
        want = sum(o.step(rows[a:a + 50]) for a in range(0, n, 50)) + 50
        got = m.train_one_epoch(epoch)
        assert got == pytest.approx(want, rel=tol)
        assert "epoch %d of literal encoder, loss: " % epoch in capsys.readouterr().out
    for p, q in zip(m.params, o.params):
        np.testing.assert_allclose(p.cpu().numpy(), q.numpy(), rtol=0, atol=50 * tol)
    enc = m.encoder_multi_batches(data)
    want_enc = o.encode(data)   # values of order 1e2 (N(0,1) weights): tolerance relative to the largest code
    np.testing.assert_allclose(enc, want_enc, rtol=0, atol=1e-5 * np.abs(want_enc).max())
    assert set(m.weights) == {"encoder_h0", "encoder_h1", "encoder_h2", "decoder_h0", "decoder_h1", "decoder_h2"}


def _hits1_lines(out, label):
    """Hits@1 of every 'quick results' line that follows a '<label> ... results:' header"""
    import re
    vals, armed = [], False
    for line in out.splitlines():
        if line.endswith("results:"):
            armed = line.startswith(label + " ")
        elif armed and line.startswith("quick results"):
            vals.append(float(re.search(r"=\s*\[\s*([0-9.]+)", line).group(1)))
            armed = False
    return vals


def test_drivers_run_itc_and_ssl_schedules(capsys, tmp_path):
    """refapi.MultiKE_CSL.MultiKE_CV.run (run_ITC.py) and refapi.MultiKE_Late.MultiKE_Late.run
    (run_SSL.py) on a small synthetic three-view dataset: the whole schedule executes on the
    device (views, cross-KG steps, soft predicate refresh, truncated neighbours, evaluation,
    save), and alignment quality rises above the name view alone."""
    import os
    import multiview_fixture as mv
    from multike_b200.refapi.MultiKE_CSL import MultiKE_CV
    from multike_b200.refapi.MultiKE_Late import MultiKE_Late
    data, args, pam = mv.make()
    args.output = str(tmp_path) + "/"
    model = MultiKE_CV(data, args, pam)
    model.run()
    out = capsys.readouterr().out
    for needle in ("epoch 4 of rel. view", "epoch 4 of att. view", "cross-kg entity inference in rel. view",
                   "cross-kg relation inference in rel. view", "cross-kg attribute inference in attr. view",
                   "epoch 4 of common space learning", "neighbor dict: 400", "Embeddings saved!", "final test results:"):
        assert needle in out, needle
    assert [p for p, _ in pam.updates] == []  # refresh is due at epochs that are multiples of 10 only
    nv, final = _hits1_lines(out, "nv"), _hits1_lines(out, "final")
    assert nv[0] >= 55                             # ~60 % of the names are shared exactly
    assert final[-1] > 10.0                        # chance is 0.4 %: four epochs already align the common space
    assert sorted(os.listdir(model.out_folder)) == ["attr_embeds.npy", "av_ent_embeds.npy", "ent_embeds.npy",
                                                    "nv_ent_embeds.npy", "rel_embeds.npy", "rv_ent_embeds.npy"]
    assert model._rv.kg1.neighbours is not None and model._rv.kg1.neighbours.shape == (800, 39)  # int((1 - 0.9) * 400) == 39
    # SSL schedule
    data, args, pam = mv.make(seed=1)
    args.output = str(tmp_path) + "/"
    late = MultiKE_Late(data, args, pam)
    late.run()
    out = capsys.readouterr().out
    for needle in ("avg valid results:", "wvag valid results:", "weights ", "epoch 3 of shared space learning",
                   "wvag test results:", "final test results:"):
        assert needle in out, needle
    assert [p for p, _ in pam.updates] == ["relation", "attribute"] * 2   # epochs 2 and 4
    # three short epochs of space mapping do not align the shared space yet: only the views are checked
    assert _hits1_lines(out, "avg")[-1] > 10.0 and len(_hits1_lines(out, "final")) == 2


@pytest.mark.parametrize("n,count", [(1, 1), (7, 7), (1000, 1000), (554173, 5000), (200000, 5000), (65536, 65536)])
def test_sample_distinct_is_a_sample_without_replacement(n, count):
    """mke_sample_distinct = random.sample(range(n), count) (the reference's cross-KG / entity batch draw): distinct indices
    of [0, n), a function of (seed, draw) only, a full permutation at count == n, and uniform (mean / spread / a chi-square
    over 16 bins of many draws)"""
    from multike_b200 import tables as T
    a = T.sample_distinct(n, count, seed=3, draw=1).cpu().numpy()
    assert a.shape == (count,) and a.min() >= 0 and a.max() < n and len(np.unique(a)) == count
    from oracle import device_sampler as ds   # index work: bit-exact against the CPU restatement
    assert a[:200].tolist() == ds.sample_distinct(n, min(count, 200), 3, 1)
    assert np.array_equal(a, T.sample_distinct(n, count, seed=3, draw=1).cpu().numpy())
    b = T.sample_distinct(n, count, seed=3, draw=2).cpu().numpy()
    if n > 1000:
        assert not np.array_equal(a, b) and (count == n or len(np.intersect1d(a, b)) < count)
    if n >= 200000:
        draws = np.concatenate([T.sample_distinct(n, count, seed=11, draw=k).cpu().numpy() for k in range(40)])
        hist = np.bincount(draws * 16 // n, minlength=16).astype(np.float64)
        expect = len(draws) / 16
        assert ((hist - expect) ** 2 / expect).sum() < 45.0          # chi-square, 15 degrees of freedom (p ~ 1e-4)
        assert abs(draws.mean() / n - 0.5) < 0.01
        # positions are independent of the output order: first and second half of a draw are alike
        assert abs(a[: count // 2].mean() - a[count // 2:].mean()) / n < 0.03
