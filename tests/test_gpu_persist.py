"""Persistent step kernel (mke_rel_view_t.variant = 4, csrc/mke_rel_persist.cu): a run of training steps in
ONE cooperative launch must end in the same tables, losses and negatives as the same steps issued as one
launch per phase (variant 3), which tests/test_gpu_parity.py pins to the oracle and the golden vectors.

Tolerances: negatives bit-exact (same counter-based draws); losses rel 1e-6 (fp32 partial sums in another
order); rows abs 2e-6 per step (float atomics in another order), scaled by the number of steps.
"""
import numpy as np
import pytest
import torch

from oracle import device_sampler as ds
from oracle import relation_view as orv

pytestmark = pytest.mark.gpu
ROW_ATOL = 2e-6


@pytest.fixture(scope="module")
def T():
    from multike_b200 import _cabi, tables
    _cabi.load()
    return tables


def _small(golden):
    g = golden("ref_batch_relation.npz")
    n_ent = int(g["n_ent"])
    t1, t2 = g["triples1"], g["triples2"]
    return n_ent, t1, t2, np.concatenate([t1, g["sup1"]]), np.concatenate([t2, g["sup2"]])


def _view(T, golden, variant, dim=75, chunk=None, K=10, seed=5):
    from multike_b200.relation_view import RelationView
    n_ent, t1, t2, all1, all2 = _small(golden)
    gen = torch.Generator().manual_seed(11)
    ent0 = T.xavier_truncated_normal(2 * n_ent, dim, gen)
    rel0 = T.xavier_truncated_normal(5, dim, gen)
    return RelationView(2 * n_ent, 5, dim, t1, t2, n_ent, batch_size=200, neg_num=K, lr=0.001, seed=seed,
                        ent_init=ent0, rel_init=rel0, filter1=all1, filter2=all2, variant=variant,
                        persist_chunk=chunk)


@pytest.mark.parametrize("dim,K", [(75, 10), (128, 25), (32, 5), (64, 12), (100, 10)])
@pytest.mark.parametrize("chunk", [None, 4])
def test_persistent_epochs_equal_stepwise_epochs(T, golden, dim, K, chunk):
    out = []
    for variant in (3, 4):
        rv = _view(T, golden, variant, dim=dim, chunk=chunk if variant == 4 else None, K=K)
        losses = [rv.train_epoch(shuffle=False)[0] for _ in range(3)]
        trained = rv.train_steps(2, 9)  # starts inside an epoch and wraps past its end
        torch.cuda.synchronize()
        out.append((losses, trained, rv.step_losses.clone(), rv.ent.var.clone(), rv.rel.var.clone(),
                    rv.ent.grad.clone(), rv.ent.touched))
    a, b = out
    assert a[1] == b[1]
    assert b[0] == pytest.approx(a[0], rel=1e-6)
    torch.testing.assert_close(b[2], a[2], rtol=1e-6, atol=0)
    torch.testing.assert_close(b[3], a[3], rtol=0, atol=4 * ROW_ATOL)
    torch.testing.assert_close(b[4], a[4], rtol=0, atol=4 * ROW_ATOL)
    # the step leaves what phase 2 must leave: gradient table all zero, no row flagged
    assert float(b[5].abs().max()) == 0.0 and (b[6] is None or int(b[6].sum()) == 0)


def test_persistent_host_fed_steps(T, golden):
    ref = _view(T, golden, 3)
    want = [ref.train_epoch(shuffle=False)[0] for _ in range(2)]
    for chunk in (None, 4, 1):
        rv = _view(T, golden, 4, chunk=chunk)
        got = [rv.train_epoch(shuffle=False, host_fed=True)[0] for _ in range(2)]
        torch.cuda.synchronize()
        assert got == pytest.approx(want, rel=1e-6)
        # the kernel stores every step's loss into pinned host memory itself
        assert float(rv.host_losses.sum()) == pytest.approx(float(rv.step_losses.sum().item()), rel=1e-12)
        assert (rv.host_losses > 0).all()
        torch.testing.assert_close(rv.ent.var, ref.ent.var, rtol=0, atol=4 * ROW_ATOL)
        torch.testing.assert_close(rv.rel.var, ref.rel.var, rtol=0, atol=4 * ROW_ATOL)


def test_persistent_negatives_are_the_stepwise_sampler_draws(T, golden):
    """after a launch of n steps, neg buffer (n-1)&1 holds the draws of the last step: equal to
    mke_sample_structured at the same RNG coordinate and to the CPU restatement"""
    n_ent, t1, t2, all1, all2 = _small(golden)
    rv = _view(T, golden, 4)
    rv.train_steps(0, 4)
    torch.cuda.synchronize()
    (a1, b1), (a2, b2) = rv.step_slices(3)
    p1, p2 = t1[a1:b1], t2[a2:b2]
    n = len(p1) + len(p2)
    got_e = rv._neg[1][0][: n * rv.K].cpu().numpy().reshape(n, rv.K)
    got_s = rv._neg[1][1][:n].cpu().numpy().view(np.uint32)
    ne, ns = T.sample_structured(p1, rv.kg1, p2, rv.kg2, rv.K, 5, 3)
    assert np.array_equal(got_e, ne.cpu().numpy().reshape(n, rv.K))
    assert np.array_equal(got_s, ns.cpu().numpy().view(np.uint32))
    ok1 = ds.KG(entity_base=0, n_entities=n_ent, triples=all1)
    ok2 = ds.KG(entity_base=n_ent, n_entities=n_ent, triples=all2)
    neg = ds.sample_batch(p1, ok1, p2, ok2, rv.K, 5, 3)
    back = orv.structured_to_negatives(np.concatenate([p1, p2]), got_e, got_s, rv.K)
    assert np.array_equal(back, neg)


def test_persistent_tiny_grid_many_passes(T, golden, monkeypatch):
    """a 2-block grid makes every quarter walk several positives per step and every warp take many
    tickets of both phase-2 queues"""
    ref = _view(T, golden, 3)
    want = ref.train_epoch(shuffle=False)[0]
    monkeypatch.setenv("MKE_PERSIST_GRID", "2")
    rv = _view(T, golden, 4)
    got = rv.train_epoch(shuffle=False)[0]
    torch.cuda.synchronize()
    assert got == pytest.approx(want, rel=1e-6)
    torch.testing.assert_close(rv.ent.var, ref.ent.var, rtol=0, atol=2 * ROW_ATOL)
    torch.testing.assert_close(rv.rel.var, ref.rel.var, rtol=0, atol=2 * ROW_ATOL)


def test_persistent_trace_is_ordered(T, golden):
    rv = _view(T, golden, 4)
    rv.train_steps(0, 5)
    torch.cuda.synchronize()
    st = rv.persist_trace(5).cpu().numpy()
    assert st.shape == (12,) and (np.diff(st) >= 0).all() and st[-1] > st[0]


def test_persistent_full_size_steps_vs_stepwise(T):
    """BASELINE.json config 2 shape: 200 000 entities, d = 75, B = 20 000, K = 10, 6 steps"""
    from multike_b200 import synthetic
    from multike_b200.relation_view import RelationView
    kgs = synthetic.make_kgs(seed=1234)
    out = []
    for variant in (3, 4):
        gen = torch.Generator().manual_seed(20190754)
        rv = RelationView(kgs["n_ent"], kgs["n_rel"], 75, kgs["triples1"], kgs["triples2"], kgs["ent_split"],
                          batch_size=20000, neg_num=10, lr=0.001, seed=1234, variant=variant, generator=gen)
        trained = rv.train_steps(44, 6)  # wraps over the epoch end (46 steps per epoch)
        torch.cuda.synchronize()
        out.append((trained, rv.step_losses.clone(), rv.ent.var.clone(), rv.rel.var.clone()))
    a, b = out
    assert a[0] == b[0]
    torch.testing.assert_close(b[1], a[1], rtol=1e-6, atol=0)
    torch.testing.assert_close(b[2], a[2], rtol=0, atol=6 * ROW_ATOL)
    # relation rows sum up to 35 000 fp32 contributions per step in atomic order
    torch.testing.assert_close(b[3], a[3], rtol=0, atol=6e-5)
