"""GPU parity of csrc/mke_sim.cu through the C-ABI: the gold-rank evaluator (mke_sim_rank) and the
truncated-epsilon neighbour search (mke_sim_topk) against oracle/alignment.py, against the outputs
of the reference's own base/alignment.py / base/batch.py (tests/golden/ref_sim.npz) and, at
BASELINE sizes, against a float64 torch computation on the same device.

Tolerances: ranks, arg-max columns and neighbour lists are INDEX work -> exact wherever the order
is decided by more than SIM_TOL = 2e-6 (fp32 rounding of a 75..128-term inner product of unit
rows is ~1e-7); exact ties (bit-equal sims) follow the stable rule and are compared exactly."""
import numpy as np
import pytest
import torch

from oracle import alignment as oa

pytestmark = pytest.mark.gpu
SIM_TOL = 2e-6


@pytest.fixture(scope="module")
def S():
    from multike_b200 import _cabi, similarity
    _cabi.load()
    return similarity


@pytest.fixture(autouse=True, params=["tcgen05", "fma"])
def impl(request):
    """every test runs on both implementations: the tensor-core tiles (default, csrc/mke_sim_tc.cu) and the fp32
    FMA tiles kept as the baseline (csrc/mke_sim.cu)"""
    from multike_b200 import _cabi
    lib = _cabi.load()
    prev = lib.mke_sim_use_tensor_cores(1 if request.param == "tcgen05" else 0)
    yield request.param
    lib.mke_sim_use_tensor_cores(prev)


@pytest.fixture(scope="module")
def ref(golden):
    return golden("ref_sim.npz")


def metrics(rank, top_k):
    r = rank.astype(np.float64) + 1
    return [int((rank < k).sum()) for k in top_k], float(r.mean()), float((1 / r).mean())


def test_rank_matches_reference_outputs(S, ref):
    top_k = ref["top_k"].tolist()
    for c in range(len(ref["align_cases"])):
        a, b = ref["align%d_a" % c], ref["align%d_b" % c]
        rank, top1 = S.sim_rank(a, b, normalize=True)
        hits, mr, mrr = metrics(rank.cpu().numpy(), top_k)
        assert hits == ref["align%d_hits" % c].tolist()
        assert mr == pytest.approx(float(ref["align%d_mr" % c]), rel=1e-12)
        assert mrr == pytest.approx(float(ref["align%d_mrr" % c]), rel=1e-12)
        assert np.array_equal(top1.cpu().numpy(), ref["align%d_top1" % c])


def test_refapi_greedy_alignment_and_evaluation(S, ref, capsys):
    from multike_b200.refapi.base import alignment as ali, evaluation as eva
    top_k = ref["top_k"].tolist()
    a, b = ref["align1_a"], ref["align1_b"]
    rest, hits1, mr, mrr = ali.greedy_alignment(a, b, top_k, 8, 'inner', True, 0, True)
    assert "accurate results: hits@[1, 5, 10, 50]" in capsys.readouterr().out
    assert hits1 == pytest.approx(float(ref["align1_hits1"]))
    assert mr == pytest.approx(float(ref["align1_mr"]), rel=1e-12) and mrr == pytest.approx(float(ref["align1_mrr"]), rel=1e-12)
    assert rest == set(zip(range(len(a)), ref["align1_top1"].tolist()))
    h1, m = eva.valid(a, b, None, top_k, 8, normalize=True)
    assert h1 == hits1 and m == mrr
    # mapping: embeds1 @ M first (base/evaluation.py:11)
    M = np.linalg.qr(np.random.default_rng(0).standard_normal((75, 75)))[0].astype(np.float32)
    _, h1m, _ = eva.test(a @ M.T, b, M, top_k, 8, normalize=True)
    assert h1m == pytest.approx(hits1, abs=0.4)
    with pytest.raises(NotImplementedError):
        ali.greedy_alignment(a, b, top_k, 8, 'inner', True, 10, True)


@pytest.mark.parametrize("n1,n2,d", [(1, 1, 3), (129, 127, 75), (300, 1000, 128), (515, 260, 8)])
def test_rank_exact_ties_integer_embeddings(S, n1, n2, d):
    """small-integer rows: every sim is exact in fp32, ties are everywhere -> the stable tie rule
    (equal sims rank by ascending column) is checked bit-exactly, incl. duplicate rows."""
    rng = np.random.default_rng(n1 * 7 + n2)
    a = rng.integers(-2, 3, (n1, d)).astype(np.float32)
    b = rng.integers(-2, 3, (n2, d)).astype(np.float32)
    b[rng.integers(0, n2, n2 // 3)] = b[rng.integers(0, n2, n2 // 3)]  # duplicate candidates
    gold = rng.integers(0, n2, n1).astype(np.int32)
    rank, top1 = S.sim_rank(a, b, gold=gold, normalize=False)
    want_rank, want_top1 = oa.gold_ranks(oa.sim(a, b, dtype=np.float64), gold)
    assert np.array_equal(rank.cpu().numpy(), want_rank)
    assert np.array_equal(top1.cpu().numpy(), want_top1)


def test_rank_gathers_rows_of_a_padded_table(S):
    """idx1/idx2 gather rows of one [rows, stride] table (what MultiKE_Late.valid does with
    ent_embeds[valid_entities1]); columns >= dim are ignored even when they hold garbage."""
    rng = np.random.default_rng(5)
    table = rng.standard_normal((900, 80)).astype(np.float32)
    table[:, 75:] = 1e6
    i1 = rng.permutation(900)[:200].astype(np.int32)
    i2 = np.concatenate([i1[::-1], rng.permutation(900)[:333].astype(np.int32)])
    gold = np.arange(200)[::-1].astype(np.int32).copy()  # row i of emb1 is the same entity as column 199 - i
    rank, top1 = S.sim_rank(torch.from_numpy(table).cuda(), torch.from_numpy(table).cuda(), idx1=i1, idx2=i2,
                            gold=gold, normalize=True, dim=75)
    s = oa.sim(table[i1, :75], table[i2, :75], normalize=True, dtype=np.float64)
    want_rank, want_top1 = oa.gold_ranks(s, gold)
    # an entity is most similar to itself: rank 0 unless the permutation drew it twice (then the
    # earlier duplicate column wins the tie)
    assert np.array_equal(rank.cpu().numpy(), want_rank) and int(rank.max()) <= 1
    assert np.array_equal(top1.cpu().numpy(), want_top1)


def rank_bounds(a, b, gold, chunk=2048):
    """float64 checker on the GPU: [lo, hi] of the admissible rank under SIM_TOL, and the exact
    float64 rank"""
    an = a.double() / a.double().norm(dim=1, keepdim=True).clamp_min(1e-300)
    bn = b.double() / b.double().norm(dim=1, keepdim=True).clamp_min(1e-300)
    lo, hi, ex = [], [], []
    for r0 in range(0, an.shape[0], chunk):
        s = an[r0:r0 + chunk] @ bn.T
        g = gold[r0:r0 + chunk].long()
        sg = s.gather(1, g[:, None])
        lo.append((s > sg + SIM_TOL).sum(1))
        hi.append((s >= sg - SIM_TOL).sum(1) - 1)
        ex.append((s > sg).sum(1))
    return torch.cat(lo), torch.cat(hi), torch.cat(ex)


@pytest.mark.parametrize("n1,n2,d", [(2000, 3333, 75), (10000, 70000, 75), (20000, 20000, 128)])
def test_rank_full_size_vs_float64(S, n1, n2, d):
    """BASELINE evaluation shapes (valid(): 10 000 x 70 000 at dim 75): every rank lies inside the
    float64 bounds, almost all equal the float64 rank, and Hits@1/10 / MRR agree."""
    gen = torch.Generator(device="cuda").manual_seed(n1 + n2)
    a = torch.randn(n1, d, device="cuda", generator=gen)
    b = torch.randn(n2, d, device="cuda", generator=gen)
    m = min(n1, n2)
    noise = torch.rand(m, 1, device="cuda", generator=gen) * 3.0
    b[:m] = a[:m] + noise * torch.randn(m, d, device="cuda", generator=gen)
    gold = torch.arange(n1, device="cuda", dtype=torch.int32) % n2
    rank, top1 = S.sim_rank(a, b, gold=gold, normalize=True)
    lo, hi, ex = rank_bounds(a, b, gold)
    r = rank.long()
    assert bool(((r >= lo) & (r <= hi)).all())
    assert float((r == ex).float().mean()) > 0.999
    for k in (1, 10):
        assert abs(float((r < k).float().mean()) - float((ex < k).float().mean())) * 100 < 0.05
    assert float((1.0 / (r + 1).double()).mean()) == pytest.approx(float((1.0 / (ex + 1).double()).mean()), rel=1e-4)
    # top1 really is a maximum of its row (float64, within tolerance)
    an = a.double() / a.double().norm(dim=1, keepdim=True)
    bn = b.double() / b.double().norm(dim=1, keepdim=True)
    pick = torch.arange(0, n1, max(1, n1 // 512), device="cuda")
    s = an[pick] @ bn.T
    assert bool((s.gather(1, top1[pick].long()[:, None])[:, 0] >= s.max(1).values - SIM_TOL).all())


def test_topk_matches_reference_outputs(S, ref):
    for c, (n, d, k) in enumerate(ref["nb_cases"].tolist()):
        e, ids = ref["nb%d_e" % c], ref["nb%d_ids" % c].astype(np.int32)
        got = S.sim_topk(e, k, id_list=ids).cpu().numpy()
        assert np.array_equal(np.sort(got, axis=1), ref["nb%d_lists" % c])
        # list order = ascending column, i.e. the oracle's stable top-k
        assert np.array_equal(got, oa.find_neighbours(ids, e, k))


@pytest.mark.parametrize("n,d,k,chunk", [(1, 4, 1, 128), (700, 16, 33, 128), (1000, 75, 999, 256), (515, 128, 64, 8192)])
def test_topk_exact_ties_integer_embeddings(S, n, d, k, chunk):
    """exact sims with masses of ties at the k-th value; several passes over the rows (small
    workspace); scattered output rows and an id base."""
    rng = np.random.default_rng(n + k)
    e = rng.integers(-1, 2, (n, d)).astype(np.float32)
    e[rng.integers(0, n, n // 4)] = e[rng.integers(0, n, n // 4)]
    want = oa.find_neighbours(np.arange(n) + 50, e, k, dtype=np.float64)
    got = S.sim_topk(e, k, id_base=50, chunk_rows=chunk).cpu().numpy()
    assert np.array_equal(got, want)
    rows = rng.permutation(3 * n)[:n].astype(np.int32)
    out = torch.full((3 * n, k), -1, dtype=torch.int32, device="cuda")
    S.sim_topk(e, k, id_base=50, out=out, out_rows=rows, chunk_rows=chunk)
    o = out.cpu().numpy()
    assert np.array_equal(o[rows], want)
    untouched = np.ones(3 * n, bool)
    untouched[rows] = False
    assert (o[untouched] == -1).all()


def test_topk_full_size_vs_float64(S):
    """DBP-WD shape: one KG's 100 000 entities at dim 75, k = int(0.02 * 100 000) = 2 000
    (MultiKE_CSL.py:91-92).  Lists are strictly ascending (so duplicate-free); for a sample of
    rows the chosen set beats every column left out, within SIM_TOL, in float64."""
    n, d, k = 100000, 75, 2000
    gen = torch.Generator(device="cuda").manual_seed(3)
    e = torch.randn(n, d, device="cuda", generator=gen)
    e = e / e.norm(dim=1, keepdim=True)
    nb = S.sim_topk(e, k, chunk_rows=8192)
    torch.cuda.synchronize()
    assert bool((nb[:, 1:] > nb[:, :-1]).all()) and int(nb.min()) >= 0 and int(nb.max()) < n
    pick = torch.arange(0, n, 997, device="cuda")
    s = e[pick].double() @ e.double().T
    chosen = torch.zeros_like(s, dtype=torch.bool)
    chosen.scatter_(1, nb[pick].long(), True)
    worst_in = torch.where(chosen, s, torch.full_like(s, 9.0)).min(1).values
    best_out = torch.where(chosen, torch.full_like(s, -9.0), s).max(1).values
    assert bool((worst_in >= best_out - SIM_TOL).all())
    assert bool(chosen[torch.arange(len(pick)), pick].all())  # every entity is its own nearest neighbour


def test_topk_fused_candidate_search_equals_the_exact_path(S, impl):
    """the search that never writes the similarity matrix (sampled threshold -> candidate lists from the tile epilogue ->
    select per row; rows whose list came out short redone exactly) returns the lists of the path that materialises the
    sims, on the same tiles: bit-equal sims, so the lists are equal element by element -- with heavy ties (integer
    rows, duplicates) and with continuous rows"""
    if impl != "tcgen05":
        pytest.skip("tensor-core path only")
    from multike_b200 import _cabi
    lib = _cabi.load()
    gen = torch.Generator(device="cuda").manual_seed(9)
    cases = [(torch.randn(20000, 32, device="cuda", generator=gen), 300),
             (torch.randint(-1, 2, (17000, 24), device="cuda", generator=gen).float(), 500)]
    cases[1][0][:4000] = cases[1][0][4000:8000]       # duplicate rows: ties at every rank
    skew = torch.randn(16500, 75, device="cuda", generator=gen)
    skew[:300] = skew[0] + 0.01 * torch.randn(300, 75, device="cuda", generator=gen)   # a tight cluster: rows whose
    cases.append((skew, 250))                                                          # sampled threshold misses
    for e, k in cases:
        lib.mke_sim_use_tensor_cores(1)
        fused = S.sim_topk(e, k, normalize=True, chunk_rows=2048)
        lib.mke_sim_use_tensor_cores(2)
        exact = S.sim_topk(e, k, normalize=True, chunk_rows=2048)
        lib.mke_sim_use_tensor_cores(1)
        assert torch.equal(fused, exact)


def test_generate_neighbours_feeds_the_device_sampler(S):
    """refapi.base.batch.generate_neighbours -> NeighbourTable -> KGSampler: negatives of the
    truncated sampler come from the anchor's list (base/batch.py:93-94)."""
    from multike_b200 import tables as T
    from multike_b200.refapi.base import batch as bat
    rng = np.random.default_rng(1)
    n, k = 600, 12
    e = rng.standard_normal((n, 75)).astype(np.float32)
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    ids = np.arange(n, dtype=np.int32)
    nb = bat.generate_neighbours(e, ids.tolist(), k, 4)
    assert len(nb) == n and nb.get(5) == oa.find_neighbours(ids, e, k)[5].tolist()
    trip = np.unique(np.stack([rng.integers(0, n, 4000), rng.integers(0, 5, 4000), rng.integers(0, n, 4000)], 1),
                     axis=0).astype(np.int32)
    kg = T.KGSampler(entity_base=0, n_entities=n, triple_set=T.TripleSet(trip), neighbours=nb.matrix)
    neg = T.sample_uniform(trip[:512], kg, None, None, 10, 3, 0).cpu().numpy().reshape(512, 10, 3)
    lists = nb.matrix.cpu().numpy()
    for i in range(512):
        h, _, t = trip[i]
        head_side = neg[i, 0, 0] != h or (neg[i, :, 2] == t).all() and (neg[i, :, 0] != h).any()
        for j in range(10):
            if neg[i, j, 2] == t and neg[i, j, 0] != h:
                assert neg[i, j, 0] in lists[h]
            elif neg[i, j, 0] == h and neg[i, j, 2] != t:
                assert neg[i, j, 2] in lists[t]
