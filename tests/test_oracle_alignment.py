"""oracle/alignment.py against the reference's own base/alignment.py and base/batch.py outputs
(tests/golden/ref_sim.npz, generated in the build container by tests/golden/make_golden_sim.py)."""
import numpy as np
import pytest

from oracle import alignment as oa


@pytest.fixture(scope="module")
def ref(golden):
    return golden("ref_sim.npz")


def test_greedy_alignment_matches_reference(ref):
    top_k = ref["top_k"].tolist()
    for c in range(len(ref["align_cases"])):
        a, b = ref["align%d_a" % c], ref["align%d_b" % c]
        rest, hits, mr, mrr = oa.greedy_alignment(a, b, top_k, normalize=True)
        assert hits[0] == pytest.approx(float(ref["align%d_hits1" % c]))
        assert np.array_equal(np.round(ref["align%d_hits" % c] / len(a) * 100, 3), hits)
        assert mr == pytest.approx(float(ref["align%d_mr" % c]), rel=1e-12)
        assert mrr == pytest.approx(float(ref["align%d_mrr" % c]), rel=1e-12)
        top1 = np.array([j for _, j in sorted(rest)])
        assert np.array_equal(top1, ref["align%d_top1" % c])


def test_find_neighbours_matches_reference(ref):
    for c, (n, d, k) in enumerate(ref["nb_cases"].tolist()):
        got = oa.find_neighbours(ref["nb%d_ids" % c], ref["nb%d_e" % c], k)
        want = ref["nb%d_lists" % c]
        # the reference's list order is argpartition's; compare as sets (both sorted by id here)
        assert np.array_equal(np.sort(got, axis=1), want)


def test_stable_tie_rule():
    s = np.array([[0.5, 0.9, 0.5, 0.5], [0.1, 0.1, 0.1, 0.1]], dtype=np.float32)
    rank, top1 = oa.gold_ranks(s, gold=np.array([2, 3]))
    assert rank.tolist() == [2, 3] and top1.tolist() == [1, 0]
    rank, _ = oa.gold_ranks(s, gold=np.array([0, 0]))
    assert rank.tolist() == [1, 0]


def test_host_side_metrics_from_ranks_match_calculate_rank(ref):
    """refapi.base.alignment.rank_metrics (the host arithmetic after mke_sim_rank) == the sums of
    calculate_rank (base/alignment.py:141-163) on the reference's own cases"""
    import torch
    from multike_b200.refapi.base.alignment import rank_metrics
    top_k = ref["top_k"].tolist()
    for c in range(len(ref["align_cases"])):
        a, b = ref["align%d_a" % c], ref["align%d_b" % c]
        rank, _ = oa.gold_ranks(oa.sim(a, b, normalize=True))
        mr, mrr, hits = rank_metrics(torch.from_numpy(rank), top_k)
        assert hits == [float(x) for x in ref["align%d_hits" % c]]
        assert mr == pytest.approx(float(ref["align%d_mr" % c]), rel=1e-12)
        assert mrr == pytest.approx(float(ref["align%d_mrr" % c]), rel=1e-12)


def test_evaluation_entry_points_forward_to_greedy_alignment(monkeypatch):
    """refapi.base.evaluation.valid / test (base/evaluation.py:6-28): argument order, defaults
    (valid is the quick ranking, test the accurate one) and return shapes -- with the ranker
    replaced by a recorder, so no device is needed."""
    from multike_b200.refapi.base import evaluation as eva
    calls = []

    def fake(embed1, embed2, top_k, nums_threads, metric, normalize, csls_k, accurate):
        calls.append((embed1, embed2, top_k, nums_threads, metric, normalize, csls_k, accurate))
        return {(0, 1)}, 12.5, 3.0, 0.25

    monkeypatch.setattr(eva, "greedy_alignment", fake)
    assert eva.valid("A", "B", None, [1, 5], 8, normalize=True) == (12.5, 0.25)
    assert calls[-1] == ("A", "B", [1, 5], 8, "inner", True, 0, False)
    assert eva.test("A", "B", None, [1], 4) == ({(0, 1)}, 12.5, 0.25)
    assert calls[-1] == ("A", "B", [1], 4, "inner", False, 0, True)
    assert eva.early_stop(0.5, 0.4, 0.3) == (0.4, 0.3, True) and eva.early_stop(0.3, 0.4, 0.5) == (0.4, 0.5, False)


def test_id_dict_files_have_the_reference_format(tmp_path, capsys):
    """MultiKE.save's id dictionaries (utils.py:60-67): "key<TAB>id" lines; a missing dict writes no file"""
    from multike_b200.refapi.MultiKE_model import write_id_dict
    path = str(tmp_path / "kg1_ent_ids")
    write_id_dict(path, {"http://dbpedia.org/resource/A": 0, "http://dbpedia.org/resource/B": 7})
    assert open(path, encoding="utf8").read() == "http://dbpedia.org/resource/A\t0\nhttp://dbpedia.org/resource/B\t7\n"
    assert "saved." in capsys.readouterr().out
    write_id_dict(str(tmp_path / "none"), None)
    assert not (tmp_path / "none").exists()


def test_literal_token_matrix_matches_reference_loop():
    """refapi.literal_encoder.literal_token_matrix == the per-literal loop of literal_encoder.py:166-172"""
    from multike_b200.refapi.literal_encoder import generate_unlisted_word2vec, literal_token_matrix
    rng = np.random.default_rng(0)
    w2v = {w: rng.standard_normal(6).astype(np.float32) for w in ("new", "york", "city", "of", "x")}
    lits = ["new york", "city of new york state capital", "unknown words", "x", ""]
    got = literal_token_matrix(lits, w2v, tokens_max_len=5, word2vec_dimension=6)
    want = []
    for literal in lits:
        vectors = np.zeros((5, 6), dtype=np.float32)
        words = literal.split(' ')
        for i in range(min(5, len(words))):
            if words[i] in w2v:
                vectors[i] = w2v[words[i]]
        want.append(vectors)
    assert np.array_equal(got, np.stack(want))
    assert sorted(generate_unlisted_word2vec(dict(w2v), ["new york", "x"])) == sorted(w2v)   # nothing missing: no gensim needed
