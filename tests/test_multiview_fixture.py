"""The real-data multi-view fixture (tests/golden/dbp_wd_100k_multiview.npz, digested through the
reference's own DataModel / PredicateAlignModel by tools/digest_dbp_wd_multiview.py) against the dataset
facts SURVEY.md section 8 records, and the stand-in literal vectors it is used with."""
import numpy as np
import pytest

from multike_b200 import synthetic


def test_literal_vectors_are_a_pure_function_of_the_literal_id():
    v = synthetic.literal_vectors(np.arange(2000), 75)
    assert v.shape == (2000, 75) and v.dtype == np.float32
    assert np.abs(np.linalg.norm(v.astype(np.float64), axis=1) - 1).max() < 1e-6
    again = synthetic.literal_vectors([1999, 5, 5, 0], 75)
    assert np.array_equal(again[0], v[1999]) and np.array_equal(again[1], v[5]) and np.array_equal(again[2], v[5])
    assert np.array_equal(again[3], v[0])
    # different literals are (nearly) orthogonal: |cos| of 75-d Gaussians stays below ~0.6
    c = v[:500].astype(np.float64) @ v[:500].astype(np.float64).T
    assert np.abs(c - np.eye(500)).max() < 0.6 and abs(float(v.mean())) < 1e-3
    # frozen values: the generator must not drift between rounds (the fixture stores ids, not vectors)
    f = synthetic.literal_vectors([0, 7, 942198], 75)
    assert f[0, :3].tolist() == pytest.approx([-0.04038414731621742, 0.11857521533966064, -0.053545426577329636], abs=1e-7)
    assert float(f[1, 11]) == pytest.approx(0.12988048791885376, abs=1e-7)
    assert float(f[2, 74]) == pytest.approx(-0.20715394616127014, abs=1e-7)


def test_multiview_fixture_matches_the_survey_numbers(golden):
    m, g = golden("dbp_wd_100k_multiview.npz"), golden("dbp_wd_100k_relation.npz")
    # cleaned attribute triples 621 595 + 989 153, 1 086 attributes, 942 199 literals
    assert len(m["attr1"]) == 621595 and len(m["attr2"]) == 989153 and int(m["attributes_num"]) == 1086
    assert int(m["n_literals"]) == 942199 and int(m["entities_num"]) == 200000 and int(m["relations_num"]) == 550
    # ids come from the same loader run as the relation fixture: KG1 entities 0..99 999, KG2 100 000..
    assert m["attr1"][:, 0].max() < 100000 <= m["attr2"][:, 0].min() and m["attr2"][:, 0].max() < 200000
    assert m["attr1"][:, 1].max() < m["attr2"][:, 1].min() and m["attr2"][:, 1].max() == 1085
    assert max(m["attr1"][:, 2].max(), m["attr2"][:, 2].max()) < len(m["value_literal"]) <= int(m["n_literals"])
    assert len(m["name_literal"]) == 200000 and m["name_literal"].max() < int(m["n_literals"])
    # 64 934 of the 100 000 gold links have identical cleaned local names on both sides (section 8c)
    links = np.concatenate([g["train_links"], g["valid_links"], g["test_links"]])
    same = int((m["name_literal"][links[:, 0]] == m["name_literal"][links[:, 1]]).sum())
    assert len(links) == 100000 and abs(same - 64934) <= 5
    # weights of predicate_alignment.add_weights: 0.2 for unmatched predicates, zoomed similarities otherwise
    for w in (m["attr1_w"], m["attr2_w"]):
        assert w.min() == pytest.approx(0.2) and w.max() == pytest.approx(1.0)
    # swapped attribute triples move a training link's attribute triples to the counterpart entity
    train1 = set(g["train_links"][:, 0].tolist())
    assert set(np.unique(m["sup_attr1"][:, 0]).tolist()) <= set(g["train_links"][:, 1].tolist())
    assert len(m["sup_attr1"]) == 186752 and len(m["sup_attr2"]) == 295985 and len(train1) == 30000
