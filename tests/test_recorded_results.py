"""The committed accuracy record: B200 path and CPU oracle on identical DBP-WD inputs
(profiles/r1_hits_*.json, produced by tools/hits_experiment.py) agree within the BASELINE tolerance."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hits_within_half_a_point_of_the_oracle():
    o = json.load(open(os.path.join(ROOT, "profiles", "r1_hits_oracle.json")))
    b = json.load(open(os.path.join(ROOT, "profiles", "r1_hits_b200.json")))
    assert o["epochs"] == b["epochs"] and o["batch"] == b["batch"] and o["neg"] == b["neg"]
    for k in ("hits@1", "hits@5", "hits@10", "hits@50"):
        assert abs(o[k] - b[k]) <= 0.5, (k, o[k], b[k])
    for x, y in zip(o["log"], b["log"]):
        assert y["rel_loss"] == pytest.approx(x["rel_loss"], rel=1e-5)
        assert y["ckge_loss"] == pytest.approx(x["ckge_loss"], rel=1e-5)


def test_multiview_oracle_record_is_complete():
    """profiles/r1_multiview_oracle.json (tools/multiview_experiment.py --impl oracle): the target of the B200 arm"""
    o = json.load(open(os.path.join(ROOT, "profiles", "r1_multiview_oracle.json")))
    assert o["impl"] == "oracle" and o["epochs"] == len(o["log"]) >= 3 and o["candidates"] == 70000
    rel = json.load(open(os.path.join(ROOT, "profiles", "r1_hits_oracle.json")))
    # same relation-view inputs and negatives as the relation-only record: epoch 1 agrees
    assert o["log"][0]["rel_loss"] == pytest.approx(rel["log"][0]["rel_loss"], rel=1e-6)
    assert o["log"][0]["ckge_rel_loss"] == pytest.approx(rel["log"][0]["ckge_loss"], rel=1e-6)
    for a, b in zip(o["log"], o["log"][1:]):
        assert b["common_loss"] < a["common_loss"] and b["attr_loss"] < a["attr_loss"]
    assert o["views"]["nv"]["hits@1"] == pytest.approx(64.83, abs=0.2)      # links with identical names
    assert set(o["views"]) == {"nv", "rv", "av", "final"}


def test_multiview_ssl_oracle_record_is_complete():
    o = json.load(open(os.path.join(ROOT, "profiles", "r1_multiview_ssl_oracle.json")))
    assert o["impl"] == "oracle" and o["mode"] == "ssl" and set(o["views"]) == {"nv", "rv", "av", "avg", "final"}
    views = [r for r in o["log"] if "epoch" in r]
    maps = [r for r in o["log"] if "shared_epoch" in r]
    assert len(views) == o["epochs"] and len(maps) == o["shared_epochs"] and all(r["common_loss"] == 0.0 for r in views)
    assert all(b["mapping_loss"] < a["mapping_loss"] for a, b in zip(maps, maps[1:]))
    itc = json.load(open(os.path.join(ROOT, "profiles", "r1_multiview_oracle.json")))
    # the first epoch does not depend on the schedule (the common-space step comes last)
    for k in ("rel_loss", "ckge_rel_loss", "attr_loss", "ckge_attr_loss"):
        assert views[0][k] == pytest.approx(itc["log"][0][k], rel=1e-6)
