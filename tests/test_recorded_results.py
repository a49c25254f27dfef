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
