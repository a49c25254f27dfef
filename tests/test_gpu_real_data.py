"""Real DBP-WD-100K relation triples: three epochs of the B200 path must reproduce the per-epoch
losses of the CPU oracle run recorded in profiles/r1_hits_oracle.json (identical inputs)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_three_epochs_on_dbp_wd_match_the_oracle_log(tmp_path):
    out = str(tmp_path / "hits.json")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "hits_experiment.py"), "--impl", "b200", "--epochs", "3",
                    "--out", out], check=True, cwd=ROOT, timeout=600, stdout=subprocess.DEVNULL)
    got = json.load(open(out))
    want = json.load(open(os.path.join(ROOT, "profiles", "r1_hits_oracle.json")))
    for g, w in zip(got["log"], want["log"][:3]):
        assert g["rel_loss"] == pytest.approx(w["rel_loss"], rel=2e-6)    # loss per positive, fp32 vs fp32
        assert g["ckge_loss"] == pytest.approx(w["ckge_loss"], rel=2e-6)
    assert 0.0 < got["hits@1"] < 100.0
    # Hits@k / MR / MRR of the 10 000 validation links: fused device evaluator == host ranking of the
    # exported rows (ranks may differ where two sims agree to fp32 rounding: <= 0.05 points)
    dev = got["device_evaluator"]
    for k in ("hits@1", "hits@5", "hits@10", "hits@50"):
        assert dev[k] == pytest.approx(got[k], abs=0.05)
    assert dev["mrr"] == pytest.approx(got["mrr"], rel=1e-3)
