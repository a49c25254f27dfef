"""bench.py's JSON contract, checked without a GPU: the reference arm (`--impl reference`, the oracle
port of the reference's CPU path) is run for one bounded step.  The B200 arm's line is checked live by
tests/test_gpu_bench.py."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], cwd=ROOT, check=True, timeout=600, stdout=subprocess.PIPE, text=True).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert BASE_KEYS <= set(j) and j["impl"] == "reference"
    assert j["metric"] == "pos_triples_per_sec_rel_view_train_step" and j["unit"] == "triples/s"
    assert j["higher_is_better"] is True and j["vs_baseline"] is None and j["value"] > 0
    assert j["config"]["workload"] == "dwy100k_rel_d75_b20000_k10"
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["sample"]
    assert j["cpu_baseline"]["value"] == j["value"] == j["e2e"]["value"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
