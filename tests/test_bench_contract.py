"""bench.py's JSON contract, checked without a GPU: the reference arm (`--impl reference`, the oracle
port of the reference's CPU path) is run for one bounded step, and the committed line of the B200
arm (profiles/r1_bench_default.json, written by `python bench.py` on a B200) is checked for every
key the driver reads."""
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


def test_recorded_b200_line_has_the_contract_keys():
    j = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_default.json")))
    assert BASE_KEYS | {"clocks", "roofline"} <= set(j) and "impl" not in j
    assert j["n_gpus"] == 1 and j["warmup"] >= 3 and j["dtype"] == "f32" and j["data"] == "synthetic"
    assert j["config"]["workload"] == "dwy100k_rel_d75_b20000_k10" and "model" not in j["config"]
    r = j["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    # achieved = algorithmic bytes per launch / mean launch duration
    assert abs(r["achieved"] - r["algorithmic_bytes"] / (r["launch_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    e = j["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != j["value"]
    assert j["gpu_launches"] >= 2 * j["steps"]
    assert not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    c = j["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["unit"] == j["unit"] and c["cores"] >= 1
    # value = positives / time: ms_per_step x value = positives per step (the 46-step epoch's mean batch)
    assert abs(j["value"] * j["ms_per_step"] * 1e-3 - r["algorithmic_bytes"] / r["bytes_per_positive"]) < 1.0
