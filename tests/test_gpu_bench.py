"""bench.py on the GPU: the driver's invocation in small (`--steps 6 --warmup 3`), every key the driver reads checked
on the live JSON line, and the line's own arithmetic (value x time = positives, roofline fraction = achieved / peak,
achieved = algorithmic bytes / launch time)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"}


def test_b200_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "6", "--warmup", "3",
                          "--no-cpu-baseline"], cwd=ROOT, check=True, timeout=600, stdout=subprocess.PIPE, text=True).stdout
    lines = [l for l in out.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert BASE_KEYS <= set(j) and "impl" not in j
    assert j["metric"] == "pos_triples_per_sec_rel_view_train_step" and j["unit"] == "triples/s"
    assert j["n_gpus"] == 1 and j["steps"] == 6 and j["warmup"] >= 3 and j["dtype"] == "f32" and j["data"] == "synthetic"
    assert j["config"]["workload"] == "dwy100k_rel_d75_b20000_k10" and "model" not in j["config"]
    assert j["higher_is_better"] is True and j["vs_baseline"] is None and j["scaling"] == "weak"
    r = j["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.2 < r["frac"] < 1.0, r      # a plausible fraction of the measured HBM peak (not a CPU path, not > peak)
    assert abs(r["achieved"] - r["algorithmic_bytes"] / (r["launch_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    e = j["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != j["value"]
    assert j["gpu_launches"] >= 1
    assert not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # value = positives / time: 6 steps of 20 000 positives (the epoch's first steps are full batches)
    assert abs(j["value"] * j["ms_per_step"] * 1e-3 - 20000) < 1.0
