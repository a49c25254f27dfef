#!/bin/bash
# GPU session for the persistent step kernel: parity tests, bench lines for a few knob settings, block trace.
# usage (on the GPU box, from the repo root): bash tools/persist_session.sh <out-dir-under-gpurun_out>
OUT=gpurun_out/${1:-persist}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_persist.py -x -q > $OUT/persist_tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/persist_tests.log
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 200 python bench.py --variant 4 --no-cpu-baseline > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$name.json"))
    ph = d["roofline"].get("phases", {})
    print("$name", "value %.1f M/s" % (d["value"] / 1e6), "e2e %.1f" % (d["e2e"]["value"] / 1e6), "us/step %.1f" % (d["ms_per_step"] * 1e3),
          "frac %.3f" % d["roofline"]["frac"], "p1 %.1f us p2 %.1f us" % (ph.get("phase1_us", 0), ph.get("phase2_us", 0)))
except Exception as e:
    print("$name failed", e); print(open("$OUT/bench_$name.err").read()[-1500:])
PY
}
timeout 200 python bench.py --variant 3 --no-cpu-baseline > $OUT/bench_v3.json 2> $OUT/bench_v3.err
python -c "
import json; d=json.load(open('$OUT/bench_v3.json')); print('v3 value %.1f e2e %.1f us/step %.1f frac %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']*1e3, d['roofline']['frac']))"
run sp1 MKE_PERSIST_SAMP_PHASE=1
run sp0 MKE_PERSIST_SAMP_PHASE=0
run sp1b MKE_PERSIST_SAMP_PHASE=1
