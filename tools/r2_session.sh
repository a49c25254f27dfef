#!/bin/bash
# Round-2 GPU session: full GPU test suite, bench lines (default = persistent step kernel, and the stepwise
# driver beside it), ncu launch list + one full capture, compute-sanitizer over smoke() and the small tests.
# usage (GPU box, repo root): bash tools/r2_session.sh <out-dir-under-gpurun_out> [quick]
OUT=gpurun_out/${1:-r2}
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/gputests.log 2>&1; echo "gpu tests rc=$?"; tail -5 $OUT/gputests.log
timeout 300 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -c 600 $OUT/bench_default.json
timeout 300 python bench.py --variant 3 --no-cpu-baseline > $OUT/bench_v3.json 2> $OUT/bench_v3.err
timeout 300 python bench.py --workload synth1m_rel_d128_b20000_k25 --no-cpu-baseline > $OUT/bench_cfg5.json 2> $OUT/bench_cfg5.err
timeout 300 python bench.py --workload synth1m_rel_d128_b20000_k25 --variant 3 --no-cpu-baseline > $OUT/bench_cfg5_v3.json 2> $OUT/bench_cfg5_v3.err
python - <<PY
import json
for n in ("default", "v3", "cfg5", "cfg5_v3"):
    try:
        d = json.load(open("$OUT/bench_%s.json" % n))
        print(n, "value %.1f M/s e2e %.1f us/step %.1f frac %.3f" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"] * 1e3, d["roofline"]["frac"]), d["roofline"].get("phases", {}).get("phase1_us"), d["roofline"].get("phases", {}).get("phase2_us"))
    except Exception as e:
        print(n, "failed", e)
PY
if [ "$2" != "quick" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 92 --warmup 46 --skip-e2e --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rel_step_persist --launch-skip 1 -c 1 -o $OUT/ncu_full_persist python bench.py --steps 46 --warmup 46 --skip-e2e --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ncu -i $OUT/ncu_full_persist.ncu-rep --page raw --csv > $OUT/ncu_full_persist_raw.csv 2>/dev/null
ncu -i $OUT/ncu_full_persist.ncu-rep --page details > $OUT/ncu_full_persist_details.txt 2>/dev/null
MKE_PERSIST_GRID=4 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_persist.py -q -x -k "tiny or host_fed or negatives" > $OUT/sanitizer_memcheck_persist.log 2>&1; echo "memcheck persist rc=$?"
compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"
tail -3 $OUT/sanitizer_*.log
fi
