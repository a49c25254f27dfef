#!/bin/bash
# Round-2, third GPU session: full GPU test suite, driver-style and default bench lines, ncu captures of the tensor-core
# similarity kernels, compute-sanitizer over the new kernels (similarity on tcgen05, row staging).
# usage (GPU box, repo root): bash tools/r2_session3.sh <out-dir-under-gpurun_out>
OUT=gpurun_out/${1:-r2s3}
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 1800 python -m pytest tests -m gpu -q -x > $OUT/gputests.log 2>&1; echo "gpu tests rc=$?"; tail -4 $OUT/gputests.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_driver_style.json 2> $OUT/bench_driver_style.err; tail -c 300 $OUT/bench_driver_style.json
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; tail -c 200 $OUT/bench_reference.json
timeout 300 python bench.py --no-cpu-baseline > $OUT/bench_default.json 2> $OUT/bench_default.err
python - <<PY
import json
for n in ("driver_style", "default"):
    try:
        d = json.load(open("$OUT/bench_%s.json" % n))
        print(n, "value %.1f M/s e2e %.1f us/step %.1f frac %.3f" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"] * 1e3, d["roofline"]["frac"]))
    except Exception as e:
        print(n, "failed", e)
PY
ncu --set full --clock-control none --import-source on -k regex:sim_tc_kernel -c 3 -o $OUT/ncu_full_sim_tc python tools/bench_sim.py --profile > $OUT/ncu_sim.log 2>&1
ncu -i $OUT/ncu_full_sim_tc.ncu-rep --page raw --csv > $OUT/ncu_full_sim_tc_raw.csv 2>/dev/null
ncu -i $OUT/ncu_full_sim_tc.ncu-rep --page details > $OUT/ncu_full_sim_tc_details.txt 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_sim.csv python tools/bench_sim.py --profile > $OUT/ncu_sim_launches.log 2>&1
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sim.py -q -x -k "ties or reference_outputs or fused or gathers" > $OUT/sanitizer_memcheck_sim.log 2>&1; echo "memcheck sim rc=$?"
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_sim.py -q -x -k "ties and tcgen05" > $OUT/sanitizer_racecheck_sim.log 2>&1; echo "racecheck sim rc=$?"
compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
tail -2 $OUT/sanitizer_*.log
rm -f $OUT/ncu_full_sim_tc.ncu-rep.tmp
