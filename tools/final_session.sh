#!/bin/bash
# Round-end style session on one B200: GPU tests, smoke, reference arm, default bench line, launch list.
out=gpurun_out/final_$1
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
python __graft_entry__.py --smoke > $out/smoke.log 2>&1; tail -1 $out/smoke.log
python bench.py --impl reference --steps 6 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; cut -c1-300 $out/bench_reference.json
python bench.py > $out/bench_default.json 2> $out/bench_default.err; cut -c1-2500 $out/bench_default.json
for e in 1 4 16; do
  python bench.py --no-cpu-baseline --p1-every $e > $out/bench_every$e.json 2>/dev/null
  python - $out/bench_every$e.json $e <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("every", sys.argv[2], "value %.1f e2e %.1f step %.2f us p1 %.2f us (%d timed) floor %.2f us"%(j["value"]/1e6,j["e2e"]["value"]/1e6,j["ms_per_step"]*1e3,j["roofline"]["launch_ms"]*1e3,j["roofline"]["timed_launches"],j["roofline"]["event_pair_floor_ms"]*1e3))
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 40 --warmup 6 --no-cpu-baseline > $out/launches_bench.log 2>&1
tail -2 $out/launches.csv
