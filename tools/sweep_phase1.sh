#!/bin/bash
# GPU session script (run under gpurun): parity tests, phase-1 schedule sweep, sim kernel timings.
out=gpurun_out/sweep_$1
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest.log
tail -15 $out/pytest.log
b() { name=$1; shift; ( "$@" ) > $out/$name.json 2> $out/$name.err; python - $out/$name.json $name <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value %.1f M/s step %.1f us  p1 %.2f us frac %.3f e2e %.1f" % (j["value"]/1e6, j["ms_per_step"]*1e3, j["roofline"]["launch_ms"]*1e3, j["roofline"]["frac"], j["e2e"]["value"]/1e6))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
B="python bench.py --no-cpu-baseline --steps 460 --warmup 46"
b v0 $B --variant 0
b v3 $B --variant 3
python tools/bench_sim.py > $out/sim.json 2> $out/sim.err; cat $out/sim.json; tail -3 $out/sim.err
ncu --set full --clock-control none --import-source on -k regex:"sim_tile" -c 2 -o $out/ncu_sim python tools/bench_sim.py --profile > $out/ncu_sim.log 2>&1
ls $out
