#!/bin/bash
out=gpurun_out/sim_$1
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
python tools/bench_sim.py > $out/sim.json 2> $out/sim.err; cat $out/sim.json; tail -3 $out/sim.err
ncu --set full --clock-control none -k regex:"row_topk" -c 1 -o $out/ncu_topk python tools/bench_sim.py --profile > $out/ncu.log 2>&1; tail -2 $out/ncu.log
