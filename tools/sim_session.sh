#!/bin/bash
out=gpurun_out/sim_$1
mkdir -p $out
python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest.log
python tools/bench_sim.py > $out/sim.json 2> $out/sim.err; cat $out/sim.json; tail -3 $out/sim.err
