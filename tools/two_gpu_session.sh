#!/bin/bash
# 2-GPU session: full GPU test suite (incl. the world-2 multi-GPU tests), sim timings, sharded bench.
out=gpurun_out/two_$1
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest.log
python tools/bench_sim.py > $out/sim.json 2> $out/sim.err; cat $out/sim.json; tail -3 $out/sim.err
bench() {
  name=$1; n=$2; shift 2
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 200 --warmup 20 > $out/$name.json 2> $out/$name.err
  python - $out/$name.json $name <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value %.1f M/s step %.1f us  p1 %.2f us" % (j["value"]/1e6, j["ms_per_step"]*1e3, j["roofline"]["launch_ms"]*1e3))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
bench n2_default 2 X=1
bench n2_ahead 2 MKE_DRAW_AHEAD=1
MKE_DRAW_AHEAD=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 tests/multi_gpu_check.py > $out/check_ahead.log 2>&1; grep MULTI_GPU_CHECK $out/check_ahead.log || tail -5 $out/check_ahead.log
python bench.py --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err; python - $out/bench_n1.json <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("n1 value %.1f e2e %.1f p1 %.2f us floor %.2f us"%(j["value"]/1e6,j["e2e"]["value"]/1e6,j["roofline"]["launch_ms"]*1e3,j["roofline"]["event_pair_floor_ms"]*1e3))
PY
