#!/bin/bash
# Multi-GPU session (run under `gpurun --gpus N`): correctness of the row-sharded relation view under
# "negatives where they live" with both phase-1 schedules, then bench lines against the old scheme.
out=gpurun_out/multi_$1
mkdir -p $out
N=${2:-4}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run $N 29601 tests/multi_gpu_check.py > $out/check_n${N}.log 2>&1; grep MULTI_GPU_CHECK $out/check_n${N}.log || tail -5 $out/check_n${N}.log
MKE_BY_KG=0 run $N 29602 tests/multi_gpu_check.py > $out/check_n${N}_mod.log 2>&1; grep MULTI_GPU_CHECK $out/check_n${N}_mod.log || tail -5 $out/check_n${N}_mod.log
run 2 29603 tests/multi_gpu_check.py > $out/check_n2.log 2>&1; grep MULTI_GPU_CHECK $out/check_n2.log || tail -5 $out/check_n2.log
bench() {  # name, n, env...
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
bench n${N}_default $N X=1
bench n${N}_inline $N MKE_DRAW_AHEAD=0
bench n2_default 2 X=1
bench n2_inline 2 MKE_DRAW_AHEAD=0
