#!/bin/bash
# Multi-GPU session (run under `gpurun --gpus 4`): correctness of the row-sharded relation view with
# both phase-1 schedules, then bench lines at N = 2 and N = 4 for each.
out=gpurun_out/multi_$1
mkdir -p $out
N=${2:-4}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for v in 0 3; do
  MKE_SHARDED_VARIANT=$v run $N 29601 tests/multi_gpu_check.py > $out/check_n${N}_v$v.log 2>&1; grep MULTI_GPU_CHECK $out/check_n${N}_v$v.log || tail -5 $out/check_n${N}_v$v.log
done
for n in 2 $N; do
  for v in 0 3; do
    MKE_SHARDED_VARIANT=$v run $n 29611 bench.py --gpus $n --steps 200 --warmup 20 > $out/bench_n${n}_v$v.json 2> $out/bench_n${n}_v$v.err
    python - $out/bench_n${n}_v$v.json n${n}_v$v <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value %.1f M/s step %.1f us  p1 %.2f us" % (j["value"]/1e6, j["ms_per_step"]*1e3, j["roofline"]["launch_ms"]*1e3))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
  done
done
python bench.py --no-cpu-baseline > $out/bench_n1.json 2>$out/bench_n1.err; tail -c 600 $out/bench_n1.json
