#!/bin/bash
# usage (GPU box with >= N GPUs): bash tools/multi_session.sh <out> <N...>   -- parity check + bench per world size
OUT=gpurun_out/$1; shift
mkdir -p $OUT
for N in "$@"; do
  for BYKG in 1 0; do
    MKE_BY_KG=$BYKG timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) tests/multi_gpu_check.py > $OUT/check_n${N}_bykg${BYKG}.log 2>&1
    grep -h "MULTI_GPU_CHECK\|d_ent" $OUT/check_n${N}_bykg${BYKG}.log | tail -3
  done
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench.py --gpus $N --steps 200 --warmup 20 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/bench_n$N.json") if l.startswith("{")][-1])
    print("N=$N value %.1f M/s e2e %.1f M/s us/step %.1f p1 %.1f us frac %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"]*1e3, d["roofline"]["launch_ms"]*1e3, d["roofline"]["frac"]))
except Exception as e:
    print("N=$N bench failed", e); print(open("$OUT/bench_n$N.err").read()[-1500:])
PY
done
