python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
python - <<'PY' > gpurun_out/s1_l2attr.log 2>&1
import torch
from cuda import cudart
for name in ["cudaDevAttrMaxPersistingL2CacheSize","cudaDevAttrL2CacheSize","cudaDevAttrMaxAccessPolicyWindowSize"]:
    print(name, cudart.cudaDeviceGetAttribute(getattr(cudart.cudaDeviceAttr,name),0))
PY
for cfg in "0 0" "64 0" "96 0" "0 1" "64 1" "96 1" "64 3" "40 1"; do
  set -- $cfg
  echo "== L2_PERSIST=$1 APPLY_HINT=$2" >> gpurun_out/s1_bench.log
  MKE_L2_PERSIST=$1 MKE_APPLY_HINT=$2 python bench.py --steps 230 --warmup 46 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.1fM ms/step %.4f p1_ms %.4f frac %.3f e2e %.1fM'%(d['value']/1e6,d['ms_per_step'],d['roofline']['launch_ms'],d['roofline']['frac'],d['e2e']['value']/1e6))
    else: print(l.rstrip())
" >> gpurun_out/s1_bench.log
done
