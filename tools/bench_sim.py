"""Times the evaluator / neighbour-search kernels (csrc/mke_sim.cu) at the DBP-WD shapes:
valid() 10 000 x 70 000, test() 60 000 x 60 000 (MultiKE_Late.py:29-61) and the truncated-epsilon
search of one KG, 100 000 rows -> top 2 000 (MultiKE_CSL.py:89-99).  CUDA events, 3 warm-ups."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multike_b200 import similarity as S  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    d = 75
    if "--profile" in sys.argv:  # one launch of each kernel for an ncu capture
        gen = torch.Generator(device="cuda").manual_seed(0)
        a = torch.randn(10000, d, device="cuda", generator=gen)
        b = torch.randn(70000, d, device="cuda", generator=gen)
        S.sim_rank(a, b, normalize=True)
        e = torch.randn(100000, d, device="cuda", generator=gen)
        S.sim_topk(e / e.norm(dim=1, keepdim=True), 2000, chunk_rows=65536)
        torch.cuda.synchronize()
        return
    gen = torch.Generator(device="cuda").manual_seed(0)
    out = {}
    from multike_b200 import _cabi
    lib = _cabi.load()
    if "--fma" in sys.argv:  # the fp32 FMA tiles (baseline) instead of the tensor-core tiles
        lib.mke_sim_use_tensor_cores(0)
    if "--materialise" in sys.argv:  # tensor-core tiles, but the neighbour search writes rows of sims and selects from them
        lib.mke_sim_use_tensor_cores(2)
    out["impl"] = {0: "fp32_fma", 1: "tcgen05_3xtf32, neighbour search without the similarity matrix",
                   2: "tcgen05_3xtf32, neighbour search from materialised sims"}[lib.mke_sim_use_tensor_cores(-1)]
    for name, n1, n2 in (("valid_10k_x_70k", 10000, 70000), ("test_60k_x_60k", 60000, 60000)):
        a = torch.randn(n1, d, device="cuda", generator=gen)
        b = torch.randn(n2, d, device="cuda", generator=gen)
        ms = timed(lambda: S.sim_rank(a, b, normalize=True))
        out[name] = {"ms": ms, "fp32_tflops": 2.0 * n1 * n2 * 80 / ms / 1e9}
    n, k = 100000, 2000
    e = torch.randn(n, d, device="cuda", generator=gen)
    e = e / e.norm(dim=1, keepdim=True)
    for chunk in (8192, 32768):
        ms = timed(lambda: S.sim_topk(e, k, chunk_rows=chunk), reps=2, warm=1)
        out["topk_100k_k2000_chunk%d" % chunk] = {"ms": ms, "fp32_tflops": 2.0 * n * n * 80 / ms / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
