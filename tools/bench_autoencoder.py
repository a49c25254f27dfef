"""Times one training step of the literal auto-encoder (code/literal_encoder.py:62-69) at the reference's shape --
batch 5000 x 1500 -> 1024 -> 512 -> 75 -> 512 -> 1024 -> 1500, 25.2 MFLOP per literal -- on the
hand-written tcgen05 GEMM (3xTF32: three tensor-core products per logical one) and on fp32 cuBLAS, same step code.
Prints one JSON line.  usage: python tools/bench_autoencoder.py [--steps 20]"""
import argparse
import json
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=5000)
    a = ap.parse_args()
    from multike_b200.refapi.literal_encoder import AutoEncoderModel
    rng = np.random.default_rng(0)
    data = rng.normal(0, 1, (a.batch, 1500)).astype(np.float32)
    dims = [1500, 1024, 512, 75]
    flop = 3 * 2 * a.batch * 2 * sum(x * y for x, y in zip(dims[:-1], dims[1:]))   # fwd + dX + dW, encoder + decoder
    out = {"batch": a.batch, "flop_per_step": flop}
    for mode in ("cublas", "tcgen05"):
        args = types.SimpleNamespace(dim=75, encoder_normalize=True, encoder_active="thah", learning_rate=0.001,
                                     batch_size=a.batch, encoder_gemm=mode)
        m = AutoEncoderModel(data, args, generator=torch.Generator().manual_seed(1))
        batch = m.word_vec_list
        for _ in range(3):
            m._step(batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            loss = m._step(batch)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        out[mode] = {"ms_per_step": ms, "logical_tflops": flop / ms / 1e9, "loss": float(loss)}
    # the GEMM alone, largest product of the step
    from multike_b200 import gemm as G
    x = torch.randn(a.batch, 1500, device="cuda")
    w = torch.randn(1024, 1500, device="cuda")
    xs, ws = G.SplitOperand(x), G.SplitOperand(w)
    for _ in range(3):
        G.gemm_nt(xs, ws)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        G.gemm_nt(xs, ws)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    f = 2.0 * a.batch * 1500 * 1024
    out["gemm_5000x1024x1500"] = {"ms": ms, "logical_tflops": f / ms / 1e9, "tensor_core_tflops": 3 * f / ms / 1e9}
    torch.backends.cuda.matmul.allow_tf32 = False
    for _ in range(3):
        x @ w.t()
    e0.record()
    for _ in range(20):
        x @ w.t()
    e1.record()
    torch.cuda.synchronize()
    out["cublas_fp32_5000x1024x1500"] = {"ms": e0.elapsed_time(e1) / 20, "tflops": f / (e0.elapsed_time(e1) / 20) / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
