"""The auto-encoder's tensor-core GEMM (csrc/mke_gemm.cu, tcgen05 + TMA + TMEM, 3xTF32 split) against an fp64
product.  Tolerance: the error of an fp32 GEMM -- the test measures torch's fp32 (non-TF32) matmul on the same
operands and requires ours to be within 3x of it (measured: 0.6x-2.1x), and 20x below what a single-pass TF32 product gives."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    from multike_b200 import _cabi, gemm
    _cabi.load()
    return gemm


SHAPES = [(128, 128, 32), (128, 128, 256), (256, 384, 96), (77, 75, 100), (5000, 1024, 1500), (5000, 75, 512),
          (1500, 1024, 5000), (1, 1, 1), (130, 129, 33), (300, 1500, 75)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_matches_fp64_at_fp32_accuracy(G, M, N, K):
    gen = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", generator=gen)
    b = torch.randn(N, K, device="cuda", generator=gen)
    bias = torch.randn(N, device="cuda", generator=gen)
    got = G.gemm_nt(a, b, bias)
    torch.cuda.synchronize()
    want = a.double() @ b.double().t() + bias.double()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        fp32 = a @ b.t() + bias
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    scale = float(want.abs().max())
    err = float((got.double() - want).abs().max()) / scale
    err32 = float((fp32.double() - want).abs().max()) / scale
    assert err <= max(3 * err32, 1e-6), (err, err32)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True          # what single-pass TF32 operands cost on the same product
    try:
        tf32 = a @ b.t() + bias
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    err_tf32 = float((tf32.double() - want).abs().max()) / scale
    assert K < 64 or err < 0.05 * err_tf32, (err, err_tf32)


def test_gemm_without_bias_into_strided_output_and_reused_split(G):
    gen = torch.Generator(device="cuda").manual_seed(5)
    w = torch.randn(200, 64, device="cuda", generator=gen)
    ws = G.SplitOperand(w)
    big = torch.full((300, 512), 7.0, device="cuda")
    for rows in (300, 17):
        x = torch.randn(rows, 64, device="cuda", generator=gen)
        out = G.gemm_nt(x, ws, out=big[:rows, 100:300])
        torch.cuda.synchronize()
        want = (x.double() @ w.double().t()).float()
        torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)
    assert float(big[:, :100].min()) == 7.0 and float(big[:, 300:].max()) == 7.0   # nothing outside the tile's columns
