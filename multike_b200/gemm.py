"""Tensor-core GEMM of the literal auto-encoder (csrc/mke_gemm.cu): C = A . B^T (+ bias) with fp32 operands at
fp32-equivalent precision (3xTF32 split on tcgen05 / TMEM, TMA-fed).  No fallback: the C-ABI library must load."""
import torch

from . import _cabi


class SplitOperand:
    """a row-major [rows, K] fp32 matrix as the (hi, lo) pair the kernel reads; rows padded to a 16-byte pitch"""

    def __init__(self, x):
        assert x.dim() == 2 and x.dtype == torch.float32 and x.is_cuda
        rows, k = x.shape
        ld = (k + 3) // 4 * 4
        if ld != k or not x.is_contiguous():
            xp = torch.zeros(rows, ld, dtype=torch.float32, device=x.device)
            xp[:, :k] = x
            x = xp
        self.rows, self.k, self.ld = rows, k, ld
        self.hi, self.lo = torch.empty_like(x), torch.empty_like(x)
        _cabi.check(_cabi.load().mke_split_tf32(x.data_ptr(), self.hi.data_ptr(), self.lo.data_ptr(), x.numel(),
                                                _cabi.current_stream()))


def gemm_nt(a, b, bias=None, out=None):
    """a [M, K] . b [N, K]^T (+ bias [N]) -> [M, N]; a, b: fp32 CUDA tensors or SplitOperand (reuse a split
    weight across calls)"""
    a = a if isinstance(a, SplitOperand) else SplitOperand(a)
    b = b if isinstance(b, SplitOperand) else SplitOperand(b)
    assert a.k == b.k, (a.k, b.k)
    if out is None:
        out = torch.empty(a.rows, b.rows, dtype=torch.float32, device=a.hi.device)
    assert out.shape == (a.rows, b.rows) and out.stride(1) == 1
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == b.rows and bias.is_contiguous()
    _cabi.check(_cabi.load().mke_gemm_tf32x3(a.hi.data_ptr(), a.lo.data_ptr(), a.ld, b.hi.data_ptr(), b.lo.data_ptr(), b.ld,
                                             a.rows, b.rows, a.k, _cabi.ptr(bias), out.data_ptr(), out.stride(0),
                                             _cabi.current_stream()))
    return out
