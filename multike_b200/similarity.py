"""Host side of csrc/mke_sim.cu: gold-rank evaluator and top-k neighbour search on device rows.

ctypes wrappers only -- the arithmetic is in the kernels (no torch.matmul / topk on this path).
Reference: base/similarity.py:9-52, base/alignment.py:8-79,141-163, base/batch.py:119-150.
"""
import numpy as np
import torch

from multike_b200 import _cabi


def _rows(x, device):
    """fp32 [n, stride] device array (a view when x already is one)."""
    t = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    t = t.to(device=device, dtype=torch.float32)
    if t.dim() != 2:
        raise ValueError("embeddings must be 2-D")
    if t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


def _idx(x, device):
    if x is None:
        return None
    t = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.int32))
    return t.to(device=device, dtype=torch.int32).contiguous()


def sim_rank(emb1, emb2, idx1=None, idx2=None, gold=None, normalize=True, dim=None, device="cuda"):
    """mke_sim_rank -> (rank [n1] int32, top1 [n1] int32), both on `device`.

    emb1/emb2 may be the same table; idx1/idx2 gather rows (None = all rows in order); `dim`
    restricts to the first columns (tables padded to a stride)."""
    lib = _cabi.load()
    a, b = _rows(emb1, device), _rows(emb2, device)
    i1, i2, g = _idx(idx1, device), _idx(idx2, device), _idx(gold, device)
    n1 = a.shape[0] if i1 is None else i1.numel()
    n2 = b.shape[0] if i2 is None else i2.numel()
    dim = int(dim if dim is not None else a.shape[1])
    if a.stride(0) != b.stride(0):
        raise ValueError("both row sets need the same row stride")
    rank = torch.empty(n1, dtype=torch.int32, device=device)
    top1 = torch.empty(n1, dtype=torch.int32, device=device)
    if n1 == 0:
        return rank, top1
    ws = torch.empty(int(lib.mke_sim_rank_workspace_floats(n1, n2, dim)), dtype=torch.float32, device=device)
    _cabi.check(lib.mke_sim_rank(a.data_ptr(), _cabi.ptr(i1), n1, b.data_ptr(), _cabi.ptr(i2), n2, a.stride(0), dim,
                                 1 if normalize else 0, _cabi.ptr(g), ws.data_ptr(), rank.data_ptr(), top1.data_ptr(),
                                 _cabi.current_stream()))
    return rank, top1


def sim_topk(emb, k, idx=None, id_list=None, id_base=0, out=None, out_rows=None, normalize=False, dim=None,
             chunk_rows=8192, device="cuda"):
    """mke_sim_topk -> int32 [n, k] (or `out`, whose rows `out_rows` are written)."""
    lib = _cabi.load()
    a = _rows(emb, device)
    ix, ids, orow = _idx(idx, device), _idx(id_list, device), _idx(out_rows, device)
    n = a.shape[0] if ix is None else ix.numel()
    dim = int(dim if dim is not None else a.shape[1])
    if out is None:
        if orow is not None:
            raise ValueError("out_rows needs an output table")
        out = torch.empty(n, k, dtype=torch.int32, device=device)
    assert out.dtype == torch.int32 and out.is_contiguous() and out.shape[1] == k
    if n == 0:
        return out
    floats = int(lib.mke_sim_topk_workspace_floats(n, dim, max(128, int(chunk_rows))))
    ws = torch.empty(floats, dtype=torch.float32, device=device)
    _cabi.check(lib.mke_sim_topk(a.data_ptr(), _cabi.ptr(ix), n, a.stride(0), dim, 1 if normalize else 0, int(k),
                                 _cabi.ptr(ids), int(id_base), _cabi.ptr(orow), ws.data_ptr(), floats,
                                 out.data_ptr(), _cabi.current_stream()))
    return out
