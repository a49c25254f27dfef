"""Mirror of code/losses.py: same eight functions, same signatures, 0-d fp32 *sums* out.

Inputs are row-major fp32 CUDA matrices [n, dim] (weights 1-D), as the reference's functions get
them from tf.nn.embedding_lookup.  The six logistic/alignment terms run on the hand-written
kernels of csrc/mke_dense.cu through the C-ABI (forward value and all input gradients in one
pass) and are differentiable through torch.autograd.Function.  space_mapping_loss / orthogonal_loss
(losses.py:53-63; 75x75 matmuls, SURVEY.md section 8 row a-14 = "next") use torch's cuBLAS matmul.
There is no CPU path: CPU tensors raise.
"""
import sys as _sys

if __name__ == "losses":  # imported under the reference's top-level name (refapi first on sys.path):
    import multike_b200.refapi.losses as _canonical  # one module object, whichever name imported it first
    _sys.modules[__name__] = _canonical
import torch

from multike_b200 import _cabi


def _prep(x):
    if not (torch.is_tensor(x) and x.is_cuda):
        raise TypeError("multike_b200 losses need CUDA tensors (no CPU fallback)")
    return x.detach().to(torch.float32).contiguous()


class _Logistic(torch.autograd.Function):
    """scale * sum_i w_i log(1 + exp(+-|h_i + m_i - t_i|^2))"""

    @staticmethod
    def forward(ctx, h, m, t, w, negative, scale):
        lib = _cabi.load()
        hh, mm, tt = _prep(h), _prep(m), _prep(t)
        n, dim = hh.shape
        assert mm.shape == hh.shape and tt.shape == hh.shape, "row matrices must agree in shape"
        ww = None if w is None else _prep(w).reshape(-1)
        assert ww is None or ww.numel() == n
        g = [torch.empty_like(hh) for _ in range(3)]
        acc = torch.zeros(1, dtype=torch.float64, device=hh.device)
        _cabi.check(lib.mke_dense_logistic_fwd_bwd(hh.data_ptr(), mm.data_ptr(), tt.data_ptr(), n, dim, dim,
                                                   _cabi.ptr(ww), int(negative), float(scale), acc.data_ptr(),
                                                   g[0].data_ptr(), g[1].data_ptr(), g[2].data_ptr(),
                                                   _cabi.current_stream()))
        ctx.save_for_backward(*g)
        return acc.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, go):
        gh, gm, gt = ctx.saved_tensors
        return go * gh, go * gm, go * gt, None, None, None


class _SqDist(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        lib = _cabi.load()
        aa, bb = _prep(a), _prep(b)
        assert aa.shape == bb.shape
        n, dim = aa.shape
        ga, gb = torch.empty_like(aa), torch.empty_like(aa)
        acc = torch.zeros(1, dtype=torch.float64, device=aa.device)
        _cabi.check(lib.mke_dense_sqdist_fwd_bwd(aa.data_ptr(), bb.data_ptr(), n, dim, dim, 1.0, acc.data_ptr(),
                                                 ga.data_ptr(), gb.data_ptr(), _cabi.current_stream()))
        ctx.save_for_backward(ga, gb)
        return acc.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, go):
        ga, gb = ctx.saved_tensors
        return go * ga, go * gb


def relation_logistic_loss(phs, prs, pts, nhs, nrs, nts):
    """losses.py:4-12"""
    return _Logistic.apply(phs, prs, pts, None, False, 1.0) + _Logistic.apply(nhs, nrs, nts, None, True, 1.0)


def attribute_logistic_loss(phs, pas, pvs, pws, nhs, nas, nvs, nws):
    """losses.py:15-27"""
    return _Logistic.apply(phs, pas, pvs, pws, False, 1.0) + _Logistic.apply(nhs, nas, nvs, nws, True, 1.0)


def relation_logistic_loss_wo_negs(phs, prs, pts):
    """losses.py:30-34"""
    return _Logistic.apply(phs, prs, pts, None, False, 1.0)


def attribute_logistic_loss_wo_negs(phs, pas, pvs):
    """losses.py:37-41"""
    return _Logistic.apply(phs, pas, pvs, None, False, 1.0)


def logistic_loss_wo_negs(phs, pas, pvs, pws):
    """losses.py:44-50"""
    return _Logistic.apply(phs, pas, pvs, pws, False, 1.0)


def orthogonal_loss(mapping, eye):
    """losses.py:61-63"""
    return ((mapping @ mapping.t() - eye) ** 2).sum()


def space_mapping_loss(view_embeds, shared_embeds, mapping, eye, orthogonal_weight, norm_w=0.0001):
    """losses.py:53-58; tf.nn.l2_normalize without axis = global norm of the batch"""
    mapped = view_embeds @ mapping
    mapped = mapped * torch.rsqrt(torch.clamp((mapped * mapped).sum(), min=1e-12))
    map_loss = _SqDist.apply(shared_embeds, mapped)
    norm_loss = (mapping ** 2).sum()
    return map_loss + orthogonal_weight * orthogonal_loss(mapping, eye) + norm_w * norm_loss


def alignment_loss(ents1, ents2):
    """losses.py:66-69"""
    return _SqDist.apply(ents1, ents2)
