"""Attribute view: the conv() score of MultiKE_model.py:34-63 as a device object.

One AttrCNN = one conv() instance of the reference (it builds three with independent weights:
attribute view, cross-KG entity inference, cross-KG attribute inference -- SURVEY.md quirk 8).
Parameters live in one flat fp32 vector (layout in include/multike_b200.h); initial values follow
tf.layers defaults [TF semantics]: gamma = 1, beta = 0, glorot-uniform kernels, zero biases.
"""
import math

import numpy as np
import torch

from . import _cabi
from . import tables as T

KH, KW, FMAPS = 2, 4, 2


def param_layout(dim):
    out, off = {}, 0
    for name, shape in (("gamma", (dim,)), ("beta", (dim,)), ("k1", (KH, KW, 1, FMAPS)), ("b1", (FMAPS,)),
                        ("k2", (KH, KW, FMAPS, FMAPS)), ("b2", (FMAPS,)), ("wd", (2 * dim * FMAPS, dim)), ("bd", (dim,))):
        out[name] = (off, shape)
        off += int(np.prod(shape))
    out["_total"] = (off, ())
    return out


class AttrCNN:
    def __init__(self, dim, device="cuda", generator=None, theta=None):
        lib = _cabi.load()
        self.dim, self.device = int(dim), torch.device(device)
        self.layout = param_layout(self.dim)
        n = self.layout["_total"][0]
        assert n == lib.mke_attr_cnn_param_count(self.dim)
        if theta is None:
            th = torch.zeros(n, dtype=torch.float32)

            def glorot(shape, fan_in, fan_out):
                lim = math.sqrt(6.0 / (fan_in + fan_out))
                return (torch.rand(*shape, generator=generator) * 2 - 1) * lim

            def put(name, val):
                off, _ = self.layout[name]
                th[off:off + val.numel()] = val.reshape(-1)

            put("gamma", torch.ones(self.dim))
            put("k1", glorot((KH, KW, 1, FMAPS), KH * KW, KH * KW * FMAPS))
            put("k2", glorot((KH, KW, FMAPS, FMAPS), KH * KW * FMAPS, KH * KW * FMAPS))
            put("wd", glorot((2 * self.dim * FMAPS, self.dim), 2 * self.dim * FMAPS, self.dim))
            theta = th
        self.theta = torch.as_tensor(np.asarray(theta, dtype=np.float32) if not torch.is_tensor(theta) else theta
                                     ).to(self.device, torch.float32).contiguous().clone()
        assert self.theta.numel() == n
        self.grad = torch.zeros_like(self.theta)
        self._slots = {}
        self._ws = None

    def adagrad_slot(self, slot):
        if slot not in self._slots:
            self._slots[slot] = torch.full_like(self.theta, T.ADAGRAD_INIT)
        return self._slots[slot]

    def fwd_bwd(self, ent, attr, val, ih, ia, iv, loss_accum, w=None, scale=1.0):
        """mke_attr_cnn_fwd_bwd: loss and every gradient of one batch (phase 1 of the attribute step)"""
        lib = _cabi.load()
        dev = self.device
        ih, ia, iv, w = T._i32(ih, dev), T._i32(ia, dev), T._i32(iv, dev), T._f32(w, dev)
        n = ih.numel()
        need = int(lib.mke_attr_cnn_workspace_floats(n, self.dim))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.float32, device=dev)
        _cabi.check(lib.mke_attr_cnn_fwd_bwd(ent.c, attr.c, val.c, ih.data_ptr(), ia.data_ptr(), iv.data_ptr(), n,
                                             _cabi.ptr(w), float(scale), self.theta.data_ptr(), self.grad.data_ptr(),
                                             self._ws.data_ptr(), _cabi.ptr(loss_accum), _cabi.current_stream()))
        return ih, ia, iv, w

    def apply_adagrad(self, slot, lr):
        lib = _cabi.load()
        acc = self.adagrad_slot(slot)
        _cabi.check(lib.mke_dense_apply_adagrad(self.theta.data_ptr(), self.grad.data_ptr(), acc.data_ptr(),
                                                self.theta.numel(), float(lr), _cabi.current_stream()))
