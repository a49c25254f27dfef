"""Restatement of the LIVE attribute-view score: conv() of code/MultiKE_model.py:34-63 and the
attribute graph of :134-151 (and its ckge / ckga copies :172-185, :203-221), with torch on CPU.

[TF semantics] encoded here (TensorFlow 1.x, not installable; see oracle/__init__.py):
  * tf.layers.batch_normalization(x, 2) -> second positional argument is `axis`: per-column (width
    axis, 75 entries) gamma/beta; training=False (default) -> moving_mean = 0, moving_variance = 1
    are used and never updated: y = gamma * x / sqrt(1 + 1e-3) + beta;
  * tf.layers.conv2d(filters=2, kernel_size=[2,4], padding="same", activation=tanh): cross-
    correlation, NHWC; SAME padding of an even kernel puts the extra cell at the END: rows (0, 1),
    columns (1, 2); kernel variable [kh, kw, in, out], glorot-uniform; bias zero;
  * tf.nn.l2_normalize(_conv, 2): over the width axis per (row, channel);
  * reshape to [-1, 2*75*2]: NHWC flattening, index (h*75 + w)*2 + c;
  * tf.layers.dense(300 -> 75, tanh): kernel [300, 75] glorot-uniform, bias zero;
  * tf.nn.l2_normalize(dense) WITHOUT axis: the global norm of the whole [B, 75] batch tensor.
Parameters travel in ONE flat vector theta with the layout of multike_b200 (see layout()).
"""
import math

import torch

from .tf_semantics import L2_EPS, l2_normalize

BN_EPS = 1e-3
KH, KW = 2, 4


def layout(dim, fmaps=2):
    """name -> (offset, shape) inside the flat parameter vector"""
    out, off = {}, 0
    for name, shape in (("gamma", (dim,)), ("beta", (dim,)), ("k1", (KH, KW, 1, fmaps)), ("b1", (fmaps,)),
                        ("k2", (KH, KW, fmaps, fmaps)), ("b2", (fmaps,)), ("wd", (2 * dim * fmaps, dim)),
                        ("bd", (dim,))):
        n = int(math.prod(shape))
        out[name] = (off, shape)
        off += n
    out["_total"] = (off, ())
    return out


def init_theta(dim, fmaps=2, generator=None, dtype=torch.float64):
    """tf.layers defaults: gamma = 1, beta = 0, glorot-uniform kernels, zero biases"""
    lay = layout(dim, fmaps)
    theta = torch.zeros(lay["_total"][0], dtype=dtype)

    def glorot(shape, fan_in, fan_out):
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(*shape, generator=generator, dtype=dtype) * 2 - 1) * lim

    def put(name, val):
        off, shape = lay[name]
        theta[off:off + val.numel()] = val.reshape(-1)

    put("gamma", torch.ones(dim, dtype=dtype))
    put("k1", glorot((KH, KW, 1, fmaps), KH * KW * 1, KH * KW * fmaps))
    put("k2", glorot((KH, KW, fmaps, fmaps), KH * KW * fmaps, KH * KW * fmaps))
    put("wd", glorot((2 * dim * fmaps, dim), 2 * dim * fmaps, dim))
    return theta


def _get(theta, lay, name):
    off, shape = lay[name]
    return theta[off:off + int(math.prod(shape))].reshape(shape)


def conv_score(attr_hs, attr_as, attr_vs, theta, dim, fmaps=2):
    """MultiKE_model.py:34-63 -> score [B] = -sum((h - dense)^2, 1)"""
    lay = layout(dim, fmaps)
    B = attr_as.shape[0]
    x = torch.stack([attr_as, attr_vs], 1)                               # [B, 2, dim]  (H = 2, W = dim, C = 1)
    x = x * (_get(theta, lay, "gamma") / math.sqrt(1.0 + BN_EPS)) + _get(theta, lay, "beta")   # BN over axis 2
    x = x.unsqueeze(1)                                                   # NCHW [B, 1, 2, dim]
    for kname, bname in (("k1", "b1"), ("k2", "b2")):
        k = _get(theta, lay, kname).permute(3, 2, 0, 1)                  # [out, in, kh, kw]
        x = torch.nn.functional.pad(x, (1, 2, 0, 1))                     # SAME for an even kernel: extra at the end
        x = torch.tanh(torch.nn.functional.conv2d(x, k, _get(theta, lay, bname)))
    x = x.permute(0, 2, 3, 1)                                            # NHWC [B, 2, dim, F]
    x = l2_normalize(x, 2)
    flat = x.reshape(B, -1)
    dense = torch.tanh(flat @ _get(theta, lay, "wd") + _get(theta, lay, "bd"))
    dense = l2_normalize(dense)                                          # important!!  (global norm, :60)
    return -((attr_hs - dense) ** 2).sum(1)


def attribute_cnn_loss(attr_hs, attr_as, attr_vs, ws, theta, dim, scale=1.0):
    """MultiKE_model.py:144-149 (scale = 1, weights), :183 (scale = 2, no weights), :214-218"""
    score = conv_score(attr_hs, attr_as, attr_vs, theta, dim)
    per = torch.log(1 + torch.exp(-score))
    if ws is not None:
        per = per * ws
    return scale * per.sum()
