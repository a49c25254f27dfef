"""The torch restatement of conv() (oracle/attr_cnn.py, MultiKE_model.py:34-63) against an
index-by-index numpy restatement of the same TF-1.x semantics written from the TF definitions
(NHWC cross-correlation, SAME padding of an even kernel, width-wise and global l2_normalize, NHWC
flattening): guards the oracle's padding / permutation / flatten order, which the GPU kernels are
then compared with."""
import math

import numpy as np
import torch

from oracle import attr_cnn as oc


def conv2d_same_nhwc(x, k, b):
    """tf.layers.conv2d(padding='same', strides 1): out[n,i,j,o] = b[o] + sum x[n,i+di-pt,j+dj-pl,c] k[di,dj,c,o]
    with total padding kh-1 / kw-1 split floor/ceil (the extra cell goes to the bottom / right)."""
    n, h, w, c = x.shape
    kh, kw, _, o = k.shape
    pt, pl = (kh - 1) // 2, (kw - 1) // 2
    out = np.zeros((n, h, w, o))
    for i in range(h):
        for j in range(w):
            for di in range(kh):
                for dj in range(kw):
                    ii, jj = i + di - pt, j + dj - pl
                    if 0 <= ii < h and 0 <= jj < w:
                        out[:, i, j, :] += x[:, ii, jj, :] @ k[di, dj]
    return out + b


def conv_score_numpy(hs, as_, vs, theta, dim):
    lay = oc.layout(dim)
    get = lambda name: theta[lay[name][0]: lay[name][0] + int(np.prod(lay[name][1]))].reshape(lay[name][1])
    x = np.stack([as_, vs], 1)[..., None]                                  # [B, 2, dim, 1]
    x = x * (get("gamma") / math.sqrt(1 + 1e-3))[None, None, :, None] + get("beta")[None, None, :, None]
    x = np.tanh(conv2d_same_nhwc(x, get("k1"), get("b1")))
    x = np.tanh(conv2d_same_nhwc(x, get("k2"), get("b2")))
    x = x / np.sqrt(np.maximum((x ** 2).sum(2, keepdims=True), 1e-12))     # l2_normalize(_conv, 2)
    flat = x.reshape(x.shape[0], -1)                                       # (h * dim + w) * 2 + c
    dense = np.tanh(flat @ get("wd") + get("bd"))
    dense = dense / math.sqrt(max((dense ** 2).sum(), 1e-12))              # l2_normalize without axis: global
    return -((hs - dense) ** 2).sum(1)


def test_conv_score_matches_index_level_restatement():
    dim, B = 7, 5
    gen = torch.Generator().manual_seed(3)
    theta = oc.init_theta(dim, generator=gen)
    theta += 0.05 * torch.randn(theta.shape, generator=gen, dtype=torch.float64)   # non-trivial biases / beta
    hs, as_, vs = (torch.randn(B, dim, generator=gen, dtype=torch.float64) for _ in range(3))
    got = oc.conv_score(hs, as_, vs, theta, dim).numpy()
    want = conv_score_numpy(hs.numpy(), as_.numpy(), vs.numpy(), theta.numpy(), dim)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    # the loss of the three graphs: weights, and the factor 2 of the cross-KG graph (:183)
    w = torch.rand(B, generator=gen, dtype=torch.float64)
    per = np.log(1 + np.exp(-want))
    assert float(oc.attribute_cnn_loss(hs, as_, vs, w, theta, dim)) == np.float64((per * w.numpy()).sum()).item() or \
        abs(float(oc.attribute_cnn_loss(hs, as_, vs, w, theta, dim)) - (per * w.numpy()).sum()) < 1e-12
    assert abs(float(oc.attribute_cnn_loss(hs, as_, vs, None, theta, dim, scale=2.0)) - 2 * per.sum()) < 1e-12


def test_parameter_layout_counts():
    lay = oc.layout(75)
    assert lay["_total"][0] == 75 + 75 + 2 * 4 * 1 * 2 + 2 + 2 * 4 * 2 * 2 + 2 + 300 * 75 + 75 == 22777   # SURVEY a-12
