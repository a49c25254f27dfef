"""TensorFlow-1.x op semantics the reference relies on, restated with torch (CPU).

Each function names the TF op and the reference call sites.  [TF semantics] = behaviour of the
third-party library (TensorFlow 1.x, version unpinned by the reference: README.md:26), restated
from its documentation because the library cannot be installed here.
"""
import math

import torch

L2_EPS = 1e-12          # tf.nn.l2_normalize(epsilon=1e-12)
ADAGRAD_INIT = 0.1      # tf.train.AdagradOptimizer(initial_accumulator_value=0.1)


def l2_normalize(x, axis=None):
    """tf.nn.l2_normalize: x * rsqrt(max(sum(x**2, axis), 1e-12)); axis=None => global norm.
    Call sites: base/initializers.py:26 (axis=1), MultiKE_model.py:55 (axis=2), :60 (global),
    losses.py:55 (global), literal_encoder.py:66 (global)."""
    if axis is None:
        ss = (x * x).sum()
    else:
        ss = (x * x).sum(dim=axis, keepdim=True)
    return x * torch.rsqrt(torch.clamp(ss, min=L2_EPS))


def xavier_normal_std(shape):
    """tf.contrib.layers.xavier_initializer(uniform=False): truncated normal with
    stddev = sqrt(1.3 * 1.0 / ((fan_in + fan_out) / 2)) (variance_scaling, FAN_AVG)."""
    fan_in, fan_out = shape[0], shape[1]
    return math.sqrt(1.3 / ((fan_in + fan_out) / 2.0))


def xavier_truncated_normal(shape, generator=None, dtype=torch.float32):
    std = xavier_normal_std(shape)
    out = torch.empty(*shape, dtype=dtype)
    torch.nn.init.trunc_normal_(out, 0.0, std, -2 * std, 2 * std, generator=generator)
    return out


def adagrad_dense_(var, acc, grad, lr):
    """tf.train.AdagradOptimizer dense apply (ApplyAdagrad, TF1: no epsilon):
    accum += grad**2 ; var -= lr * grad * rsqrt(accum).  MultiKE_model.py:15-31."""
    acc += grad * grad
    var -= lr * grad * torch.rsqrt(acc)
    return var, acc


def adagrad_sparse_(var, acc, indices, grad_rows, lr):
    """AdagradOptimizer on an IndexedSlices gradient (attr_embeds, is_l2_norm=False):
    duplicates are summed first (_apply_sparse_duplicate_indices -> unsorted_segment_sum),
    then SparseApplyAdagrad on the unique rows."""
    uniq, inv = torch.unique(indices, return_inverse=True)
    summed = torch.zeros(uniq.numel(), grad_rows.shape[1], dtype=grad_rows.dtype)
    summed.index_add_(0, inv, grad_rows)
    acc[uniq] += summed * summed
    var[uniq] -= lr * summed * torch.rsqrt(acc[uniq])
    return var, acc
