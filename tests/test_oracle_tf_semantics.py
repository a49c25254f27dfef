"""Cross-checks of the [TF semantics] restatements (oracle/tf_semantics.py) against independent
implementations of the same published definitions in torch: they do not replace TensorFlow (which
cannot run here) but rule out slips in the restatement that every parity test leans on."""
import math

import numpy as np
import pytest
import torch

from oracle import tf_semantics as tfs


def test_adagrad_equals_torch_optim_adagrad_without_epsilon():
    """ApplyAdagrad (accum0 = 0.1, no epsilon) == torch.optim.Adagrad(initial_accumulator_value=0.1, eps=0)"""
    gen = torch.Generator().manual_seed(0)
    v0 = torch.randn(7, 5, generator=gen, dtype=torch.float64)
    var, acc = v0.clone(), torch.full_like(v0, tfs.ADAGRAD_INIT)
    p = torch.nn.Parameter(v0.clone())
    opt = torch.optim.Adagrad([p], lr=0.004, initial_accumulator_value=0.1, eps=0.0)
    for _ in range(4):
        g = torch.randn(7, 5, generator=gen, dtype=torch.float64)
        g[2] = 0.0                                        # a row without gradient: a no-op
        tfs.adagrad_dense_(var, acc, g, 0.004)
        p.grad = g.clone()
        opt.step()
    torch.testing.assert_close(var, p.detach(), rtol=0, atol=1e-15)
    assert torch.equal(var[2], v0[2])


def test_sparse_adagrad_sums_duplicates_before_the_update():
    gen = torch.Generator().manual_seed(1)
    v0 = torch.randn(6, 4, generator=gen, dtype=torch.float64)
    idx = torch.tensor([1, 4, 1, 1, 5])
    rows = torch.randn(5, 4, generator=gen, dtype=torch.float64)
    var, acc = v0.clone(), torch.full_like(v0, 0.1)
    tfs.adagrad_sparse_(var, acc, idx, rows, 0.01)
    dense = torch.zeros_like(v0).index_add_(0, idx, rows)
    var2, acc2 = v0.clone(), torch.full_like(v0, 0.1)
    tfs.adagrad_dense_(var2, acc2, dense, 0.01)
    torch.testing.assert_close(var, var2, rtol=0, atol=1e-15)
    torch.testing.assert_close(acc, acc2, rtol=0, atol=1e-15)


def test_l2_normalize_equals_functional_normalize_with_matching_epsilon():
    """x * rsqrt(max(sum x^2, 1e-12)) == x / max(||x||, 1e-6) (torch.nn.functional.normalize)"""
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(9, 6, generator=gen, dtype=torch.float64)
    x[3] *= 1e-9                                          # below the clamp: scaled by 1e6, not to unit norm
    x[4] = 0.0
    torch.testing.assert_close(tfs.l2_normalize(x, 1), torch.nn.functional.normalize(x, dim=1, eps=1e-6), rtol=1e-12, atol=0)
    g = tfs.l2_normalize(x)                               # axis=None: one global norm
    assert float((g * g).sum()) == pytest.approx(1.0, rel=1e-12)
    assert float(tfs.l2_normalize(x, 1)[3].norm()) == pytest.approx(float(x[3].norm()) * 1e6, rel=1e-9)


def test_xavier_truncated_normal_statistics():
    """xavier_initializer(uniform=False) = variance_scaling (FAN_AVG, truncated normal): the reference's
    [200 000, 75] entity table gets stddev sqrt(1.3 * 2 / (fan_in + fan_out)) = 0.003605 before truncation
    (SURVEY.md a-3), values within 2 stddev"""
    std = tfs.xavier_normal_std((200000, 75))
    assert std == pytest.approx(math.sqrt(1.3 * 2 / 200075), rel=1e-12) and std == pytest.approx(0.003605, rel=1e-3)
    x = tfs.xavier_truncated_normal((20000, 75), torch.Generator().manual_seed(3), torch.float64)
    s = tfs.xavier_normal_std((20000, 75))
    assert float(x.abs().max()) <= 2 * s and abs(float(x.mean())) < 5 * s / math.sqrt(x.numel())
    assert float(x.std()) == pytest.approx(0.8796 * s, rel=0.01)     # std of a normal truncated at 2 sigma
