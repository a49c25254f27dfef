"""Host logic of the device-resident triple lists (refapi/MultiKE_model.py::_device_columns): a
Python list of (h, p, t[, w]) tuples is converted once, found again by identity + content fingerprint, reconverted when
the list object or its content changes, and the epoch's shuffle replaces the cached columns."""
import types

import torch

from multike_b200.refapi.MultiKE_model import MultiKE


def _holder():
    return types.SimpleNamespace(device=torch.device("cpu"))


def test_columns_are_cached_by_identity_and_fingerprint():
    m = _holder()
    lst = [(i, i % 3, 2 * i, 0.5 + 0.01 * i) for i in range(50)]
    c = MultiKE._device_columns(m, lst)
    assert [x.dtype for x in c] == [torch.int32] * 3 + [torch.float32]
    assert c[0].tolist() == [t[0] for t in lst] and c[2].tolist() == [t[2] for t in lst]
    assert abs(float(c[3][7]) - lst[7][3]) < 1e-6
    assert MultiKE._device_columns(m, lst) is c                      # same object, same content: the cached copy
    same_content = list(lst)
    assert MultiKE._device_columns(m, same_content) is not c         # another list object: converted on its own
    lst[0] = (99, 0, 0, 1.0)                                         # in-place change of a fingerprinted element
    c2 = MultiKE._device_columns(m, lst)
    assert c2 is not c and int(c2[0][0]) == 99
    lst.append((7, 7, 7, 0.7))                                       # length change
    assert MultiKE._device_columns(m, lst)[0].numel() == 51


def test_triples_without_weights_get_unit_weights_and_shuffles_are_kept():
    m = _holder()
    lst = [(i, 1, i + 1) for i in range(10)]
    c = MultiKE._device_columns(m, lst)
    assert c[3].tolist() == [1.0] * 10
    perm = torch.randperm(10)
    MultiKE._store_columns(m, lst, tuple(x[perm] for x in c))
    again = MultiKE._device_columns(m, lst)
    assert again[0].tolist() == c[0][perm].tolist()                  # the shuffled copy is what the next epoch reads
    assert MultiKE._device_columns(m, [])[0].numel() == 0
