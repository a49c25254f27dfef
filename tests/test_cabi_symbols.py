"""The C-ABI library must load on a CPU-only box and export every symbol include/*.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

from multike_b200 import _cabi
from multike_b200 import build as mke_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "multike_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mke_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_cabi.SIGNATURES)


def test_library_builds_loads_and_exports_everything():
    path = mke_build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    lib2 = _cabi.load()
    assert lib2.mke_abi_version() == _cabi.MKE_ABI_VERSION
    assert lib2.mke_launch_count() == 0 or lib2.mke_launch_count() > 0


def test_bad_arguments_are_reported_not_thrown():
    lib = _cabi.load()
    rc = lib.mke_rows_apply_adagrad(None, None, 0.1, None)
    assert rc == _cabi.MKE_EINVAL
    assert b"apply needs" in lib.mke_last_error()
    with pytest.raises(_cabi.MkeError):
        _cabi.check(rc)


def test_struct_layout_matches_header():
    # mke_table_t: 3 pointers + 9 int32 (+4 pad) + 3 x 8 peer pointers; mke_tripleset_t; mke_kg_sampler_t
    assert ctypes.sizeof(_cabi.MkeTable) == 3 * 8 + 9 * 4 + 4 + 3 * 8 * 8
    assert ctypes.sizeof(_cabi.MkeTripleSet) == 16
    assert ctypes.sizeof(_cabi.MkeKgSampler) == 8 + 4 + 4 + 8 + 4 + 4 + 16


def test_similarity_and_ownership_entry_points_validate_on_the_host():
    """argument errors of the newer entry points are reported before any device work"""
    lib = _cabi.load()
    # both row sets as TF32 part + remainder, the gathered gold rows likewise, gold scores, arg-max keys
    assert lib.mke_sim_rank_workspace_floats(10000, 70000, 75) == 80 * (4 * 10000 + 2 * 70000) + 10000 + 2 * 10000 + 8
    assert lib.mke_sim_rank_workspace_floats(-1, 5, 75) == -1
    assert lib.mke_sim_topk_workspace_floats(100000, 75, 8192) == 2 * 100000 * 80 + 8192 * 100000 + 8
    assert lib.mke_sim_rank(None, None, 5, None, None, 5, 80, 75, 1, None, None, None, None, None) == _cabi.MKE_EINVAL
    assert b"null pointer" in lib.mke_last_error()
    assert lib.mke_sim_rank(None, None, 0, None, None, 5, 80, 75, 1, None, None, None, None, None) == 0  # no rows: no-op
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p).value
    assert lib.mke_sim_topk(p, None, 10, 80, 75, 0, 11, None, 0, None, p, 1 << 20, p, None) == _cabi.MKE_EINVAL
    assert b"k=11" in lib.mke_last_error()
    assert lib.mke_sim_topk(p, None, 10, 200, 150, 0, 3, None, 0, None, p, 1 << 20, p, None) == _cabi.MKE_EINVAL  # dim > 128
    assert lib.mke_neg_keep_owned(p, 4, 10, 3, 0, 0, 0, p, None) == _cabi.MKE_EINVAL
    assert b"n_shards=3" in lib.mke_last_error()
    # the dummy row must live on the caller's shard (KG-block placement: KG2 starts at id 100)
    assert lib.mke_neg_keep_owned(p, 4, 10, 4, 100, 3, 0, p, None) == _cabi.MKE_EINVAL
    assert b"dummy row" in lib.mke_last_error()
    assert lib.mke_timing_stride(0) == _cabi.MKE_EINVAL and lib.mke_timing_stride(4) == 0 and lib.mke_timing_stride(1) == 0
