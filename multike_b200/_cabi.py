"""ctypes binding of ``include/multike_b200.h`` (the thin C-ABI over the sm_100a kernels).

There is deliberately NO fallback: if ``csrc/libmultike_b200.so`` is missing or a symbol is
absent, importing/using the product path raises.  PyTorch is used only to own device memory and
streams; every pointer handed to the library is ``tensor.data_ptr()``.
"""
import ctypes
import os

from . import build as _build

_c = ctypes
c_i32p = _c.c_void_p  # device pointers travel as integers
MKE_ABI_VERSION = 5
MKE_EINVAL = -100000
MKE_MAX_NEG = 32
MKE_MAX_TRY = 10
MKE_MAX_SHARDS = 8


class MkeTable(_c.Structure):
    _fields_ = [
        ("var", _c.c_void_p),
        ("grad", _c.c_void_p),
        ("touched", _c.c_void_p),
        ("rows", _c.c_int32),
        ("stride", _c.c_int32),
        ("dim", _c.c_int32),
        ("normalised", _c.c_int32),
        ("grad_replicas", _c.c_int32),
        ("n_shards", _c.c_int32),
        ("shard_rank", _c.c_int32),
        ("shard_split", _c.c_int32),
        ("shard_pad", _c.c_int32),
        ("peer_var", _c.c_void_p * MKE_MAX_SHARDS),
        ("peer_grad", _c.c_void_p * MKE_MAX_SHARDS),
        ("peer_touched", _c.c_void_p * MKE_MAX_SHARDS),
    ]


class MkeTripleSet(_c.Structure):
    _fields_ = [("slots", _c.c_void_p), ("capacity", _c.c_uint64)]


class MkeKgSampler(_c.Structure):
    _fields_ = [
        ("entity_list", _c.c_void_p),
        ("entity_base", _c.c_int32),
        ("n_entities", _c.c_int32),
        ("neighbours", _c.c_void_p),
        ("n_neighbours", _c.c_int32),
        ("set", MkeTripleSet),
    ]


class MkeRelView(_c.Structure):
    _fields_ = [
        ("ent", _c.POINTER(MkeTable)), ("rel", _c.POINTER(MkeTable)),
        ("ent_acc", _c.c_void_p), ("rel_acc", _c.c_void_p),
        ("lr", _c.c_float),
        ("triples1", _c.c_void_p), ("triples2", _c.c_void_p),
        ("host_triples1", _c.c_void_p), ("host_triples2", _c.c_void_p),
        ("stage1", _c.c_void_p * 2), ("stage2", _c.c_void_p * 2),
        ("n1", _c.c_int32), ("n2", _c.c_int32),
        ("kg1", _c.POINTER(MkeKgSampler)), ("kg2", _c.POINTER(MkeKgSampler)),
        ("batch_size", _c.c_int32), ("K", _c.c_int32),
        ("seed", _c.c_uint64),
        ("neg_ent", _c.c_void_p * 2), ("neg_side", _c.c_void_p * 2),
        ("step_loss", _c.c_void_p), ("host_step_loss", _c.c_void_p),
        ("variant", _c.c_int32),
        ("persist_ws", _c.c_void_p), ("persist_ws_bytes", _c.c_int64),
        ("persist_chunk", _c.c_int32), ("persist_flag_src", _c.c_void_p),
    ]


class MkeRelShardedView(_c.Structure):
    _fields_ = [
        ("ent", _c.POINTER(MkeTable)), ("rel", _c.POINTER(MkeTable)),
        ("ent_acc", _c.c_void_p), ("rel_acc", _c.c_void_p),
        ("lr", _c.c_float),
        ("triples1", _c.c_void_p), ("triples2", _c.c_void_p),
        ("n1", _c.c_int32), ("n2", _c.c_int32),
        ("kg1", _c.POINTER(MkeKgSampler)), ("kg2", _c.POINTER(MkeKgSampler)),
        ("global_batch", _c.c_int32), ("K", _c.c_int32),
        ("seed", _c.c_uint64),
        ("world", _c.c_int32), ("rank", _c.c_int32),
        ("by_kg", _c.c_int32), ("owner_negs", _c.c_int32), ("dummy_row", _c.c_int32), ("variant", _c.c_int32),
        ("neg_ent", _c.c_void_p * 2), ("neg_side", _c.c_void_p * 2), ("neg_valid", _c.c_void_p * 2),
        ("step_loss", _c.c_void_p),
        ("xchg", _c.c_void_p * MKE_MAX_SHARDS), ("sync", _c.c_void_p * MKE_MAX_SHARDS),
        ("host_triples1", _c.c_void_p), ("host_triples2", _c.c_void_p),
        ("stage1", _c.c_void_p * 2), ("stage2", _c.c_void_p * 2),
        ("host_step_loss", _c.c_void_p),
    ]


_PT = _c.POINTER(MkeTable)
_PS = _c.POINTER(MkeTripleSet)
_PK = _c.POINTER(MkeKgSampler)
_vp, _i32, _u64, _f32 = _c.c_void_p, _c.c_int32, _c.c_uint64, _c.c_float

# name -> (restype, argtypes); one entry per declaration in include/multike_b200.h
SIGNATURES = {
    "mke_abi_version": (_i32, []),
    "mke_last_error": (_c.c_char_p, []),
    "mke_launch_count": (_u64, []),
    "mke_triple_fwd_bwd": (_i32, [_PT, _PT, _PT, _vp, _vp, _vp, _i32, _vp, _i32, _f32, _vp, _vp, _vp]),
    "mke_rel_step_sampled": (_i32, [_PT, _PT, _vp, _i32, _PK, _vp, _i32, _PK, _i32, _u64, _u64, _vp,
                                    _f32, _vp, _vp, _i32, _vp]),
    "mke_rel_step_structured": (_i32, [_PT, _PT, _vp, _i32, _i32, _vp, _vp, _vp, _f32, _vp, _i32, _vp]),
    "mke_rel_step_structured2": (_i32, [_PT, _PT, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _f32, _vp, _i32, _vp]),
    "mke_rel_step_structured3": (_i32, [_PT, _PT, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _f32, _vp,
                                        _i32, _vp]),
    "mke_neg_keep_owned": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "mke_neg_keep_owned2": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "mke_rel_step_structured4": (_i32, [_PT, _PT, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _f32,
                                        _vp, _i32, _vp]),
    "mke_rel_train_steps": (_i32, [_c.POINTER(MkeRelView), _i32, _i32, _u64, _c.POINTER(_c.c_int64), _vp, _vp]),
    "mke_rel_sharded_train_steps": (_i32, [_c.POINTER(MkeRelShardedView), _i32, _i32, _u64, _c.POINTER(_c.c_uint32),
                                           _c.POINTER(_c.c_int64), _vp, _vp]),
    "mke_split_tf32": (_i32, [_vp, _vp, _vp, _c.c_int64, _vp]),
    "mke_gemm_tf32x3": (_i32, [_vp, _vp, _c.c_int64, _vp, _vp, _c.c_int64, _i32, _i32, _i32, _vp, _vp, _c.c_int64, _vp]),
    "mke_rel_persist_workspace_bytes": (_c.c_int64, [_i32, _i32, _i32, _i32]),
    "mke_attr_cnn_param_count": (_c.c_int64, [_i32]),
    "mke_attr_cnn_workspace_floats": (_c.c_int64, [_i32, _i32]),
    "mke_attr_cnn_fwd_bwd": (_i32, [_PT, _PT, _PT, _vp, _vp, _vp, _i32, _vp, _f32, _vp, _vp, _vp, _vp, _vp]),
    "mke_dense_apply_adagrad": (_i32, [_vp, _vp, _vp, _c.c_int64, _f32, _vp]),
    "mke_align_fwd_bwd": (_i32, [_PT, _PT, _PT, _PT, _vp, _i32, _f32, _f32, _vp, _vp]),
    "mke_space_mapping_workspace_floats": (_c.c_int64, [_i32, _i32]),
    "mke_space_mapping_fwd_bwd": (_i32, [_PT, _PT, _PT, _PT, _vp, _i32, _vp, _vp, _f32, _f32, _vp, _vp, _vp]),
    "mke_dense_logistic_fwd_bwd": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _i32, _f32, _vp, _vp, _vp, _vp, _vp]),
    "mke_dense_sqdist_fwd_bwd": (_i32, [_vp, _vp, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _vp]),
    "mke_timing_enable": (_i32, [_i32]),
    "mke_timing_stride": (_i32, [_i32]),
    "mke_timing_read": (_i32, [_c.POINTER(_c.c_double), _c.POINTER(_i32)]),
    "mke_rows_apply_adagrad": (_i32, [_PT, _vp, _f32, _vp]),
    "mke_rows_apply_adagrad_pair": (_i32, [_PT, _vp, _f32, _PT, _vp, _f32, _vp]),
    "mke_tripleset_build": (_i32, [_PS, _vp, _i32, _vp]),
    "mke_tripleset_contains": (_i32, [_PS, _vp, _i32, _vp, _vp]),
    "mke_sample_uniform": (_i32, [_vp, _i32, _PK, _vp, _i32, _PK, _i32, _u64, _u64, _vp, _vp]),
    "mke_sample_structured": (_i32, [_vp, _i32, _PK, _vp, _i32, _PK, _i32, _u64, _u64, _vp, _vp, _vp]),
    "mke_sample_structured_at": (_i32, [_vp, _i32, _PK, _vp, _i32, _PK, _i32, _u64, _u64, _i32, _vp, _vp, _vp]),
    "mke_sample_attribute_heads": (_i32, [_vp, _i32, _PK, _vp, _i32, _PK, _i32, _u64, _u64, _i32, _vp, _vp]),
    "mke_peer_alloc": (_i32, [_u64, _c.POINTER(_c.c_void_p)]),
    "mke_peer_free": (_i32, [_vp]),
    "mke_ipc_export": (_i32, [_vp, _c.c_char_p]),
    "mke_ipc_open": (_i32, [_c.c_char_p, _c.POINTER(_c.c_void_p)]),
    "mke_ipc_close": (_i32, [_vp]),
    "mke_sample_distinct": (_i32, [_i32, _i32, _c.c_uint64, _c.c_uint64, _vp, _vp]),
    "mke_table_stage_rows": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "mke_table_commit_grads": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "mke_peer_barrier": (_i32, [_vp, _i32, _i32, _c.c_uint32, _vp]),
    "mke_sim_use_tensor_cores": (_i32, [_i32]),
    "mke_sim_rank_workspace_floats": (_c.c_int64, [_i32, _i32, _i32]),
    "mke_sim_rank": (_i32, [_vp, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mke_sim_topk_workspace_floats": (_c.c_int64, [_i32, _i32, _i32]),
    "mke_sim_topk": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _c.c_int64, _vp, _vp]),
    "mke_table_export": (_i32, [_PT, _vp, _i32, _vp, _vp]),
    "mke_fill_rows": (_i32, [_vp, _i32, _i32, _i32, _f32, _vp]),
}

_lib = None


def library_path():
    return _build.LIB


def load():
    """dlopen the in-tree library (building it first when nvcc and sources allow)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    override = os.environ.get("MKE_LIB_OVERRIDE")  # development only: A/B two builds on one GPU box
    if override:
        path = override
    elif not os.path.exists(path) or (_build._nvcc() and not _build.is_fresh()):
        path = _build.build()
    if not os.path.exists(path):
        raise RuntimeError("multike_b200: CUDA library %s is missing (run __graft_entry__.build())" % path)
    lib = _c.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError => the library does not match the header
        fn.restype = res
        fn.argtypes = args
    if lib.mke_abi_version() != MKE_ABI_VERSION:
        raise RuntimeError("multike_b200: ABI version mismatch")
    _lib = lib
    return lib


class MkeError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = load().mke_last_error().decode("utf-8", "replace")
        kind = "invalid argument" if rc == MKE_EINVAL else "CUDA error %d" % (-rc)
        raise MkeError("multike_b200: %s: %s" % (kind, msg))


def launch_count():
    return int(load().mke_launch_count())


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream():
    import torch

    return torch.cuda.current_stream().cuda_stream
