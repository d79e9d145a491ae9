"""ctypes binding of libskm_b200.so (the C ABI in include/skm_b200.h).

There is no CPU fallback: if the library has not been built, or no B200 is visible, every
entry point raises.  The oracle under oracle/ is test infrastructure and is never imported
from this package.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_lib", "libskm_b200.so")

SKM_OK, SKM_ERR_INVALID, SKM_ERR_CUDA, SKM_ERR_NOMEM, SKM_ERR_UNSUPPORTED, SKM_ERR_STATE = range(6)
SKM_F32, SKM_F64, SKM_I32, SKM_I64, SKM_U16 = range(5)

_i64 = C.c_int64
_vp = C.c_void_p
_dbl = C.c_double
_int = C.c_int


class SkmError(RuntimeError):
    """A non-zero status from libskm_b200 (`code` is one of the SKM_ERR_* values)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libskm_b200 error {code}: {message}")
        self.code = code
        self.message = message


REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p)


class DatasetInfo(C.Structure):
    _fields_ = [("p", _i64), ("n", _i64), ("nnz", _i64), ("max_col_nnz", _i64),
                ("store_dtype", C.c_int32), ("reserved", C.c_int32),
                ("device_bytes", _i64), ("stream_bytes", _i64)]


class IterStats(C.Structure):
    _fields_ = [("dff", _dbl), ("sumsq", _dbl), ("n_empty", _i64), ("n_rechecked", _i64),
                ("n_points", _i64), ("has_nan", C.c_int32), ("reserved", C.c_int32)]


# name -> (restype, argtypes); every symbol include/skm_b200.h declares
SIGNATURES = {
    "skm_abi_version": (_int, []),
    "skm_ctx_create": (_int, [_int, _vp, C.POINTER(_vp)]),
    "skm_ctx_destroy": (None, [_vp]),
    "skm_last_error": (C.c_char_p, [_vp]),
    "skm_ctx_stream": (_vp, [_vp]),
    "skm_ctx_device": (_int, [_vp]),
    "skm_ctx_sync": (_int, [_vp]),
    "skm_ctx_launch_count": (_i64, [_vp]),
    "skm_ctx_tc_chunks": (_int, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "skm_ctx_timing_enable": (_int, [_vp, _int]),
    "skm_ctx_timing_read": (_int, [_vp, _vp, _vp]),
    "skm_sparse_matrix_minus_cluster": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _int, _dbl, _vp]),
    "skm_sparse_matrix_inner_product": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "skm_sparse_matrix_column_normsq": (_int, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "skm_hadamard": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "skm_dataset_create_csc": (_int, [_vp, _i64, _i64, _vp, _int, _vp, _int, _vp, _int, _int, _int,
                                      C.POINTER(_vp)]),
    "skm_dataset_create_csc_hint": (_int, [_vp, _i64, _i64, _vp, _int, _vp, _int, _vp, _int, _int, _int, _i64,
                                           C.POINTER(_vp)]),
    "skm_dataset_alloc_csc": (_int, [_vp, _i64, _i64, _i64, C.POINTER(_vp)]),
    "skm_dataset_csc_ptrs": (_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "skm_dataset_commit": (_int, [_vp]),
    "skm_dataset_minmax": (_int, [_vp, C.POINTER(_dbl), C.POINTER(_dbl)]),
    "skm_dataset_destroy": (None, [_vp]),
    "skm_dataset_get_info": (_int, [_vp, C.POINTER(DatasetInfo)]),
    "skm_dataset_get_column": (_int, [_vp, _i64, _vp]),
    "skm_dataset_layout_check": (_int, [_vp, _int, _vp]),
    "skm_assign": (_int, [_vp, _vp, _i64, _int, _dbl, _vp, _vp]),
    "skm_assign_sparse_centers": (_int, [_vp, _vp, _i64, _int, _dbl, _vp, _vp]),
    "skm_masked_distances": (_int, [_vp, _vp, _i64, _vp]),
    "skm_lloyd_create": (_int, [_vp, _i64, C.POINTER(_vp)]),
    "skm_lloyd_destroy": (None, [_vp]),
    "skm_lloyd_set_centers": (_int, [_vp, _vp]),
    "skm_lloyd_get_centers": (_int, [_vp, _vp]),
    "skm_lloyd_get_centers_old": (_int, [_vp, _vp]),
    "skm_lloyd_set_center_column": (_int, [_vp, _i64, _vp]),
    "skm_lloyd_assign": (_int, [_vp, _int, _dbl]),
    "skm_lloyd_assign_sparse": (_int, [_vp, _int, _dbl]),
    "skm_lloyd_accumulate": (_int, [_vp]),
    "skm_lloyd_set_assign_mode": (_int, [_vp, _int]),
    "skm_lloyd_last_assign": (_int, [_vp, C.POINTER(_i64)]),
    "skm_lloyd_set_prune": (_int, [_vp, _int]),
    "skm_lloyd_last_prune": (_int, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "skm_lloyd_set_tc_filter": (_int, [_vp, _int]),
    "skm_lloyd_last_tc": (_int, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "skm_debug_tc_scores": (_int, [_vp, _int, C.c_double, _vp, C.POINTER(_i64)]),
    "skm_lloyd_set_update_mode": (_int, [_vp, _int]),
    "skm_lloyd_last_update": (_int, [_vp, C.POINTER(_int), C.POINTER(_i64)]),
    "skm_lloyd_partials": (_vp, [_vp, C.POINTER(_i64)]),
    "skm_lloyd_finalize": (_int, [_vp, _dbl, _int, C.POINTER(IterStats)]),
    "skm_lloyd_refresh_diff": (_int, [_vp, C.POINTER(IterStats)]),
    "skm_lloyd_get_counts": (_int, [_vp, _vp]),
    "skm_lloyd_get_assignments": (_int, [_vp, _vp, _vp]),
    "skm_lloyd_argmax_distance": (_int, [_vp, C.POINTER(_dbl), C.POINTER(_i64)]),
    "skm_lloyd_kernel_name": (C.c_char_p, [_vp]),
    "skm_lloyd_assign_ptr": (_vp, [_vp]),
    "skm_lloyd_dist_ptr": (_vp, [_vp, C.POINTER(_int)]),
    "skm_lloyd_step_host": (_int, [_vp, _i64, _i64, _vp, _int, _vp, _int, _vp, _int, _vp, _i64, _int, _dbl, _dbl,
                                   _int, _i64, _vp, _vp, _vp, C.POINTER(IterStats), _vp, _vp]),
    "skm_second_pass": (_int, [_vp, _i64, _i64, _vp, _int, _int, _dbl, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64,
                               C.POINTER(_i64)]),
    "skm_kpp_update": (_int, [_vp, _vp, _int, _dbl, _int, C.POINTER(_dbl)]),
    "skm_kpp_update_sparse": (_int, [_vp, _vp, _int, C.POINTER(_dbl)]),
    "skm_kpp_pick": (_int, [_vp, _dbl, C.POINTER(_i64)]),
    "skm_kpp_get_mindist": (_int, [_vp, _vp]),
    "skm_multi_create": (_int, [_int, _vp, C.POINTER(_vp)]),
    "skm_multi_destroy": (None, [_vp]),
    "skm_multi_ndev": (_int, [_vp]),
    "skm_multi_ctx": (_vp, [_vp, _int]),
    "skm_multi_peer_access": (_int, [_vp]),
    "skm_multi_dataset_create_csc": (_int, [_vp, _i64, _i64, _vp, _int, _vp, _int, _vp, _int, _int, C.POINTER(_vp)]),
    "skm_multi_dataset_from_shards": (_int, [_vp, _vp, C.POINTER(_vp)]),
    "skm_multi_dataset_destroy": (None, [_vp]),
    "skm_multi_dataset_shard": (_vp, [_vp, _int, C.POINTER(_i64)]),
    "skm_multi_dataset_get_info": (_int, [_vp, C.POINTER(DatasetInfo)]),
    "skm_multi_dataset_get_column": (_int, [_vp, _i64, _vp]),
    "skm_multi_kpp_update": (_int, [_vp, _vp, _int, _dbl, _int, _int, C.POINTER(_dbl)]),
    "skm_multi_kpp_pick": (_int, [_vp, _dbl, C.POINTER(_i64)]),
    "skm_multi_lloyd_create": (_int, [_vp, _i64, C.POINTER(_vp)]),
    "skm_multi_lloyd_destroy": (None, [_vp]),
    "skm_multi_lloyd_set_modes": (_int, [_vp, _int, _int]),
    "skm_multi_lloyd_set_centers": (_int, [_vp, _vp]),
    "skm_multi_lloyd_set_center_column": (_int, [_vp, _i64, _vp]),
    "skm_multi_lloyd_get_centers": (_int, [_vp, _vp]),
    "skm_multi_lloyd_get_centers_old": (_int, [_vp, _vp]),
    "skm_multi_lloyd_get_centers_of": (_int, [_vp, _int, _vp]),
    "skm_multi_lloyd_step": (_int, [_vp, _int, _dbl, _dbl, _int, _int, C.POINTER(IterStats)]),
    "skm_multi_lloyd_refresh_diff": (_int, [_vp, C.POINTER(IterStats)]),
    "skm_multi_lloyd_get_counts": (_int, [_vp, _vp]),
    "skm_multi_lloyd_get_assignments": (_int, [_vp, _vp, _vp]),
    "skm_multi_lloyd_argmax_distance": (_int, [_vp, C.POINTER(_dbl), C.POINTER(_i64)]),
    "skm_multi_lloyd_launch_count": (_int, [_vp, C.POINTER(_i64)]),
    "skm_mix_hadamard": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _int, _vp]),
    "skm_fwht_sample_f32": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, C.c_uint64, _i64, C.POINTER(_vp)]),
    "skm_dataset_from_dense_host": (_int, [_vp, _i64, _i64, _i64, _vp, _int, _vp, _i64, C.c_uint64, _i64, _i64,
                                           C.POINTER(_vp)]),
    "skm_sample_rows": (_int, [_vp, _i64, _i64, _i64, C.c_uint64, _i64, _vp]),
    "skm_dct_mix": (_int, [_vp, _i64, _i64, _vp, _vp, _int, _vp]),
    "skm_dataset_from_dense_host_dct": (_int, [_vp, _i64, _i64, _vp, _int, _vp, _i64, C.c_uint64, _i64, _vp, _i64,
                                               C.POINTER(_vp)]),
    "skm_sample_rows_general": (_int, [_vp, _i64, _i64, _i64, C.c_uint64, _i64, _vp]),
    "skm_fwht_f32_inplace": (_int, [_vp, _i64, _i64, _vp, _vp]),
}

_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Load libskm_b200.so and bind every symbol; raises if the library is missing."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m sparsifiedkmeans_b200.build` "
                "(or __graft_entry__.build()). sparsifiedkmeans_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int) -> None:
    if rc != SKM_OK:
        msg = load().skm_last_error(None)
        raise SkmError(rc, msg.decode() if msg else "")
