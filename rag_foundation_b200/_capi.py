"""ctypes binding of include/rf_b200.h (the C-ABI of librf_b200.so).

There is no CPU fallback: if the shared library is missing this module raises at import of the
first symbol, and every compute call fails loudly when no sm_100 GPU is visible.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librf_b200.so")

RF_DIM = 256          # default row width; engines can also be created with 512 or 1024 (RF_DIM_MAX)
RF_DIM_MAX = 1024
RF_TOPK_MAX = 32
RF_SCOPE_MAX = 16
RF_TOMBSTONE = 0xFFFFFFFF

RF_OK, RF_EINVAL, RF_ENOMEM, RF_ECUDA, RF_ENODEVICE, RF_ENOTFOUND, RF_EBUSY, RF_ECAPACITY = 0, -1, -2, -3, -4, -5, -6, -7


class rf_config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("dim", C.c_uint32),
                ("n_contexts", C.c_uint32), ("capacity_rows", C.c_uint64), ("id_base", C.c_uint64)]


class rf_peer_exchange(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("rank", C.c_uint32), ("world", C.c_uint32), ("nq_cap", C.c_uint32),
                ("k", C.c_uint32), ("seq", C.c_uint32), ("keys_ptrs", C.c_void_p), ("flag_ptrs", C.c_void_p),
                ("timeout_flag_dev", C.c_void_p), ("nq_total", C.c_uint32), ("q_index", C.c_void_p), ("owner_masks", C.c_void_p)]


class rf_group_config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_devices", C.c_uint32), ("devices", C.c_void_p), ("id_bases", C.c_void_p),
                ("n_contexts", C.c_uint32), ("placement", C.c_uint32), ("capacity_rows", C.c_uint64),
                ("dim", C.c_uint32), ("reserved", C.c_uint32)]


RF_PLACE_STORE, RF_PLACE_SPREAD = 0, 1
RF_GROUP_MAX = 8


class rf_stats(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("capacity_rows", C.c_uint64), ("n_stores", C.c_uint64),
                ("n_docs", C.c_uint64), ("hbm_bytes", C.c_uint64), ("searches", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("free_rows", C.c_uint64), ("ingest_bytes", C.c_uint64),
                ("ingest_kernel_ns", C.c_uint64)]


# name -> (restype, argtypes); the test-suite checks every one of these is exported
_vp, _u32, _u64, _i32, _sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_size_t
SIGNATURES = {
    "rf_engine_create": (_i32, [C.POINTER(rf_config), C.POINTER(_vp)]),
    "rf_engine_destroy": (_i32, [_vp]),
    "rf_engine_stats": (_i32, [_vp, C.POINTER(rf_stats)]),
    "rf_strerror": (C.c_char_p, [_i32]),
    "rf_last_error": (C.c_char_p, []),
    "rf_build_info": (_i32, [C.c_char_p, _sz]),
    "rf_debug_timestamps": (_i32, [_vp, _vp, _u64, _i32]),
    "rf_store_open": (_i32, [_vp, C.c_char_p, C.POINTER(_u32)]),
    "rf_store_lookup": (_i32, [_vp, C.c_char_p, C.POINTER(_u32)]),
    "rf_store_drop": (_i32, [_vp, _u32]),
    "rf_ingest_text": (_i32, [_vp, _u32, _u64, _vp, _sz, C.POINTER(_u64), C.POINTER(_u32), _vp, _u32]),
    "rf_host_alloc": (_i32, [_sz, C.POINTER(_vp)]),
    "rf_host_free": (_i32, [_vp]),
    "rf_ingest_features": (_i32, [_vp, _u32, _u64, _vp, _u64, _i32, C.POINTER(_u64)]),
    "rf_ingest_synthetic": (_i32, [_vp, _u32, _u64, _u64, _u64, _u64, _vp, C.POINTER(_u64)]),
    "rf_doc_tombstone": (_i32, [_vp, _u64]),
    "rf_snapshot_save": (_i32, [_vp, C.c_char_p]),
    "rf_snapshot_load": (_i32, [_vp, C.c_char_p]),
    "rf_engine_export": (_i32, [_vp, _vp, _sz, C.POINTER(_sz)]),
    "rf_engine_attach": (_i32, [_vp, _sz, _u32, C.POINTER(_vp)]),
    "rf_engine_refresh": (_i32, [_vp, _vp, _sz]),
    "rf_rows_read": (_i32, [_vp, _u64, _u64, _vp, _vp, _vp]),
    "rf_search": (_i32, [_vp, _vp, _u32, _vp, _vp, _u32, _vp, _vp, _vp, _vp]),
    "rf_search_begin": (_i32, [_vp, _vp, _u32, _vp, _vp, _u32, C.POINTER(_vp)]),
    "rf_search_text_begin": (_i32, [_vp, _vp, _sz, _vp, _u32, _vp, _u32, _vp, _u32, C.POINTER(_vp)]),
    "rf_search_end": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rf_search_text": (_i32, [_vp, _vp, _sz, _vp, _u32, _u32, _vp, _vp, _vp, _vp, _vp]),
    "rf_search_text_in": (_i32, [_vp, _vp, _sz, _vp, _u32, _vp, _u32, _u32, _vp, _vp, _vp, _vp, _vp]),
    "rf_search_keys_device": (_i32, [_vp, _vp, _u32, _vp, _u32, _u32, _vp, _vp]),
    "rf_stream_set_overlap": (_i32, [_vp, _vp, _i32]),
    "rf_search_keys_device_scoped": (_i32, [_vp, _vp, _u32, _vp, _vp, _u32, _vp, _vp]),
    "rf_search_keys_device_fused": (_i32, [_vp, _vp, _u32, _vp, _u32, _u32, C.POINTER(rf_peer_exchange), _vp, _vp]),
    "rf_search_keys_device_scoped_fused": (_i32, [_vp, _vp, _u32, _vp, _vp, _u32, C.POINTER(rf_peer_exchange), _vp, _vp]),
    "rf_merge_topk_device": (_i32, [_vp, _vp, _u32, _u32, _u32, _vp, _vp]),
    "rf_featurize_query": (_i32, [_vp, _vp, _sz, _vp]),
    "rf_scope_df": (_i32, [_vp, _vp, _u32, _vp, C.POINTER(_u64)]),
    "rf_scope_df_device": (_i32, [_vp, _vp, _u32, _vp, _vp]),
    "rf_idf_weights": (_i32, [_vp, _u64, _u32, _vp]),
    "rf_weight_query": (_i32, [_vp, _vp, _u32, _vp]),
    "rf_search_text_w": (_i32, [_vp, _vp, _sz, _vp, _u32, _vp, _u32, _vp, _u32, _vp, _vp, _vp, _vp, _vp]),
    "rf_probe_int8_peak": (_i32, [_i32, _u32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "rf_group_create": (_i32, [C.POINTER(rf_group_config), C.POINTER(_vp)]),
    "rf_group_destroy": (_i32, [_vp]),
    "rf_group_size": (_i32, [_vp, C.POINTER(_u32)]),
    "rf_group_engine": (_i32, [_vp, _u32, C.POINTER(_vp)]),
    "rf_group_stats": (_i32, [_vp, C.POINTER(rf_stats), _vp]),
    "rf_group_last_error": (C.c_char_p, []),
    "rf_group_store_open": (_i32, [_vp, C.c_char_p, C.POINTER(_u32)]),
    "rf_group_store_lookup": (_i32, [_vp, C.c_char_p, C.POINTER(_u32)]),
    "rf_group_store_drop": (_i32, [_vp, _u32]),
    "rf_group_ingest_text": (_i32, [_vp, _u32, _u64, _vp, _sz, C.POINTER(_u64), C.POINTER(_u32), _vp, _u32]),
    "rf_group_ingest_features": (_i32, [_vp, _u32, _u64, _vp, _u64, C.POINTER(_u64)]),
    "rf_group_ingest_synthetic": (_i32, [_vp, _u32, _u64, _u64, _u64, _u64, _vp]),
    "rf_group_doc_tombstone": (_i32, [_vp, _u64]),
    "rf_group_search": (_i32, [_vp, _vp, _u32, _vp, _vp, _u32, _vp, _vp, _vp, _vp]),
    "rf_group_search_text": (_i32, [_vp, _vp, _sz, _vp, _u32, _vp, _u32, _vp, _u32, _vp, _vp, _vp, _vp, _vp]),
    "rf_group_scope_df": (_i32, [_vp, _vp, _u32, _vp, C.POINTER(_u64)]),
    "rf_group_snapshot_save": (_i32, [_vp, C.c_char_p]),
    "rf_group_snapshot_load": (_i32, [_vp, C.c_char_p]),
}

_lib = None


class RfError(RuntimeError):
    def __init__(self, code: int, detail: str):
        super().__init__(f"librf_b200: {detail} (code {code})")
        self.code = code


def lib() -> C.CDLL:
    """Load librf_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "rag_foundation_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code: int, group: bool = False) -> None:
    """Map C-ABI error codes onto the exceptions the reference's callers handle:
    RF_EBUSY -> TimeoutError (retryable: gemini_rag.py:22-27, routes/chat.py:1076);
    everything else -> RuntimeError (-> `unexpected_error` frame, routes/chat.py:1130-1143)."""
    if code == RF_OK:
        return
    L = lib()
    detail = ((L.rf_group_last_error() if group else L.rf_last_error()) or b"").decode("utf-8", "replace") or L.rf_strerror(code).decode()
    if code == RF_EBUSY:
        raise TimeoutError(f"librf_b200: {detail}")
    raise RfError(code, detail)
