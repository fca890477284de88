"""ctypes loader for oracle/librf1_oracle.so (TEST INFRASTRUCTURE; see rf1_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Builds the library with `make -C oracle` when it is missing or stale.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librf1_oracle.so")
_SRC = os.path.join(_HERE, "rf1_oracle.c")
D = 256


def build(force: bool = False) -> str:
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(_SRC)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "librf1_oracle.so"], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, i64, u64, i32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int
        L.rf1_fnv1a32.restype = C.c_uint32
        L.rf1_fnv1a32.argtypes = [vp, C.c_size_t]
        L.rf1_tokenize.restype = i64
        L.rf1_tokenize.argtypes = [vp, C.c_size_t, i32, vp, vp, vp, i64]
        L.rf1_n_chunks.restype = i64
        L.rf1_n_chunks.argtypes = [i64]
        L.rf1_featurize_doc.restype = i64
        L.rf1_featurize_doc.argtypes = [vp, C.c_size_t, i32, vp, vp, vp, i64, vp]
        L.rf1_query_vector.restype = None
        L.rf1_query_vector.argtypes = [vp, C.c_size_t, i32, vp]
        L.rf1_dot_isa.restype = C.c_char_p
        L.rf1_max_threads.restype = i32
        L.rf1_score_topk_keys.restype = i32
        L.rf1_score_topk_keys.argtypes = [vp, vp, i32, i64, i64, vp, vp, i32, i32, u64, vp, i32]
        L.rf1_score_topk.restype = i32
        L.rf1_score_topk.argtypes = [vp, vp, vp, i32, i64, vp, vp, i32, i32, u64, vp, vp, vp, i32]
        L.rf1_score_topk_batch.restype = i32
        L.rf1_score_topk_batch.argtypes = [vp, vp, i32, i64, vp, i32, vp, vp, i32, u64, vp, i32]
        L.rf1_merge_topk.restype = i32
        L.rf1_merge_topk.argtypes = [vp, i64, i32, vp]
        L.rf1_cosine.restype = C.c_float
        L.rf1_cosine.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        L.rf1_mix64.restype = u64
        L.rf1_mix64.argtypes = [u64, u64, u64]
        L.rf1_synth_rows.restype = None
        L.rf1_synth_rows.argtypes = [u64, u64, i64, vp, i32, vp, vp, i32]
        L.rf1_synth_query.restype = None
        L.rf1_synth_query.argtypes = [u64, u64, i32, vp, i32, vp]
        L.rf1_bucket_df.restype = i64
        L.rf1_bucket_df.argtypes = [vp, vp, i32, i64, vp, i32, vp]
        L.rf1_idf_weights.restype = None
        L.rf1_idf_weights.argtypes = [vp, u64, i32, vp]
        L.rf1_weight_query.restype = None
        L.rf1_weight_query.argtypes = [vp, vp, i32, vp]
        _lib = L
    return _lib


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


def fnv1a32(data: bytes) -> int:
    return int(lib().rf1_fnv1a32(data, len(data)))


def tokenize(data: bytes, dim: int = D):
    n = int(lib().rf1_tokenize(data, len(data), dim, None, None, None, 0))
    st = np.zeros(max(n, 1), np.int64)
    en = np.zeros(max(n, 1), np.int64)
    bk = np.zeros(max(n, 1), np.uint16)
    lib().rf1_tokenize(data, len(data), dim, _p(st), _p(en), _p(bk), n)
    return st[:n], en[:n], bk[:n]


def featurize_doc(data: bytes, dim: int = D):
    ntok = C.c_int64(0)
    n = int(lib().rf1_featurize_doc(data, len(data), dim, None, None, None, 0, C.byref(ntok)))
    F = np.zeros((max(n, 1), dim), np.int8)
    ff = np.zeros(max(n, 1), np.int32)
    spans = np.zeros((max(n, 1), 2), np.int64)
    lib().rf1_featurize_doc(data, len(data), dim, _p(F), _p(ff), _p(spans), n, C.byref(ntok))
    return F[:n], ff[:n], spans[:n], int(ntok.value)


def query_vector(data: bytes, dim: int = D) -> np.ndarray:
    q = np.zeros(dim, np.int8)
    lib().rf1_query_vector(data, len(data), dim, _p(q))
    return q


def score_topk(F, store_seg, q, scope, k=10, id_base=0, ff=None, threads=0):
    F = np.ascontiguousarray(F, np.int8)
    store_seg = np.ascontiguousarray(store_seg, np.uint32)
    q = np.ascontiguousarray(q, np.int8)
    scope = np.ascontiguousarray(scope, np.uint32)
    ids = np.zeros(k, np.uint64)
    sc = np.zeros(k, np.int32)
    cs = np.zeros(k, np.float32)
    ffp = None if ff is None else _p(np.ascontiguousarray(ff, np.int32))
    if ff is not None:
        ff = np.ascontiguousarray(ff, np.int32)
        ffp = _p(ff)
    n = lib().rf1_score_topk(_p(F), _p(store_seg), ffp, F.shape[1], F.shape[0], _p(q), _p(scope), len(scope), k,
                             id_base, _p(ids), _p(sc), _p(cs), threads)
    if n < 0:
        raise RuntimeError("rf1_score_topk failed")
    return ids[:n], sc[:n], cs[:n]


def score_topk_keys(F, store_seg, q, scope, k=10, id_base=0, row_lo=0, row_hi=None, threads=0):
    F = np.ascontiguousarray(F, np.int8)
    store_seg = np.ascontiguousarray(store_seg, np.uint32)
    q = np.ascontiguousarray(q, np.int8)
    scope = np.ascontiguousarray(scope, np.uint32)
    keys = np.zeros(k, np.uint64)
    hi = F.shape[0] if row_hi is None else row_hi
    n = lib().rf1_score_topk_keys(_p(F), _p(store_seg), F.shape[1], row_lo, hi, _p(q), _p(scope), len(scope), k,
                                  id_base, _p(keys), threads)
    if n < 0:
        raise RuntimeError("rf1_score_topk_keys failed")
    return keys


def merge_topk(keys, k=10):
    keys = np.ascontiguousarray(keys, np.uint64).ravel()
    out = np.zeros(k, np.uint64)
    lib().rf1_merge_topk(_p(keys), keys.size, k, _p(out))
    return out


def synth_rows(seed, start, n, zb, threads=0, with_ff=False, dim: int = D):
    """zb: bucket table for `dim` (rf1.zipf_bucket_table(dim=dim)); any integer dtype."""
    zb = np.ascontiguousarray(zb, np.uint16)
    assert int(zb.max()) < dim, "bucket table was made for a wider row"
    F = np.zeros((n, dim), np.int8)
    ff = np.zeros(n, np.int32)
    lib().rf1_synth_rows(seed, start, n, _p(zb), dim, _p(F), _p(ff), threads)
    return (F, ff) if with_ff else F


def synth_query(seed, qi, zb, n_tokens=8, dim: int = D):
    zb = np.ascontiguousarray(zb, np.uint16)
    q = np.zeros(dim, np.int8)
    lib().rf1_synth_query(seed, qi, n_tokens, _p(zb), dim, _p(q))
    return q


def bucket_df(F, store_seg, scope):
    F = np.ascontiguousarray(F, np.int8)
    seg = np.ascontiguousarray(store_seg, np.uint32)
    sc = np.ascontiguousarray(list(scope), np.uint32)
    df = np.zeros(F.shape[1], np.uint64)
    n = lib().rf1_bucket_df(_p(F), _p(seg), F.shape[1], len(F), _p(sc), len(sc), _p(df))
    return df, int(n)


def idf_weights(df, n):
    df = np.ascontiguousarray(df, np.uint64)
    w = np.zeros(len(df), np.uint8)
    lib().rf1_idf_weights(_p(df), int(n), len(df), _p(w))
    return w


def weight_query(q, w):
    q = np.ascontiguousarray(q, np.int8)
    w = np.ascontiguousarray(w, np.uint8)
    out = np.zeros(len(q), np.int8)
    lib().rf1_weight_query(_p(q), _p(w), len(q), _p(out))
    return out


def dot_isa() -> str:
    return lib().rf1_dot_isa().decode()


def max_threads() -> int:
    return int(lib().rf1_max_threads())
