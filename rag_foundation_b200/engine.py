"""Python host wrapper over the C-ABI engine (include/rf_b200.h).

numpy in / numpy out for host buffers, raw device pointers + stream handles for the device-resident
entry points (PyTorch supplies those: `tensor.data_ptr()`, `torch.cuda.current_stream().cuda_stream`).
All compute happens in librf_b200.so's sm_100a kernels; this file only marshals arguments.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import threading
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _capi
from ._capi import RF_DIM, RF_SCOPE_MAX, RF_TOPK_MAX, check, lib

_HERE = os.path.dirname(os.path.abspath(__file__))
ZIPF_VOCAB_PATH = os.path.join(_HERE, "data", "zipf_vocab_u16.bin")
NO_ID = np.uint64(0xFFFFFFFFFFFFFFFF)


def load_zipf_vocab() -> np.ndarray:
    t = np.fromfile(ZIPF_VOCAB_PATH, dtype="<u2")
    if t.shape != (65536,):
        raise RuntimeError("zipf_vocab_u16.bin is corrupt")
    return np.ascontiguousarray(t)


def _ptr(a: Optional[np.ndarray]) -> Optional[int]:
    # __array_interface__ is ~10x cheaper than .ctypes.data (no ctypes helper object); it matters
    # on the single-query path, where the whole call is tens of microseconds
    return None if a is None else a.__array_interface__["data"][0]


def scopes_to_csr(scopes: Sequence[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray]:
    off = np.zeros(len(scopes) + 1, np.uint32)
    flat: List[int] = []
    for i, s in enumerate(scopes):
        s = list(s)
        if len(s) > RF_SCOPE_MAX:
            raise ValueError(f"scope of query {i} has {len(s)} store segments (max {RF_SCOPE_MAX})")
        flat.extend(int(x) for x in s)
        off[i + 1] = len(flat)
    segs = np.asarray(flat if flat else [0], dtype=np.uint32)
    return segs, off


def _search_text(owner, call, text: bytes, scope: Sequence[int], k: int, ranges, weights):
    """Shared body of Engine.search_text / EngineGroup.search_text: argument marshalling and the split of a
    long range list over several launches (the kernel plan holds 64 extents)."""
    text = bytes(text)
    if weights is not None:
        weights = np.ascontiguousarray(weights, dtype=np.uint8)
        if weights.shape != (owner.dim,):
            raise ValueError(f"weights must be uint8 [{owner.dim}]")
    rng = None
    if ranges is not None:
        rng = np.ascontiguousarray(np.asarray(list(ranges), dtype=np.uint64).reshape(-1, 2))
        if rng.shape[0] == 0:
            q0 = owner.featurize_query(text)
            return (np.zeros(0, np.uint64), np.zeros(0, np.int32), np.zeros(0, np.float32),
                    q0 if weights is None else owner.weight_query(q0, weights))
        if rng.shape[0] > owner.RANGES_PER_CALL:   # many matching documents: several launches, merged here
            parts = [_search_text(owner, call, text, scope, k, rng[i:i + owner.RANGES_PER_CALL], weights)
                     for i in range(0, rng.shape[0], owner.RANGES_PER_CALL)]
            ids = np.concatenate([p[0] for p in parts]); sc = np.concatenate([p[1] for p in parts])
            cs = np.concatenate([p[2] for p in parts])
            order = np.lexsort((ids, -sc.astype(np.int64)))[:k]   # (score desc, id asc): the RF-1 order
            return ids[order], sc[order], cs[order], parts[0][3]
    segs = np.asarray(list(scope) if len(scope) else [0], dtype=np.uint32)
    ids = np.zeros(k, np.uint64)
    sc = np.zeros(k, np.int32)
    cs = np.zeros(k, np.float32)
    cnt = C.c_uint32()
    q = np.zeros(owner.dim, np.int8)
    call(text, segs, len(scope), rng, weights, k, ids, sc, cs, cnt, q)
    m = int(cnt.value)
    return ids[:m], sc[:m], cs[:m], q


class PendingSearch:
    """A search in flight (Engine.search_begin).  `result()` must be called exactly once."""

    def __init__(self, engine, handle, nq: int, k: int):
        self._e, self._p, self._nq, self._k = engine, handle, nq, k

    def result(self):
        """-> ids uint64 [nq,k], scores int32, cos float32, counts uint32 (as Engine.search)."""
        e, p = self._e, self._p
        if p is None:
            raise RuntimeError("result() was already taken")
        self._p = None
        ids = np.empty((self._nq, self._k), np.uint64)
        sc = np.empty((self._nq, self._k), np.int32)
        cs = np.empty((self._nq, self._k), np.float32)
        cnt = np.empty(self._nq, np.uint32)
        rc = e._L.rf_search_end(e._h, p, _ptr(ids), _ptr(sc), _ptr(cs), _ptr(cnt), None)
        if rc:
            check(rc)
        return ids, sc, cs, cnt


class Engine:
    """One GPU's share of the chunk index (feature arena in HBM + store extents)."""

    RANGES_PER_CALL = 16   # row-range restrictions per launch (the kernel plan holds 64 extents)

    def __init__(self, capacity_rows: int, device: int = 0, id_base: int = 0, n_contexts: int = 8, dim: int = RF_DIM):
        self._L = lib()
        self.dim = int(dim)       # features per chunk row: 256 (default), 512 or 1024
        cfg = _capi.rf_config(C.sizeof(_capi.rf_config), int(device), self.dim, int(n_contexts), int(capacity_rows),
                              int(id_base))
        h = C.c_void_p()
        check(self._L.rf_engine_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.device = int(device)
        self.id_base = int(id_base)
        self.capacity_rows = int(capacity_rows)
        self._zipf = None
        self._lock = threading.Lock()
        self._csr_cache = {}

    # ------------------------------------------------------------------ lifecycle
    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.rf_engine_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self) -> C.c_void_p:
        if not self._h:
            raise RuntimeError("engine is closed")
        return self._h

    def stats(self) -> dict:
        st = _capi.rf_stats()
        check(self._L.rf_engine_stats(self.handle, C.byref(st)))
        return {f: int(getattr(st, f)) for f, _ in st._fields_}

    # ------------------------------------------------------------------ stores
    def open_store(self, fs_name: str) -> int:
        seg = C.c_uint32()
        check(self._L.rf_store_open(self.handle, fs_name.encode("utf-8"), C.byref(seg)))
        return int(seg.value)

    def lookup_store(self, fs_name: str) -> Optional[int]:
        seg = C.c_uint32()
        rc = self._L.rf_store_lookup(self.handle, fs_name.encode("utf-8"), C.byref(seg))
        if rc == _capi.RF_ENOTFOUND:
            return None
        check(rc)
        return int(seg.value)

    def drop_store(self, seg: int) -> None:
        check(self._L.rf_store_drop(self.handle, int(seg)))

    # ------------------------------------------------------------------ ingest
    def ingest_text(self, seg: int, doc_id: int, data: bytes, want_spans: bool = True):
        """-> (first_chunk_id, n_chunks, spans int64 [n_chunks, 2])."""
        data = bytes(data)
        first = C.c_uint64()
        nch = C.c_uint32()
        # a kept token needs >= 1 byte + 1 separator; a chunk past the first needs 112 new tokens
        max_spans = (len(data) // 2 + 1 + 111) // 112 + 1 if want_spans else 0
        spans = np.zeros((max(max_spans, 1), 2), np.int64)
        # a bytes object goes to a void* parameter as a pointer to its own buffer: no copy
        check(self._L.rf_ingest_text(self.handle, int(seg), int(doc_id), data or b"\0", len(data), C.byref(first),
                                     C.byref(nch), _ptr(spans) if want_spans else None, max_spans))
        n = int(nch.value)
        return int(first.value), n, spans[:n].copy()

    def ingest_text_ptr(self, seg: int, doc_id: int, ptr: int, n: int):
        """ingest_text from a raw pointer -- pinned host memory (`PinnedBuffer`) or device memory; the engine
        detects which.  -> (first_chunk_id, n_chunks)."""
        first = C.c_uint64()
        nch = C.c_uint32()
        check(self._L.rf_ingest_text(self.handle, int(seg), int(doc_id), int(ptr), int(n), C.byref(first), C.byref(nch), None, 0))
        return int(first.value), int(nch.value)

    def ingest_features(self, seg: int, doc_id: int, rows, n_rows: Optional[int] = None, on_device: bool = False) -> int:
        first = C.c_uint64()
        if on_device:
            check(self._L.rf_ingest_features(self.handle, int(seg), int(doc_id), int(rows), int(n_rows), 1, C.byref(first)))
        else:
            rows = np.ascontiguousarray(rows, dtype=np.int8).reshape(-1, self.dim)
            check(self._L.rf_ingest_features(self.handle, int(seg), int(doc_id), _ptr(rows), rows.shape[0], 0,
                                             C.byref(first)))
        return int(first.value)

    def ingest_synthetic(self, first_seg: int, rows_per_store: int, seed: int, start_counter: int, n_rows: int) -> int:
        if self._zipf is None:
            self._zipf = load_zipf_vocab()
        first = C.c_uint64()
        check(self._L.rf_ingest_synthetic(self.handle, int(first_seg), int(rows_per_store), int(seed), int(start_counter),
                                          int(n_rows), _ptr(self._zipf), C.byref(first)))
        return int(first.value)

    def tombstone_doc(self, doc_id: int) -> None:
        check(self._L.rf_doc_tombstone(self.handle, int(doc_id)))

    def save_snapshot(self, path: str) -> None:
        check(self._L.rf_snapshot_save(self.handle, os.fsencode(path)))

    def load_snapshot(self, path: str) -> None:
        """Into a freshly created engine (same id_base, capacity >= the snapshot's rows)."""
        check(self._L.rf_snapshot_load(self.handle, os.fsencode(path)))

    # ---- one index, several processes (CUDA IPC; rf_engine_export / attach / refresh) --------------
    def export_state(self) -> bytes:
        """Description of this engine's arena (IPC handles, geometry, store table) for Engine.attach in ANOTHER process."""
        n = C.c_size_t()
        check(self._L.rf_engine_export(self.handle, None, 0, C.byref(n)))
        while True:
            buf = C.create_string_buffer(int(n.value) + 4096)      # (stores may be added between the two calls)
            rc = self._L.rf_engine_export(self.handle, buf, len(buf), C.byref(n))
            if rc == _capi.RF_ENOMEM:
                continue
            check(rc)
            return buf.raw[:int(n.value)]

    @classmethod
    def attach(cls, blob: bytes, n_contexts: int = 8) -> "Engine":
        """A read-only engine over the arena another process exported: searches run here, on this process's own
        streams; ingest / deletes / snapshots stay with the owner."""
        self = cls.__new__(cls)
        self._L = lib()
        h = C.c_void_p()
        check(self._L.rf_engine_attach(blob, len(blob), int(n_contexts), C.byref(h)))
        self._h = h
        dim, device, cap, id_base = struct.unpack_from("<IIQQ", blob, 8)
        self.dim, self.device, self.capacity_rows, self.id_base = int(dim), int(device), int(cap), int(id_base)
        self._zipf = None
        self._lock = threading.Lock()
        self._csr_cache = {}
        self.attached = True
        return self

    def refresh(self, blob: bytes) -> None:
        """Attached engines: take the owner's newer export (rows and stores added since)."""
        check(self._L.rf_engine_refresh(self.handle, blob, len(blob)))

    def read_rows(self, first_row: int, n: int):
        F = np.zeros((n, self.dim), np.int8)
        seg = np.zeros(n, np.uint32)
        ff = np.zeros(n, np.int32)
        check(self._L.rf_rows_read(self.handle, int(first_row), int(n), _ptr(F), _ptr(seg), _ptr(ff)))
        return F, seg, ff

    # ------------------------------------------------------------------ query
    def search(self, q, scopes: Sequence[Sequence[int]], k: int = 10):
        """q int8 [nq, dim] (host). -> ids uint64 [nq,k], scores int32, cos float32, counts uint32.

        `scopes` is one list of store segments per query, or -- for large batches, to skip the
        Python loop -- the CSR pair (segs uint32 [n], off uint32 [nq + 1]) the C-ABI takes."""
        q = np.ascontiguousarray(q, dtype=np.int8).reshape(-1, self.dim)
        nq = q.shape[0]
        if isinstance(scopes, tuple) and len(scopes) == 2 and isinstance(scopes[0], np.ndarray):
            segs = np.ascontiguousarray(scopes[0], dtype=np.uint32)
            off = np.ascontiguousarray(scopes[1], dtype=np.uint32)
            if off.shape != (nq + 1,) or int(off[-1]) > segs.size:
                raise ValueError("CSR scopes: off must have nq + 1 entries ending within segs")
        else:
            if len(scopes) != nq:
                raise ValueError("one scope per query")
            key = tuple(tuple(s) for s in scopes) if nq <= 4 else None
            csr = self._csr_cache.get(key) if key is not None else None
            if csr is None:
                csr = scopes_to_csr(scopes)
                if key is not None and len(self._csr_cache) < 1024:
                    self._csr_cache[key] = csr
            segs, off = csr
        ids = np.empty((nq, k), np.uint64)
        sc = np.empty((nq, k), np.int32)
        cs = np.empty((nq, k), np.float32)
        cnt = np.empty(nq, np.uint32)
        rc = self._L.rf_search(self._h, _ptr(q), nq, _ptr(segs), _ptr(off), k, _ptr(ids), _ptr(sc), _ptr(cs), _ptr(cnt))
        if rc:
            check(rc)
        return ids, sc, cs, cnt

    def search_begin(self, q, scopes: Sequence[Sequence[int]], k: int = 10) -> "PendingSearch":
        """First half of `search` (rf_search_begin): validates, takes a search context and enqueues the work, returns at
        once.  `.result()` (rf_search_end) blocks until the answer is in host memory.  A serving loop keeps a few of
        these in flight per thread -- each on its own context and stream -- so one query's host work and result
        delivery overlap the next query's scan."""
        q = np.ascontiguousarray(q, dtype=np.int8).reshape(-1, self.dim)
        nq = q.shape[0]
        if isinstance(scopes, tuple) and len(scopes) == 2 and isinstance(scopes[0], np.ndarray):
            segs = np.ascontiguousarray(scopes[0], dtype=np.uint32)
            off = np.ascontiguousarray(scopes[1], dtype=np.uint32)
        else:
            if len(scopes) != nq:
                raise ValueError("one scope per query")
            key = tuple(tuple(s) for s in scopes) if nq <= 4 else None
            csr = self._csr_cache.get(key) if key is not None else None
            if csr is None:
                csr = scopes_to_csr(scopes)
                if key is not None and len(self._csr_cache) < 1024:
                    self._csr_cache[key] = csr
            segs, off = csr
        p = C.c_void_p()
        rc = self._L.rf_search_begin(self._h, _ptr(q), nq, _ptr(segs), _ptr(off), k, C.byref(p))
        if rc:
            check(rc)
        return PendingSearch(self, p, nq, k)

    def search_text(self, text: bytes, scope: Sequence[int], k: int = 10, ranges=None, weights=None):
        """-> ids uint64 [m], scores int32 [m], cos float32 [m], q int8 [dim]   (m <= k results).

        `ranges`: optional sorted, disjoint [(lo, hi), ...] global chunk id ranges to stay inside
        (doc-level metadata filters).  `weights`: optional uint8 [dim] RF-1w bucket weights
        (`idf_weights`); the returned q is then the weighted vector."""
        return _search_text(self, self._search_text_call, text, scope, k, ranges, weights)

    def _search_text_call(self, text, segs, n_segs, rng, weights, k, ids, sc, cs, cnt, q):
        check(self._L.rf_search_text_w(self.handle, text or b"\0", len(text), _ptr(segs), n_segs, _ptr(rng),
                                       0 if rng is None else rng.shape[0], _ptr(weights), int(k), _ptr(ids), _ptr(sc),
                                       _ptr(cs), C.byref(cnt), _ptr(q)))

    def featurize_query(self, text: bytes) -> np.ndarray:
        text = bytes(text)
        q = np.zeros(self.dim, np.int8)
        check(self._L.rf_featurize_query(self.handle, text or b"\0", len(text), _ptr(q)))
        return q

    # ---- RF-1w (IDF-weighted variant, oracle/SPEC.md) ------------------------------------------
    def scope_df(self, scope: Sequence[int]) -> Tuple[np.ndarray, int]:
        """-> (df uint64 [dim], n): per-bucket document frequencies over the scope's live rows."""
        segs = np.asarray(list(scope) if len(scope) else [0], dtype=np.uint32)
        df = np.zeros(self.dim, np.uint64)
        n = C.c_uint64()
        check(self._L.rf_scope_df(self.handle, _ptr(segs), len(scope), _ptr(df), C.byref(n)))
        return df, int(n.value)

    def scope_df_device(self, scope: Sequence[int], df_ptr: int, stream: int = 0) -> None:
        """Adds the scope's df[dim] and row count into the dim + 1 u64 at device pointer `df_ptr`."""
        segs = np.asarray(list(scope) if len(scope) else [0], dtype=np.uint32)
        check(self._L.rf_scope_df_device(self.handle, _ptr(segs), len(scope), df_ptr, stream))

    def idf_weights(self, df: np.ndarray, n: int) -> np.ndarray:
        df = np.ascontiguousarray(df, dtype=np.uint64)
        w = np.zeros(self.dim, np.uint8)
        check(self._L.rf_idf_weights(_ptr(df), int(n), self.dim, _ptr(w)))
        return w

    def scope_weights(self, scope: Sequence[int]) -> np.ndarray:
        return self.idf_weights(*self.scope_df(scope))

    def weight_query(self, q: np.ndarray, w: np.ndarray) -> np.ndarray:
        q = np.ascontiguousarray(q, dtype=np.int8)
        w = np.ascontiguousarray(w, dtype=np.uint8)
        out = np.zeros_like(q)
        for row_in, row_out in zip(q.reshape(-1, self.dim), out.reshape(-1, self.dim)):
            check(self._L.rf_weight_query(_ptr(row_in), _ptr(w), self.dim, _ptr(row_out)))
        return out

    def search_keys_device(self, q_ptr: int, nq: int, scope: Sequence[int], k: int, out_keys_ptr: int,
                           stream: int = 0) -> None:
        segs = np.asarray(list(scope) if len(scope) else [0], dtype=np.uint32)
        check(self._L.rf_search_keys_device(self.handle, int(q_ptr), int(nq), _ptr(segs), len(scope), int(k),
                                            int(out_keys_ptr), int(stream) or None))

    def set_stream_overlap(self, stream: int, allow: bool = True) -> None:
        """Promise for the device-resident searches on `stream` (see rf_stream_set_overlap in rf_b200.h): no
        kernel enqueued between two searches writes their query buffer, so scan launches may overlap the
        previous kernel's tail (programmatic dependent launch)."""
        check(self._L.rf_stream_set_overlap(self.handle, int(stream) or None, 1 if allow else 0))

    def search_keys_device_scoped(self, q_ptr: int, nq: int, scopes, k: int, out_keys_ptr: int, stream: int = 0) -> None:
        """Device-resident search with one scope per query (`scopes`: list of lists, or a CSR tuple)."""
        segs, off = scopes if isinstance(scopes, tuple) else scopes_to_csr(scopes)
        if len(off) != nq + 1:
            raise ValueError("one scope per query")
        check(self._L.rf_search_keys_device_scoped(self.handle, q_ptr, nq, _ptr(segs), _ptr(off), k, out_keys_ptr, stream))

    def search_keys_device_fused(self, q_ptr: int, nq: int, scope: Sequence[int], k: int, out_keys_ptr: int, stream: int,
                                 rank: int, world: int, nq_cap: int, seq: int, keys_ptrs: np.ndarray, flag_ptrs: np.ndarray,
                                 timeout_flag_ptr: int) -> None:
        """Sharded search with the top-k exchange fused into the kernels (NVLink peer stores +
        flags) instead of an NCCL all-gather; see rf_search_keys_device_fused in rf_b200.h."""
        segs = np.asarray(list(scope) if len(scope) else [0], dtype=np.uint32)
        px = _capi.rf_peer_exchange(C.sizeof(_capi.rf_peer_exchange), int(rank), int(world), int(nq_cap), int(k), int(seq),
                                    keys_ptrs.ctypes.data, flag_ptrs.ctypes.data, int(timeout_flag_ptr))
        check(self._L.rf_search_keys_device_fused(self.handle, int(q_ptr), int(nq), _ptr(segs), len(scope), int(k), C.byref(px),
                                                  int(out_keys_ptr), int(stream) or None))

    def search_keys_device_scoped_fused(self, q_ptr: int, nq: int, scopes, k: int, out_keys_ptr: int, stream: int,
                                        rank: int, world: int, nq_cap: int, seq: int, keys_ptrs: np.ndarray, flag_ptrs: np.ndarray,
                                        timeout_flag_ptr: int, nq_total: int = 0, q_index: Optional[np.ndarray] = None,
                                        owner_masks: Optional[np.ndarray] = None) -> None:
        """Store-sharded batch with the exchange in the kernels (rf_search_keys_device_scoped_fused): one scope
        per LAUNCHED query (`scopes`: list of lists or a CSR tuple of THIS rank's segments).  `q_index` (uint32 [nq])
        names the batch queries this rank launches (the device query buffer holds all `nq_total`), `owner_masks`
        (uint8 [nq_total]) the ranks each batch query's merge waits for."""
        segs, off = scopes if isinstance(scopes, tuple) else scopes_to_csr(scopes)
        if len(off) != nq + 1:
            raise ValueError("one scope per query")
        px = _capi.rf_peer_exchange(C.sizeof(_capi.rf_peer_exchange), int(rank), int(world), int(nq_cap), int(k), int(seq),
                                    keys_ptrs.ctypes.data, flag_ptrs.ctypes.data, int(timeout_flag_ptr),
                                    int(nq_total or 0), _ptr(q_index), _ptr(owner_masks))
        check(self._L.rf_search_keys_device_scoped_fused(self.handle, int(q_ptr), int(nq), _ptr(segs), _ptr(off), int(k), C.byref(px),
                                                         int(out_keys_ptr), int(stream) or None))

    def merge_topk_device(self, keys_ptr: int, n_lists: int, nq: int, k: int, out_keys_ptr: int, stream: int = 0) -> None:
        check(self._L.rf_merge_topk_device(self.handle, int(keys_ptr), int(n_lists), int(nq), int(k), int(out_keys_ptr),
                                           int(stream) or None))


class PinnedBuffer:
    """Page-locked host memory from rf_host_alloc, exposed as a numpy uint8 array (`.array`): read an upload
    straight into it (`file.readinto(buf.array)`) and hand `buf.ptr` to `Engine.ingest_text_ptr` -- the DMA then
    runs from the caller's own buffer, no staging copy."""

    def __init__(self, n: int):
        self._L = lib()
        p = C.c_void_p()
        check(self._L.rf_host_alloc(int(n), C.byref(p)))
        self.ptr = int(p.value)
        self.n = int(n)
        self.array = np.ctypeslib.as_array((C.c_uint8 * self.n).from_address(self.ptr))

    def close(self) -> None:
        p, self.ptr = getattr(self, "ptr", 0), 0
        if p:
            self.array = None
            self._L.rf_host_free(p)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class EngineGroup:
    """Several GPUs behind one index in THIS process (include/rf_b200.h, rf_group_*): the same surface as
    Engine for everything the adapter uses, so `Registry(EngineGroup([...]))` is a drop-in.  Stores keep one
    number on every device, chunk ids are global, a search runs on every device that holds rows of the scope
    and the per-device top-k lists are merged on the host (bit-identical to one engine holding everything).
    `devices` may repeat an ordinal (two engines on one GPU) -- how the single-GPU test-suite covers it."""

    RANGES_PER_CALL = Engine.RANGES_PER_CALL

    def __init__(self, devices: Sequence[int], capacity_rows: int, n_contexts: int = 8, placement: str = "store",
                 id_bases: Optional[Sequence[int]] = None, dim: int = RF_DIM):
        self._L = lib()
        placements = {"store": _capi.RF_PLACE_STORE, "spread": _capi.RF_PLACE_SPREAD}
        if placement not in placements:
            raise ValueError("placement must be 'store' or 'spread'")
        devs = np.asarray(list(devices), dtype=np.int32)
        bases = None if id_bases is None else np.asarray(list(id_bases), dtype=np.uint64)
        if bases is not None and bases.shape != devs.shape:
            raise ValueError("one id base per device")
        cfg = _capi.rf_group_config(C.sizeof(_capi.rf_group_config), int(devs.size), _ptr(devs), _ptr(bases), int(n_contexts),
                                    placements[placement], int(capacity_rows), int(dim), 0)
        h = C.c_void_p()
        check(self._L.rf_group_create(C.byref(cfg), C.byref(h)), group=True)
        self._h = h
        self.dim = int(dim)
        self.devices = [int(d) for d in devs]
        self.placement = placement
        self.capacity_rows = int(capacity_rows)
        self._zipf = None

    # ------------------------------------------------------------------ lifecycle
    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.rf_group_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self) -> C.c_void_p:
        if not self._h:
            raise RuntimeError("engine group is closed")
        return self._h

    def stats(self, per_device: bool = False):
        st = _capi.rf_stats()
        per = (_capi.rf_stats * len(self.devices))()
        check(self._L.rf_group_stats(self.handle, C.byref(st), C.cast(per, C.c_void_p)), group=True)
        tot = {f: int(getattr(st, f)) for f, _ in st._fields_}
        if not per_device:
            return tot
        return tot, [{f: int(getattr(p, f)) for f, _ in p._fields_} for p in per]

    # ------------------------------------------------------------------ stores / ingest / deletes
    def open_store(self, fs_name: str) -> int:
        s = C.c_uint32()
        check(self._L.rf_group_store_open(self.handle, fs_name.encode("utf-8"), C.byref(s)), group=True)
        return int(s.value)

    def lookup_store(self, fs_name: str) -> Optional[int]:
        s = C.c_uint32()
        rc = self._L.rf_group_store_lookup(self.handle, fs_name.encode("utf-8"), C.byref(s))
        if rc == _capi.RF_ENOTFOUND:
            return None
        check(rc, group=True)
        return int(s.value)

    def drop_store(self, store: int) -> None:
        check(self._L.rf_group_store_drop(self.handle, int(store)), group=True)

    def ingest_text(self, store: int, doc_id: int, data: bytes, want_spans: bool = True):
        data = bytes(data)
        first = C.c_uint64()
        nch = C.c_uint32()
        max_spans = (len(data) // 2 + 1 + 111) // 112 + 1 if want_spans else 0
        spans = np.zeros((max(max_spans, 1), 2), np.int64)
        check(self._L.rf_group_ingest_text(self.handle, int(store), int(doc_id), data or b"\0", len(data), C.byref(first),
                                           C.byref(nch), _ptr(spans) if want_spans else None, max_spans), group=True)
        n = int(nch.value)
        return int(first.value), n, spans[:n].copy()

    def ingest_features(self, store: int, doc_id: int, rows) -> int:
        rows = np.ascontiguousarray(rows, dtype=np.int8).reshape(-1, self.dim)
        first = C.c_uint64()
        check(self._L.rf_group_ingest_features(self.handle, int(store), int(doc_id), _ptr(rows), rows.shape[0], C.byref(first)), group=True)
        return int(first.value)

    def ingest_synthetic(self, first_store: int, rows_per_store: int, seed: int, start_counter: int, n_rows: int) -> None:
        if self._zipf is None:
            self._zipf = load_zipf_vocab()
        check(self._L.rf_group_ingest_synthetic(self.handle, int(first_store), int(rows_per_store), int(seed), int(start_counter),
                                                int(n_rows), _ptr(self._zipf)), group=True)

    def tombstone_doc(self, doc_id: int) -> None:
        check(self._L.rf_group_doc_tombstone(self.handle, int(doc_id)), group=True)

    def save_snapshot(self, path: str) -> None:
        check(self._L.rf_group_snapshot_save(self.handle, os.fsencode(path)), group=True)

    def load_snapshot(self, path: str) -> None:
        check(self._L.rf_group_snapshot_load(self.handle, os.fsencode(path)), group=True)

    # ------------------------------------------------------------------ query
    def search(self, q, scopes, k: int = 10):
        """As Engine.search: q int8 [nq, dim] (host), one list of stores per query or a CSR pair."""
        q = np.ascontiguousarray(q, dtype=np.int8).reshape(-1, self.dim)
        nq = q.shape[0]
        if isinstance(scopes, tuple) and len(scopes) == 2 and isinstance(scopes[0], np.ndarray):
            segs = np.ascontiguousarray(scopes[0], dtype=np.uint32)
            off = np.ascontiguousarray(scopes[1], dtype=np.uint32)
            if off.shape != (nq + 1,) or int(off[-1]) > segs.size:
                raise ValueError("CSR scopes: off must have nq + 1 entries ending within segs")
        else:
            if len(scopes) != nq:
                raise ValueError("one scope per query")
            segs, off = scopes_to_csr(scopes)
        ids = np.empty((nq, k), np.uint64)
        sc = np.empty((nq, k), np.int32)
        cs = np.empty((nq, k), np.float32)
        cnt = np.empty(nq, np.uint32)
        rc = self._L.rf_group_search(self._h, _ptr(q), nq, _ptr(segs), _ptr(off), k, _ptr(ids), _ptr(sc), _ptr(cs), _ptr(cnt))
        if rc:
            check(rc, group=True)
        return ids, sc, cs, cnt

    def search_text(self, text: bytes, scope: Sequence[int], k: int = 10, ranges=None, weights=None):
        return _search_text(self, self._search_text_call, text, scope, k, ranges, weights)

    def _search_text_call(self, text, segs, n_segs, rng, weights, k, ids, sc, cs, cnt, q):
        check(self._L.rf_group_search_text(self.handle, text or b"\0", len(text), _ptr(segs), n_segs, _ptr(rng),
                                           0 if rng is None else rng.shape[0], _ptr(weights), int(k), _ptr(ids), _ptr(sc),
                                           _ptr(cs), C.byref(cnt), _ptr(q)), group=True)

    def engine_handle(self, index: int) -> C.c_void_p:
        h = C.c_void_p()
        check(self._L.rf_group_engine(self.handle, int(index), C.byref(h)), group=True)
        return h

    def featurize_query(self, text: bytes) -> np.ndarray:
        text = bytes(text)
        q = np.zeros(self.dim, np.int8)
        check(self._L.rf_featurize_query(self.engine_handle(0), text or b"\0", len(text), _ptr(q)))
        return q

    def scope_df(self, scope: Sequence[int]) -> Tuple[np.ndarray, int]:
        segs = np.asarray(list(scope) if len(scope) else [0], dtype=np.uint32)
        df = np.zeros(self.dim, np.uint64)
        n = C.c_uint64()
        check(self._L.rf_group_scope_df(self.handle, _ptr(segs), len(scope), _ptr(df), C.byref(n)), group=True)
        return df, int(n.value)

    idf_weights = Engine.idf_weights
    weight_query = Engine.weight_query

    def scope_weights(self, scope: Sequence[int]) -> np.ndarray:
        return self.idf_weights(*self.scope_df(scope))


def probe_int8_peak(device: int = 0, n_batches: int = 4000) -> Tuple[float, float]:
    """-> (int8 ops/s, ms): the tensor pipe's achievable dense int8 rate on `device` (rf_probe_int8_peak)."""
    ops, ms = C.c_double(), C.c_double()
    check(lib().rf_probe_int8_peak(int(device), int(n_batches), C.byref(ops), C.byref(ms)))
    return float(ops.value), float(ms.value)


def unpack_keys(keys: np.ndarray):
    """Packed RF-1 keys -> (ids uint64, scores int32, valid bool)."""
    keys = np.asarray(keys, dtype=np.uint64)
    valid = keys != 0
    ids = np.where(valid, np.uint64(0xFFFFFFFF) - (keys & np.uint64(0xFFFFFFFF)), NO_ID)
    scores = (keys >> np.uint64(32)).astype(np.int64).astype(np.int32)
    return ids, scores, valid
