#!/usr/bin/env python3
"""Secondary measurements for the BASELINE.json configs that are not the bench.py headline:

  cfg3   synthetic single store, 1M chunks, batched 1024 queries, top-10 on 1 B200
  cfg5   multi-tenant: 10k stores x 10k chunks, 1024 store-scoped queries per batch (+ CPU oracle on a slice)
  ingest featurisation throughput (text MB/s in, chunks/s out) on synthetic ASCII text

One JSON line per measurement (CUDA events on the launching stream, >= 3 warm-ups, results
parity-checked against the C oracle on a sample first).  Run under gpurun; copies go to profiles/.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle as co, rf1  # noqa: E402  (checker + CPU baseline only)
from rag_foundation_b200 import Engine  # noqa: E402

HBM_PEAK = 6550.1
try:
    HBM_PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def emit(d):
    print(json.dumps(d), flush=True)


def cfg3(n_rows: int, nq: int, reps: int):
    zb = rf1.zipf_bucket_table()
    Q = np.stack([co.synth_query(0, i, zb) for i in range(nq)])
    with Engine(capacity_rows=n_rows) as e:
        s = e.open_store("fileSearchStores/cfg3")
        e.ingest_synthetic(s, 0, seed=0, start_counter=0, n_rows=n_rows)
        qd = torch.from_numpy(Q).cuda()
        out = torch.zeros((nq, 10), dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        e.search_keys_device(qd.data_ptr(), nq, [s], 10, out.data_ptr(), st)
        torch.cuda.synchronize()
        got = out.cpu().numpy().view(np.uint64)
        F = co.synth_rows(0, 0, n_rows, zb)
        seg = np.zeros(n_rows, np.uint32)
        bad = sum(got[i].tolist() != co.score_topk_keys(F, seg, Q[i], [s]).tolist() for i in range(0, nq, max(1, nq // 16)))
        for _ in range(3):
            e.search_keys_device(qd.data_ptr(), nq, [s], 10, out.data_ptr(), st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            e.search_keys_device(qd.data_ptr(), nq, [s], 10, out.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        t0 = time.perf_counter()
        ncpu = 8
        for i in range(ncpu):
            co.score_topk_keys(F, seg, Q[i], [s])
        cpu_q = (time.perf_counter() - t0) / ncpu
        emit({"config": "cfg3: 1M chunks, batched %d queries, top-10, 1 B200" % nq, "kernel": "score_topk_gemm (tcgen05 kind::i8) unless RF_GEMM=0 / nq < 64, then score_topk_scan",
              "ms_per_batch": ms, "qps": nq / (ms * 1e-3), "chunks_per_s": nq * n_rows / (ms * 1e-3),
              "int8_mac_ops_per_s": 2.0 * nq * n_rows * 256 / (ms * 1e-3), "parity_mismatches_sampled": bad,
              "cpu_oracle_ms_per_query": cpu_q * 1e3, "cpu_threads": co.max_threads()})


def cfg5(n_stores: int, per_store: int, nq: int, reps: int):
    zb = rf1.zipf_bucket_table()
    n_rows = n_stores * per_store
    rng = np.random.default_rng(5)
    with Engine(capacity_rows=n_rows) as e:
        first = e.open_store("fileSearchStores/mt0")
        for i in range(1, n_stores):
            e.open_store(f"fileSearchStores/mt{i}")
        e.ingest_synthetic(first, per_store, seed=5, start_counter=0, n_rows=n_rows)
        scopes = [[int(first + rng.integers(0, n_stores))] for _ in range(nq)]
        Q = np.stack([co.synth_query(5, i, zb) for i in range(nq)])
        ids, sc, cs, cnt = e.search(Q, scopes, k=10)
        bad = 0
        t_cpu = 0.0
        n_cpu = min(nq, 32)
        for i in range(n_cpu):   # oracle on the scoped store only (regenerated from its counters)
            st_i = scopes[i][0] - first
            F = co.synth_rows(5, st_i * per_store, per_store, zb)
            t0 = time.perf_counter()
            w_ids, w_sc, _ = co.score_topk(F, np.zeros(per_store, np.uint32), Q[i], [0], id_base=st_i * per_store)
            t_cpu += time.perf_counter() - t0
            bad += int(ids[i].tolist() != w_ids.tolist() or sc[i].tolist() != w_sc.tolist())
        from rag_foundation_b200.engine import scopes_to_csr
        csr = scopes_to_csr(scopes)                      # the C-ABI's own scope format
        ids2, sc2, _, _ = e.search(Q, csr, k=10)
        bad += int(not (ids2 == ids).all() or not (sc2 == sc).all())
        for _ in range(3):
            e.search(Q, csr, k=10)
        reps = max(reps, 20)
        t0 = time.perf_counter()
        for _ in range(reps):
            e.search(Q, csr, k=10)
        wall_ms = (time.perf_counter() - t0) / reps * 1e3
        alg = nq * per_store * 260
        emit({"config": "cfg5: %d stores x %d chunks, %d store-scoped queries per batch, 1 B200" % (n_stores, per_store, nq),
              "kernel": "score_topk_scan (grid.y = queries, per-query extents + row mask)", "e2e_ms_per_batch": wall_ms,
              "e2e_qps": nq / (wall_ms * 1e-3), "algorithmic_bytes_per_batch": alg,
              "e2e_GBps": alg / (wall_ms * 1e-3) / 1e9, "frac_of_measured_hbm_peak_e2e": alg / (wall_ms * 1e-3) / 1e9 / HBM_PEAK,
              "parity_mismatches": bad, "parity_checked": n_cpu,
              "cpu_oracle_ms_per_query": t_cpu / n_cpu * 1e3, "cpu_threads": co.max_threads()})


def ingest(mb: int, reps: int):
    n_tokens = mb * 1_000_000 // 4
    data = rf1.synth_text(0, n_tokens)
    with Engine(capacity_rows=max(1024, 2 * (n_tokens // 112) * (reps + 2))) as e:
        s = e.open_store("fileSearchStores/ingest")
        first, n_chunks, spans = e.ingest_text(s, 1, data)
        wF, wff, wsp, ntok = co.featurize_doc(data)
        F, sg, ff = e.read_rows(0, n_chunks)
        ok = bool(n_chunks == len(wF) and (F == wF).all() and (ff == wff).all() and (spans == wsp).all())
        t0 = time.perf_counter()
        for r in range(reps):
            e.ingest_text(s, 2 + r, data, want_spans=False)
        wall = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        co.featurize_doc(data)
        cpu = time.perf_counter() - t0
        emit({"config": "ingest featurisation: %.1f MB synthetic ASCII text per document" % (len(data) / 1e6), "n_tokens": ntok,
              "n_chunks": n_chunks, "parity_ok": ok, "e2e_ms_per_doc": wall * 1e3, "text_MBps": len(data) / wall / 1e6,
              "chunks_per_s": n_chunks / wall, "cpu_oracle_ms_per_doc_1_thread": cpu * 1e3,
              "note": "rf_ingest_text from a pageable host buffer: H2D copy + 4 kernels + 2 stream syncs inside the timed region"})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="cfg3,cfg5,ingest")
    ap.add_argument("--cfg3-nq", type=int, default=1024)
    ap.add_argument("--cfg5-stores", type=int, default=10_000)
    ap.add_argument("--cfg5-per-store", type=int, default=10_000)
    args = ap.parse_args()
    which = args.only.split(",")
    if "cfg3" in which:
        cfg3(1_000_000, args.cfg3_nq, 2)
    if "cfg5" in which:
        cfg5(args.cfg5_stores, args.cfg5_per_store, 1024, 5)
    if "ingest" in which:
        ingest(24, 3)
        ingest(1, 10)
