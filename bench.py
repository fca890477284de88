#!/usr/bin/env python3
"""bench.py -- the reference's headline metric on B200: chunks scored/s and QPS of store-scoped
top-10 retrieval, with the fraction of the HBM roofline and the CPU oracle timed beside it.

  python bench.py --gpus 1 --steps K --warmup W            # configs[1]: 1M chunks, 1 query/step
  torchrun ... bench.py --gpus N --steps K --warmup W      # configs[3]: 100M chunks sharded by chunk
  python bench.py --impl reference ...                     # the CPU arm (RF-1 C oracle, all host cores)

A step is ONE query through the hot path: the fused score + top-10 kernel over every chunk in
scope (at N > 1: per-rank scan with the top-k exchange fused into the kernel over NVLink peer
memory; `--exchange nccl` selects the all-gather + merge-kernel path instead).
`value` is chunks scored per second over all GPUs with inputs resident in HBM, timed with CUDA
events on the launching stream (max over ranks).  `e2e` is the same metric through the public
C-ABI call with HOST buffers (query upload and result download inside the timed region).
The reference has no retrieval arithmetic of its own (oracle/SPEC.md), so the reference arm is the
RF-1 C oracle -- `cpu_baseline.kind == "port"`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BYTES_PER_CHUNK = 260            # 256 B int8 features + 4 B store-segment word
CFG2_ROWS = 1_000_000
CFG4_ROWS = 100_000_000
SEED = 0
N_DISTINCT_QUERIES = 64
METRIC = "chunks scored/s, top-10 store-scoped retrieval (QPS alongside)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(workload: str):
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload)
        except Exception:
            return None
    return None


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self._nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_oracle_leg(n_rows: int, budget_s: float, queries: np.ndarray, gpu_keys=None, seed: int = SEED):
    """Time the RF-1 C oracle (all host threads) on a bounded sample: the first min(n_rows, 4M)
    chunks of the same synthetic corpus, as many of `queries` as fit in ~budget_s."""
    from oracle import c_oracle as co, rf1
    zb = rf1.zipf_bucket_table()
    sample_rows = min(n_rows, 4_000_000)
    F = co.synth_rows(seed, 0, sample_rows, zb)
    seg = np.zeros(sample_rows, np.uint32)
    threads = co.max_threads()
    co.score_topk_keys(F, seg, queries[0], [0])   # warm
    t0 = time.perf_counter()
    done = 0
    mismatches = 0
    while True:   # cycle through the queries until ~budget_s of CPU work has been timed
        i = done % len(queries)
        keys = co.score_topk_keys(F, seg, queries[i], [0])
        if done < len(queries) and gpu_keys is not None and sample_rows == n_rows and keys.tolist() != gpu_keys[i].tolist():
            mismatches += 1
        done += 1
        if time.perf_counter() - t0 > budget_s and done >= min(len(queries), 8):
            break
    dt = time.perf_counter() - t0
    return {"value": sample_rows * done / dt, "unit": "chunks/s", "cores": threads, "kind": "port",
            "sample": f"{done} queries x first {sample_rows} chunks of the workload corpus (seed {seed}), "
                      f"RF-1 C oracle ({co.dot_isa()}, OpenMP {threads} threads), {dt:.2f} s",
            "qps_on_sample": done / dt, "parity_mismatches": mismatches if gpu_keys is not None and sample_rows == n_rows else None}


def make_queries(n: int, seed: int = SEED) -> np.ndarray:
    """RF-1 synthetic queries (oracle/SPEC.md).  Generated with the product's own table so the
    bench never needs the oracle on the GPU arm: same mix64 / bucket rule, vectorised in numpy."""
    from rag_foundation_b200.engine import load_zipf_vocab
    zv = load_zipf_vocab()

    def fnv_bucket(v: int) -> int:
        h = 0x811C9DC5
        for b in str(v).encode():
            h = ((h ^ b) * 0x01000193) & 0xFFFFFFFF
        return h & 255

    bucket_of = {}
    M = (1 << 64) - 1
    out = np.zeros((n, 256), np.int8)
    for i in range(n):
        for j in range(8):
            x = ((seed ^ 0x51) * 0x9E3779B97F4A7C15 + i * 0xBF58476D1CE4E5B9 + j * 0x94D049BB133111EB + 0x2545F4914F6CDD1D) & M
            x ^= x >> 30; x = (x * 0xBF58476D1CE4E5B9) & M
            x ^= x >> 27; x = (x * 0x94D049BB133111EB) & M
            x ^= x >> 31
            v = int(zv[x >> 48])
            b = bucket_of.get(v)
            if b is None:
                b = bucket_of[v] = fnv_bucket(v)
            out[i, b] = min(int(out[i, b]) + 1, 127)
    return out


def run_reference(args) -> None:
    """--impl reference: the CPU arm.  Rank 0 only; other ranks exit 0 without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import c_oracle as co, rf1
    zb = rf1.zipf_bucket_table()
    n_total = CFG2_ROWS if args.gpus == 1 else CFG4_ROWS
    sample_rows = min(n_total, 4_000_000)
    workload = workload_name(args.gpus)
    F = co.synth_rows(SEED, 0, sample_rows, zb)
    seg = np.zeros(sample_rows, np.uint32)
    Q = np.stack([co.synth_query(SEED, i, zb) for i in range(N_DISTINCT_QUERIES)])
    threads = co.max_threads()
    steps, warmup = args.steps, args.warmup
    # keep the whole run within a few minutes whatever K the caller asked for
    t0 = time.perf_counter()
    co.score_topk_keys(F, seg, Q[0], [0])
    per = max(time.perf_counter() - t0, 1e-4)
    steps_run = max(1, min(steps, int(120.0 / per)))
    for i in range(min(warmup, 5)):
        co.score_topk_keys(F, seg, Q[i % len(Q)], [0])
    t0 = time.perf_counter()
    for i in range(steps_run):
        co.score_topk_keys(F, seg, Q[i % len(Q)], [0])
    dt = time.perf_counter() - t0
    value = sample_rows * steps_run / dt
    sample = (f"each step = 1 query x first {sample_rows} chunks of the {n_total}-chunk corpus; "
              f"{steps_run} of {steps} requested steps timed; RF-1 C oracle ({co.dot_isa()}), OpenMP {threads} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": args.gpus,
            "steps": steps_run, "warmup": min(warmup, 5), "ms_per_step": 1e3 * dt / steps_run, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "s8 x s8 -> s32",
            "data": "synthetic", "config": {"workload": workload, "chunks": n_total, "dim": 256, "k": 10},
            "qps": steps_run / dt * (sample_rows / n_total),
            "cpu_baseline": {"value": value, "unit": "chunks/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "the reference has no local retriever (gemini_rag.py:704-718); this arm is the frozen RF-1 CPU oracle"}
    print(json.dumps(line))


def workload_name(n_gpus: int) -> str:
    if n_gpus == 1:
        return "configs[1]: synthetic single store, 1M chunks, 1 query at a time, top-10 on 1 B200"
    return f"configs[3]: synthetic 100M-chunk corpus sharded by chunk across {n_gpus} B200 with NCCL top-k merge"


def run_b200(args) -> None:
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    from rag_foundation_b200 import Engine
    from rag_foundation_b200.sharded import FusedShardedSearcher, ShardedSearcher, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    n_gpus = max(world, 1)
    if args.gpus != n_gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using {n_gpus}", file=sys.stderr)

    n_total = args.chunks or (CFG2_ROWS if n_gpus == 1 else CFG4_ROWS)
    lo, hi = shard_range(n_total, rank, n_gpus)
    k = 10
    hbm_peak, peak_src = load_peaks()

    eng = Engine(capacity_rows=hi - lo, device=local_rank, id_base=lo)
    seg = eng.open_store("fileSearchStores/bench")
    eng.ingest_synthetic(seg, 0, seed=SEED, start_counter=lo, n_rows=hi - lo)
    searcher = ShardedSearcher.for_engine(eng)          # NCCL all-gather + merge kernel
    exchange = "single GPU"
    if world > 1:
        exchange = "nccl all-gather + merge kernel"
        if args.exchange == "fused":
            fused = None
            try:    # compute + collective in one kernel over NVLink peer memory (symmetric memory)
                fused = FusedShardedSearcher(eng, nq_cap=8, k=10)
            except Exception as exc:   # noqa: BLE001
                print(f"bench.py: rank {rank}: fused exchange unavailable ({exc})", file=sys.stderr)
            ok = torch.tensor([1 if fused is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)          # every rank or none
            if int(ok.item()):
                searcher = fused
                exchange = "fused in the scan kernel: NVLink peer stores + flags (symmetric memory), no collective launch"

    Qh = make_queries(N_DISTINCT_QUERIES)
    Qd = torch.from_numpy(Qh).to(dev)
    stream = torch.cuda.current_stream(dev)
    out_keys = torch.zeros((N_DISTINCT_QUERIES, k), dtype=torch.int64, device=dev)
    # the query batch is resident and complete before the first search: back-to-back scans may overlap
    torch.cuda.synchronize(dev)
    eng.set_stream_overlap(stream.cuda_stream, True)

    def step(i: int):
        qi = i % N_DISTINCT_QUERIES
        q = Qd[qi:qi + 1]
        if n_gpus == 1:
            eng.search_keys_device(q.data_ptr(), 1, [seg], k, out_keys[qi].data_ptr(), stream.cuda_stream)
        else:
            searcher.search_keys(q, [seg], k, out=out_keys[qi:qi + 1])

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # size-independent check at N > 1: the fused NVLink exchange and the NCCL all-gather + merge
    # kernel must return identical packed keys, and the winners' scores must match a CPU re-scoring
    # of exactly those rows (regenerated from their counters by the oracle on rank 0)
    exchange_check = None
    if world > 1:
        nccl_ref = ShardedSearcher.for_engine(eng)
        agree = 1
        first_keys = None
        for qi in range(4):
            a_keys = searcher.search_keys(Qd[qi:qi + 1], [seg], k).clone()
            b_keys = nccl_ref.search_keys(Qd[qi:qi + 1], [seg], k)
            agree &= int(torch.equal(a_keys, b_keys))
            if qi == 0:
                first_keys = a_keys.cpu().numpy().view(np.uint64)[0]
        t_ag = torch.tensor([agree], device=dev)
        dist.all_reduce(t_ag, op=dist.ReduceOp.MIN)
        exchange_check = {"fused_equals_nccl": bool(int(t_ag.item()))}
        if rank == 0:
            from oracle import c_oracle as co, rf1
            zb = rf1.zipf_bucket_table()
            ok = True
            for key in first_keys:
                if int(key) == 0:
                    continue
                gid = 0xFFFFFFFF - (int(key) & 0xFFFFFFFF)
                row = co.synth_rows(SEED, gid, 1, zb)[0]
                ok &= int(row.astype(np.int32) @ Qh[0].astype(np.int32)) == int(key) >> 32
            exchange_check["winner_scores_match_cpu_rescoring"] = bool(ok)

    steps, warmup = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    with sampler:
        for i in range(warmup):
            step(i)
        sync_all()
        launches0 = eng.stats()["kernel_launches"]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for i in range(steps):
            step(i)
        ev1.record(stream)
        sync_all()
        ms = ev0.elapsed_time(ev1)
        launches = eng.stats()["kernel_launches"] - launches0

        # ---- scan kernel alone (the dominant kernel), same stream, CUDA events
        kq = min(steps, 512)
        ek0, ek1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        scratch = torch.zeros((1, k), dtype=torch.int64, device=dev)
        ek0.record(stream)
        for i in range(kq):
            qi = i % N_DISTINCT_QUERIES
            eng.search_keys_device(Qd[qi:qi + 1].data_ptr(), 1, [seg], k, scratch.data_ptr(), stream.cuda_stream)
        ek1.record(stream)
        sync_all()
        kernel_ms = ek0.elapsed_time(ek1) / kq

        # ---- isolated launches (a lone caller's latency: no overlap with a neighbouring query)
        lat_us = []
        for i in range(min(200, max(steps, 20))):
            qi = i % N_DISTINCT_QUERIES
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            a0.record(stream)
            if n_gpus == 1:
                eng.search_keys_device(Qd[qi:qi + 1].data_ptr(), 1, [seg], k, scratch.data_ptr(), stream.cuda_stream)
            else:
                searcher.search_keys(Qd[qi:qi + 1], [seg], k, out=scratch)
            a1.record(stream)
            a1.synchronize()
            lat_us.append(a0.elapsed_time(a1) * 1e3)
        sync_all()

        # ---- e2e: the public C-ABI call with HOST buffers (query H2D + result D2H inside)
        e2e_steps = max(10, min(steps, 2000))
        if n_gpus == 1:
            for i in range(10):
                eng.search(Qh[i:i + 1], [[seg]], k=k)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                qi = i % N_DISTINCT_QUERIES
                eng.search(Qh[qi:qi + 1], [[seg]], k=k)
            e2e_s = time.perf_counter() - t0
            h2d = 256 + 80 + 16 + 16 + 16   # query row + ScanPlan + one extent (lo, hi, tile prefix), 16-byte aligned
            d2h = k * (8 + 4 + 4) + 4
            # the reference serves up to 50 concurrent streams per process (routes/chat.py:40): the
            # same call from 4 host threads (each search owns a context + stream; ctypes drops the GIL)
            n_thr, per_thr = 4, max(50, e2e_steps // 4)

            def _client(tid):
                for i in range(per_thr):
                    qi = (tid * 17 + i) % N_DISTINCT_QUERIES
                    eng.search(Qh[qi:qi + 1], [[seg]], k=k)
            ths = [threading.Thread(target=_client, args=(t,)) for t in range(n_thr)]
            t0 = time.perf_counter()
            [t.start() for t in ths]
            [t.join() for t in ths]
            e2e_conc_qps = n_thr * per_thr / (time.perf_counter() - t0)
        else:
            pinned_q = torch.from_numpy(Qh).pin_memory()
            host_out = torch.zeros((1, k), dtype=torch.int64).pin_memory()
            qbuf = torch.zeros((1, 256), dtype=torch.int8, device=dev)

            def e2e_step(i):
                qi = i % N_DISTINCT_QUERIES
                qbuf.copy_(pinned_q[qi:qi + 1], non_blocking=True)
                keys = searcher.search_keys(qbuf, [seg], k)
                host_out.copy_(keys, non_blocking=True)
                torch.cuda.current_stream(dev).synchronize()
            for i in range(5):
                e2e_step(i)
            sync_all()
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                e2e_step(i)
            sync_all()
            e2e_s = time.perf_counter() - t0
            h2d, d2h = 256, k * 8
            e2e_conc_qps = None

    t = torch.tensor([ms, kernel_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, kernel_ms, e2e_s = (float(x) for x in t.tolist())
    clocks = sampler.summary()

    if rank == 0:
        ms_per_step = ms / steps
        value = n_total * steps / (ms * 1e-3)
        shard_rows = hi - lo
        achieved = shard_rows * BYTES_PER_CHUNK / (kernel_ms * 1e-3) / 1e9
        keys_host = out_keys.cpu().numpy().view(np.uint64)
        cpu = None
        if n_gpus == 1 and not args.no_cpu_baseline:
            cpu = cpu_oracle_leg(n_total, args.cpu_budget_s, Qh, gpu_keys=keys_host if steps + warmup >= N_DISTINCT_QUERIES else None)
        workload = workload_name(n_gpus) if not args.chunks else f"custom: {n_total} chunks over {n_gpus} GPU(s)"
        if world > 1 and isinstance(searcher, FusedShardedSearcher) and searcher.timed_out():
            raise SystemExit("bench.py: a peer's top-k never arrived (fused exchange timed out)")
        line = {
            "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if n_gpus > 1 else "weak",
            "vs_baseline": None, "dtype": "s8 x s8 -> s32", "data": "synthetic",
            "config": {"workload": workload, "chunks": n_total, "chunks_per_gpu": shard_rows, "dim": 256, "k": k,
                       "queries_per_step": 1, "seed": SEED,
                       "l2": f"inputs larger than L2: each step streams {shard_rows * BYTES_PER_CHUNK / 1e6:.0f} MB per GPU (L2 = 126 MB); no flush",
                       "parallelism": "single GPU" if n_gpus == 1 else f"chunk-sharded x{n_gpus}; top-k exchange: {exchange}"},
            "qps": steps / (ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": load_traffic("cfg2" if n_gpus == 1 else "cfg4"),
                         "kernel": "score_topk_scan_kernel", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": shard_rows * BYTES_PER_CHUNK, "peak_source": peak_src,
                         "frac_of_8TBps": achieved / 8000.0},
            "e2e": {"value": n_total * e2e_steps / e2e_s, "unit": "chunks/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "qps": e2e_steps / e2e_s, "ms_per_query": 1e3 * e2e_s / e2e_steps,
                    "steps": e2e_steps, "qps_4_host_threads": e2e_conc_qps, "api": "rf_search (C-ABI, host buffers)" if n_gpus == 1 else "ShardedSearcher.search_keys with pinned host query/result"},
            "isolated_launch_latency_us": {"p50": float(np.percentile(lat_us, 50)), "p95": float(np.percentile(lat_us, 95)),
                                           "n": len(lat_us), "note": "one query at a time with a device sync between launches (device-resident query)"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if exchange_check is not None:
            line["exchange_check"] = exchange_check
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chunks", type=int, default=0, help="override the corpus size (not the headline workload)")
    ap.add_argument("--cpu-budget-s", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"], help="multi-GPU top-k exchange")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
