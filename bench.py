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
host-facing call with HOST buffers: `rf_search` at N = 1, and at N > 1 `rf_group_search` -- ONE process
(rank 0) driving all N GPUs through the engine group that sits behind the adapter (`B200Rag.retrieve`).
Before anything is timed the 64 bench queries are checked against the C oracle (`parity`).
`configs` carries the other BASELINE.json configurations, each with sampled parity and its own roofline:
configs[2] (1024 batched queries, tensor-core path, against an in-run int8 peak probe), configs[4]
(10 k stores x 10 k chunks, 1024 store-scoped queries), ingest featurisation, and -- at N = 1 -- the
100 M-chunk corpus on ONE GPU (`scaling_base`: the same-corpus origin of the 2/4/8-GPU lines).
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
K = 10
N_DISTINCT_QUERIES = 64
METRIC = "chunks scored/s, top-10 store-scoped retrieval (QPS alongside)"
SCAN_KERNEL = "score_topk_scan_tma_kernel<6, 12, 1>"    # as ncu prints it: <consumer warps, ring stages, 256-feature sub-rows per row>


def host_threads() -> int:
    """Cores this process may run on.  torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU legs size
    their OpenMP team from the affinity mask instead and pass it to the oracle explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this workload (profiles/roofline_traffic.json); None when there is none."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:
            return None
    return None


def workload_name(n_gpus: int) -> str:
    if n_gpus == 1:
        return "configs[1]: synthetic single store, 1M chunks, 1 query at a time, top-10 on 1 B200"
    return f"configs[3]: synthetic 100M-chunk corpus sharded by chunk across {n_gpus} B200 with NCCL top-k merge"


def make_config(n_gpus: int, n_total: int, custom: bool = False) -> dict:
    """The `config` object -- built by this one function for BOTH arms, so they are identical key for key."""
    per = n_total // n_gpus
    return {"workload": workload_name(n_gpus) if not custom else f"custom: {n_total} chunks over {n_gpus} GPU(s)",
            "chunks": n_total, "chunks_per_gpu": per, "dim": 256, "k": K, "queries_per_step": 1, "seed": SEED,
            "l2": f"inputs larger than L2: each step streams {per * BYTES_PER_CHUNK / 1e6:.0f} MB per GPU (L2 = 126 MB); no flush",
            "parallelism": "single GPU" if n_gpus == 1 else f"chunk-sharded x{n_gpus}"}


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


def make_queries(n: int, seed: int = SEED, dim: int = 256) -> np.ndarray:
    """RF-1 synthetic queries (oracle/SPEC.md).  Generated with the product's own table so the
    bench never needs the oracle on the GPU arm: same mix64 / bucket rule, vectorised in numpy."""
    from rag_foundation_b200.engine import load_zipf_vocab
    zv = load_zipf_vocab()

    def fnv_bucket(v: int) -> int:
        h = 0x811C9DC5
        for b in str(v).encode():
            h = ((h ^ b) * 0x01000193) & 0xFFFFFFFF
        return h & (dim - 1)

    bucket_of = {}
    M = (1 << 64) - 1
    out = np.zeros((n, dim), np.int8)
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


def make_text(n_bytes: int, seed: int = SEED) -> bytes:
    """Synthetic ASCII document for the ingest leg: Zipf-distributed vocabulary ids rendered as decimal
    words, one space apart (the token rule of the synthetic corpus, oracle/SPEC.md), a line break every 16."""
    from rag_foundation_b200.engine import load_zipf_vocab
    zv = load_zipf_vocab()
    rng = np.random.default_rng(seed)
    n_tok = n_bytes // 3 + 64
    ids = zv[rng.integers(0, 65536, n_tok)]
    words = [str(i).encode() for i in range(int(zv.max()) + 1)]
    parts = [words[i] for i in ids.tolist()]
    lines = [b" ".join(parts[i:i + 16]) for i in range(0, len(parts), 16)]
    return b"\n".join(lines)[:n_bytes]


# ---------------------------------------------------------------------------------------------- CPU legs
def cpu_oracle_leg(n_total: int, budget_s: float, queries: np.ndarray, seed: int = SEED):
    """Time the RF-1 C oracle (all host cores) on a bounded sample: the first min(n_total, 4M) chunks of the
    same synthetic corpus, as many of `queries` as fit in ~budget_s."""
    from oracle import c_oracle as co, rf1
    zb = rf1.zipf_bucket_table()
    threads = host_threads()
    sample_rows = min(n_total, 4_000_000)
    F = co.synth_rows(seed, 0, sample_rows, zb, threads=threads)
    seg = np.zeros(sample_rows, np.uint32)
    co.score_topk_keys(F, seg, queries[0], [0], threads=threads)   # warm
    t0 = time.perf_counter()
    done = 0
    while True:   # cycle through the queries until ~budget_s of CPU work has been timed
        co.score_topk_keys(F, seg, queries[done % len(queries)], [0], threads=threads)
        done += 1
        if time.perf_counter() - t0 > budget_s and done >= min(len(queries), 8):
            break
    dt = time.perf_counter() - t0
    return {"value": sample_rows * done / dt, "unit": "chunks/s", "cores": threads, "kind": "port",
            "sample": f"{done} queries x first {sample_rows} chunks of the workload corpus (seed {seed}), "
                      f"RF-1 C oracle ({co.dot_isa()}, OpenMP {threads} threads), {dt:.2f} s",
            "qps_on_sample": done / dt}


def oracle_parity(n_rows: int, queries: np.ndarray, gpu_keys: np.ndarray, id_base: int = 0, start_counter: int = 0, seed: int = SEED):
    """GPU packed keys vs the C oracle over rows [start_counter, start_counter + n_rows) of the corpus."""
    from oracle import c_oracle as co, rf1
    dim = queries.shape[1]
    zb = rf1.zipf_bucket_table(dim=dim)
    threads = host_threads()
    F = co.synth_rows(seed, start_counter, n_rows, zb, threads=threads, dim=dim)
    seg = np.zeros(n_rows, np.uint32)
    bad = 0
    for i in range(len(queries)):
        want = co.score_topk_keys(F, seg, queries[i], [0], k=gpu_keys.shape[1], id_base=id_base, threads=threads)
        bad += int(want.tolist() != gpu_keys[i].tolist())
    return bad


def run_reference(args) -> None:
    """--impl reference: the CPU arm.  Rank 0 only; other ranks exit 0 without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import c_oracle as co, rf1
    zb = rf1.zipf_bucket_table()
    threads = host_threads()
    n_total = args.chunks or (CFG2_ROWS if args.gpus == 1 else CFG4_ROWS)
    sample_rows = min(n_total, 4_000_000)
    F = co.synth_rows(SEED, 0, sample_rows, zb, threads=threads)
    seg = np.zeros(sample_rows, np.uint32)
    Q = np.stack([co.synth_query(SEED, i, zb) for i in range(N_DISTINCT_QUERIES)])
    steps, warmup = args.steps, max(3, min(args.warmup, 10))
    for i in range(warmup):
        co.score_topk_keys(F, seg, Q[i % len(Q)], [0], threads=threads)
    # exactly `steps` timed steps when that is between 2 s and 2 minutes of work; otherwise as many as that takes
    t0 = time.perf_counter()
    done = 0
    while True:
        co.score_topk_keys(F, seg, Q[done % len(Q)], [0], threads=threads)
        done += 1
        el = time.perf_counter() - t0
        if (done >= steps and el >= 2.0) or el > 120.0:
            break
    dt = time.perf_counter() - t0
    value = sample_rows * done / dt
    sample = (f"each step = 1 query x first {sample_rows} chunks of the {n_total}-chunk corpus; "
              f"{done} steps timed ({steps} requested, >= 2 s); RF-1 C oracle ({co.dot_isa()}), OpenMP {threads} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": args.gpus,
            "steps": done, "warmup": warmup, "ms_per_step": 1e3 * dt / done, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "s8 x s8 -> s32",
            "data": "synthetic", "config": make_config(args.gpus, n_total, custom=bool(args.chunks)),
            "qps": done / dt * (sample_rows / n_total),
            "cpu_baseline": {"value": value, "unit": "chunks/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "the reference has no local retriever (gemini_rag.py:704-718); this arm is the frozen RF-1 CPU oracle"}
    emit(line)


# ---------------------------------------------------------------------------------------------- other configs
def events_ms(torch, stream, fn, reps: int, warm: int = 3) -> float:
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def leg_cfg2(torch, dev, eng, seg, hbm_peak, sample_parity: bool):
    """configs[2]: 1 M chunks x 1024 batched queries on the tensor-core path, device-resident."""
    from rag_foundation_b200.engine import probe_int8_peak
    nq = 1024
    Q = make_queries(nq, seed=SEED + 2)
    qd = torch.from_numpy(Q).to(dev)
    out = torch.zeros((nq, K), dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev)
    l0 = eng.stats()["kernel_launches"]
    eng.search_keys_device(qd.data_ptr(), nq, [seg], K, out.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize(dev)
    launches = eng.stats()["kernel_launches"] - l0
    keys = out.cpu().numpy().view(np.uint64)
    bad = oracle_parity(CFG2_ROWS, Q[::16], keys[::16]) if sample_parity else None
    ms = events_ms(torch, stream, lambda: eng.search_keys_device(qd.data_ptr(), nq, [seg], K, out.data_ptr(), stream.cuda_stream), reps=20)
    peak_ops, peak_ms = probe_int8_peak(dev.index or 0)
    ops = 2.0 * nq * CFG2_ROWS * 256
    # two batches in flight (alternating caller streams): a batch's small kernels (k-th largest, list merges) and the
    # ragged end of its passes overlap the other batch's GEMM
    two = {}
    try:
        s2 = torch.cuda.Stream(dev)
        outs, sts = [out, torch.zeros_like(out)], [stream, s2]
        torch.cuda.synchronize(dev)

        def call(i):
            eng.search_keys_device(qd.data_ptr(), nq, [seg], K, outs[i & 1].data_ptr(), sts[i & 1].cuda_stream)
        for i in range(4):
            call(i)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        s2.wait_event(e0)
        for i in range(20):
            call(i)
        j = torch.cuda.Event()
        j.record(s2)
        stream.wait_event(j)
        e1.record(stream)
        e1.synchronize()
        ms2 = e0.elapsed_time(e1) / 20
        same = bool((outs[0].cpu().numpy().view(np.uint64) == keys).all() and (outs[1].cpu().numpy().view(np.uint64) == keys).all())
        two = {"two_in_flight": {"ms_per_batch": ms2, "qps": nq / (ms2 * 1e-3), "tensor_frac": ops / (ms2 * 1e-3) / peak_ops, "equals_sequential": same}}
    except Exception as exc:   # noqa: BLE001
        two = {"two_in_flight": {"error": f"{type(exc).__name__}: {exc}"}}
    # e2e: the host entry point (queries H2D, keys -> ids/scores/cosines, results D2H inside)
    csr = (np.full(nq, seg, np.uint32), np.arange(nq + 1, dtype=np.uint32))
    for _ in range(3):
        eng.search(Q, csr, k=K)
    t0 = time.perf_counter()
    for _ in range(10):
        eng.search(Q, csr, k=K)
    e2e_ms = (time.perf_counter() - t0) / 10 * 1e3
    return {"workload": "configs[2]: synthetic single store, 1M chunks, batched 1024 queries (GEMM-shaped scoring), top-10 on 1 B200",
            "ms_per_batch": ms, "qps": nq / (ms * 1e-3), "chunks_per_s": nq * CFG2_ROWS / (ms * 1e-3), "launches_per_batch": launches,
            "kernel": "score_topk_gemm_pair_kernel (tcgen05.mma.cta_group::2.kind::i8) + floor pass + merge_lists_kernel",
            "roofline": {"bound": "tensor", "achieved": ops / (ms * 1e-3) / 1e12, "peak": peak_ops / 1e12, "unit": "TOP/s (int8, 2*M*N*K)",
                         "frac": ops / (ms * 1e-3) / peak_ops, "peak_source": f"rf_probe_int8_peak in this run ({peak_ms:.2f} ms of back-to-back 256x256x32 pair MMAs)",
                         "hbm_GBps": CFG2_ROWS * BYTES_PER_CHUNK / (ms * 1e-3) / 1e9, "hbm_frac": CFG2_ROWS * BYTES_PER_CHUNK / (ms * 1e-3) / 1e9 / hbm_peak},
            "e2e_ms_per_batch": e2e_ms, "e2e_qps": nq / (e2e_ms * 1e-3), "h2d_bytes": nq * 256, "d2h_bytes": nq * (K * 16 + 4),
            "parity_mismatches": bad, "parity_checked": len(Q[::16]) if sample_parity else 0, **two}


def cfg4_parity(ids, sc, Q, scopes, per_store, id_of_store_row0, n_check=32):
    from oracle import c_oracle as co, rf1
    zb = rf1.zipf_bucket_table()
    bad = 0
    for i in range(min(n_check, len(scopes))):   # oracle on the scoped store only (regenerated from its counters)
        st = scopes[i][0]
        F = co.synth_rows(SEED + 4, st * per_store, per_store, zb)
        w_ids, w_sc, _ = co.score_topk(F, np.zeros(per_store, np.uint32), Q[i], [0], id_base=id_of_store_row0(st))
        bad += int(ids[i].tolist() != w_ids.tolist() or sc[i].tolist() != w_sc.tolist())
    return bad


def leg_cfg4(torch, dev, hbm_peak, n_stores, per_store, sample_parity: bool, devices=None):
    """configs[4]: n_stores x per_store chunks, 1024 store-scoped queries per batch.  devices = None: one engine
    on `dev`; else an engine group over `devices` with whole stores per GPU (rf_group_search, host buffers)."""
    from rag_foundation_b200 import Engine, EngineGroup
    from rag_foundation_b200.engine import scopes_to_csr
    nq = 1024
    n_rows = n_stores * per_store
    rng = np.random.default_rng(5)
    Q = make_queries(nq, seed=SEED + 4)
    scopes = [[int(rng.integers(0, n_stores))] for _ in range(nq)]
    csr = scopes_to_csr(scopes)
    alg = nq * per_store * BYTES_PER_CHUNK
    out = {"workload": f"configs[4]: multi-tenant, {n_stores} stores x {per_store} chunks, {nq} store-scoped queries per batch",
           "algorithmic_bytes_per_batch": alg, "kernel": SCAN_KERNEL + " (grid.y = queries, per-query extents + tenant mask)"}
    if devices is None:
        eng = Engine(capacity_rows=n_rows, device=dev.index or 0)
        id0 = lambda st: st * per_store   # noqa: E731
    else:
        G = len(devices)
        eng = EngineGroup(devices, capacity_rows=(n_stores + G - 1) // G * per_store, placement="store")
        stride = 0xFFFFFFFE // G
        id0 = lambda st: (st % G) * stride + (st // G) * per_store   # noqa: E731
        out["parallelism"] = f"whole stores per GPU x{G} (store g on GPU g % {G}), one process, host-merged"
    try:
        for i in range(n_stores):
            eng.open_store(f"fileSearchStores/mt{i}")
        eng.ingest_synthetic(0, per_store, seed=SEED + 4, start_counter=0, n_rows=n_rows)
        ids, sc, cs, cnt = eng.search(Q, csr, k=K)
        out["parity_mismatches"] = cfg4_parity(ids, sc, Q, scopes, per_store, id0) if sample_parity else None
        out["parity_checked"] = 32 if sample_parity else 0
        for _ in range(3):
            eng.search(Q, csr, k=K)
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.search(Q, csr, k=K)
        e2e_ms = (time.perf_counter() - t0) / reps * 1e3
        out.update({"e2e_ms_per_batch": e2e_ms, "e2e_qps": nq / (e2e_ms * 1e-3), "e2e_GBps": alg / (e2e_ms * 1e-3) / 1e9,
                    "h2d_bytes": nq * 256 + csr[0].nbytes + csr[1].nbytes, "d2h_bytes": nq * (K * 16 + 4),
                    "api": "rf_search (host buffers)" if devices is None else "rf_group_search (host buffers)"})
        n_dev = 1 if devices is None else len(devices)
        out["roofline_e2e"] = {"bound": "hbm", "achieved": alg / (e2e_ms * 1e-3) / 1e9, "peak": hbm_peak * n_dev, "unit": "GB/s",
                               "frac": alg / (e2e_ms * 1e-3) / 1e9 / (hbm_peak * n_dev)}
        if devices is None:   # the kernel alone, device-resident queries and keys
            qd = torch.from_numpy(Q).to(dev)
            keys = torch.zeros((nq, K), dtype=torch.int64, device=dev)
            stream = torch.cuda.current_stream(dev)
            ms = events_ms(torch, stream, lambda: eng.search_keys_device_scoped(qd.data_ptr(), nq, csr, K, keys.data_ptr(), stream.cuda_stream), reps=20)
            out.update({"ms_per_batch": ms, "qps": nq / (ms * 1e-3),
                        "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": alg / (ms * 1e-3) / 1e9 / hbm_peak, "traffic": load_traffic("cfg4")}})
    finally:
        eng.close()
    return out


def leg_cfg4_spmd(torch, dist, dev, rank, world, hbm_peak, n_stores, per_store, sample_parity: bool):
    """configs[4] sharded by whole stores, one process per GPU (all ranks call this): every rank holds the stores
    g with g % world == rank, receives the whole 1024-query batch (device-resident), scans what it owns and
    exchanges inside the kernels (publish-only scan + merge_wait over NVLink peer memory).  CUDA events, max
    over ranks."""
    from rag_foundation_b200 import Engine
    from rag_foundation_b200.sharded import FusedStoreShardedSearcher, StoreShardedSearcher
    nq = 1024
    rng = np.random.default_rng(5)
    Q = make_queries(nq, seed=SEED + 4)
    scopes = [[int(rng.integers(0, n_stores))] for _ in range(nq)]
    owned = (n_stores + world - 1 - rank) // world
    base = StoreShardedSearcher.id_base_for(rank, world)
    eng, srch, err = None, None, ""
    try:   # set-up can fail on one rank only (memory, symmetric-memory rendezvous): agree before any collective step
        eng = Engine(capacity_rows=owned * per_store, device=dev.index or 0, id_base=base)
    except Exception as exc:   # noqa: BLE001
        err = f"{type(exc).__name__}: {exc}"
    ok = torch.tensor([0 if err else 1], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if not int(ok.item()):
        if eng is not None:
            eng.close()
        return {"spmd_error": err or "another rank could not set up"}
    try:
        srch = FusedStoreShardedSearcher(eng, nq_cap=nq, k=K)
        for g in range(n_stores):
            srch.open_store(f"fileSearchStores/mt{g}")
        for g in range(rank, n_stores, world):
            eng.ingest_synthetic(srch.local_seg[g], 0, seed=SEED + 4, start_counter=g * per_store, n_rows=per_store)
        qd = torch.from_numpy(Q).to(dev)
        local = srch.prepare_fused(scopes)
        out = torch.zeros((nq, K), dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev)
        torch.cuda.synchronize(dev)
        eng.set_stream_overlap(stream.cuda_stream, True)
        srch.search_keys(qd, local, K, out=out)
        torch.cuda.synchronize(dev)
        dist.barrier()
        keys = out.cpu().numpy().view(np.uint64)
        bad = None
        if rank == 0 and sample_parity:
            ids = np.where(keys != 0, np.uint64(0xFFFFFFFF) - (keys & np.uint64(0xFFFFFFFF)), np.uint64(0xFFFFFFFFFFFFFFFF))
            sc = (keys >> np.uint64(32)).astype(np.int64).astype(np.int32)
            stride = (1 << 32) // world
            bad = cfg4_parity(ids, sc, Q, scopes, per_store, lambda st: (st % world) * stride + (st // world) * per_store)
        for _ in range(3):
            srch.search_keys(qd, local, K, out=out)
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            srch.search_keys(qd, local, K, out=out)
        e1.record(stream)
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        timed_out = srch.timed_out()
        again = out.cpu().numpy().view(np.uint64)
        alg = nq * per_store * BYTES_PER_CHUNK
        # The same batches with TWO in flight per rank (alternating caller streams, as a server that keeps receiving
        # batches would): a batch's plan upload, scan start-up and exchange wait overlap its neighbour's scan.  Every rank
        # alternates identically, so a gather slot (4 per rank, indexed by the call number) is still free when it comes round.
        two = {}
        try:
            s2 = torch.cuda.Stream(dev)
            eng.set_stream_overlap(s2.cuda_stream, True)
            pair_streams, outs = [stream, s2], [out, torch.zeros_like(out)]

            def call(i):
                with torch.cuda.stream(pair_streams[i & 1]):
                    srch.search_keys(qd, local, K, out=outs[i & 1])
            for i in range(4):
                call(i)
            torch.cuda.synchronize(dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
            reps2 = 20
            e0.record(stream)
            s2.wait_event(e0)
            for i in range(reps2):
                call(i)
            j = torch.cuda.Event()
            j.record(s2)
            stream.wait_event(j)
            e1.record(stream)
            e1.synchronize()
            t2 = torch.tensor([e0.elapsed_time(e1) / reps2], dtype=torch.float64, device=dev)
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            ms2 = float(t2.item())
            same2 = bool((outs[1].cpu().numpy().view(np.uint64) == keys).all() and (outs[0].cpu().numpy().view(np.uint64) == keys).all())
            two = {"spmd_two_in_flight": {"ms_per_batch": ms2, "qps": nq / (ms2 * 1e-3), "equals_sequential": same2 and not srch.timed_out(),
                                          "roofline": {"bound": "hbm", "achieved": alg / (ms2 * 1e-3) / 1e9, "peak": hbm_peak * world, "unit": "GB/s",
                                                       "frac": alg / (ms2 * 1e-3) / 1e9 / (hbm_peak * world)}}}
        except Exception as exc:   # noqa: BLE001
            two = {"spmd_two_in_flight": {"error": f"{type(exc).__name__}: {exc}"}}
        # weak-scaling point: a batch of 1024 x world queries (the per-GPU work of the single-GPU batch): what is left of
        # the fixed per-batch costs (plan upload, two launches, the exchange) once a rank has 0.4 ms of scanning again
        weak = {}
        try:
            nq_w = nq * world
            Qw = np.concatenate([Q] * world)
            scopes_w = [[int(rng.integers(0, n_stores))] for _ in range(nq_w)]
            srch_w = FusedStoreShardedSearcher(eng, nq_cap=nq_w, k=K)
            srch_w.by_name, srch_w.local_seg = srch.by_name, srch.local_seg
            qw = torch.from_numpy(Qw).to(dev)
            local_w = srch_w.prepare_fused(scopes_w)
            out_w = torch.zeros((nq_w, K), dtype=torch.int64, device=dev)
            for _ in range(3):
                srch_w.search_keys(qw, local_w, K, out=out_w)
            torch.cuda.synchronize(dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
            e0.record(stream)
            for _ in range(10):
                srch_w.search_keys(qw, local_w, K, out=out_w)
            e1.record(stream)
            e1.synchronize()
            tw = torch.tensor([e0.elapsed_time(e1) / 10], dtype=torch.float64, device=dev)
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            ms_w = float(tw.item())
            alg_w = nq_w * per_store * BYTES_PER_CHUNK
            weak = {"spmd_weak": {"queries_per_batch": nq_w, "ms_per_batch": ms_w, "qps": nq_w / (ms_w * 1e-3),
                                  "roofline": {"bound": "hbm", "achieved": alg_w / (ms_w * 1e-3) / 1e9, "peak": hbm_peak * world, "unit": "GB/s",
                                               "frac": alg_w / (ms_w * 1e-3) / 1e9 / (hbm_peak * world)},
                                  "timed_out": srch_w.timed_out()}}
        except Exception as exc:   # noqa: BLE001
            weak = {"spmd_weak": {"error": f"{type(exc).__name__}: {exc}"}}
        return {**weak, **two, "spmd_ms_per_batch": ms, "spmd_qps": nq / (ms * 1e-3), "spmd_parity_mismatches": bad, "spmd_stable": bool((again == keys).all()) and not timed_out,
                "spmd_exchange": "publish-only scan + merge_wait over NVLink peer memory (rf_search_keys_device_scoped_fused), plans from the device-resident store table",
                "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": hbm_peak * world, "unit": "GB/s",
                             "frac": alg / (ms * 1e-3) / 1e9 / (hbm_peak * world)}}
    finally:
        eng.close()


def leg_ingest(torch, dev, hbm_peak, sample_parity: bool):
    """Ingest featurisation throughput: one 22.8 MB synthetic document through rf_ingest_text (host buffer)."""
    from rag_foundation_b200 import Engine
    data = make_text(22_800_000)
    reps = 5
    with Engine(capacity_rows=(len(data) // 400 + 64) * (3 * reps + 8) * 2, device=dev.index or 0) as e:
        s = e.open_store("fileSearchStores/ingest")
        first, n_chunks, spans = e.ingest_text(s, 1, data)
        ok = None
        if sample_parity:
            from oracle import c_oracle as co
            wF, wff, wsp, _ = co.featurize_doc(data)
            F, sg, ff = e.read_rows(0, n_chunks)
            ok = bool(n_chunks == len(wF) and (F == wF).all() and (ff == wff).all() and (spans == wsp).all())
        e.ingest_text(s, 2, data, want_spans=False)
        t0 = time.perf_counter()
        for r in range(reps):
            e.ingest_text(s, 3 + r, data, want_spans=False)
        wall = (time.perf_counter() - t0) / reps
        alg = len(data) + n_chunks * 264
        out = {"workload": f"ingest featurisation: one {len(data) / 1e6:.1f} MB synthetic ASCII document (upload cap 25 MB, config.py:118)",
               "n_chunks": n_chunks, "e2e_ms_per_doc": wall * 1e3, "text_GBps_e2e": len(data) / wall / 1e9, "chunks_per_s": n_chunks / wall,
               "h2d_bytes": len(data), "algorithmic_bytes": alg, "parity_ok": ok,
               "api": "rf_ingest_text (pageable host buffer: staging + H2D + featurise kernels + sync inside the timed region)"}
        # the same document from pinned host memory (the caller read the upload straight into an rf_host_alloc buffer)
        from rag_foundation_b200.engine import PinnedBuffer
        pb = PinnedBuffer(len(data))
        pb.array[:] = np.frombuffer(data, np.uint8)
        e.ingest_text_ptr(s, 50, pb.ptr, len(data))
        t0 = time.perf_counter()
        for r in range(reps):
            e.ingest_text_ptr(s, 51 + r, pb.ptr, len(data))
        pin = (time.perf_counter() - t0) / reps
        pb.close()
        out.update({"pinned_ms_per_doc": pin * 1e3, "text_GBps_pinned_source": len(data) / pin / 1e9})
        # and with the text already in HBM: the featurise kernels + the call's two stream synchronisations
        dd = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
        torch.cuda.synchronize(dev)
        e.ingest_text_ptr(s, 100, dd.data_ptr(), len(data))
        k0 = e.stats()["ingest_kernel_ns"]
        t0 = time.perf_counter()
        for r in range(reps):
            e.ingest_text_ptr(s, 101 + r, dd.data_ptr(), len(data))
        kern = (time.perf_counter() - t0) / reps
        kern_dev = (e.stats()["ingest_kernel_ns"] - k0) / reps * 1e-9     # CUDA events on the ingest stream, inside the engine
        out.update({"resident_ms_per_doc": kern * 1e3, "text_GBps_resident": len(data) / kern / 1e9,
                    "kernel_us_per_doc": kern_dev * 1e6,
                    "kernels": "tokenize_span_kernel (single pass: one contiguous span per CTA in shared memory, token counts summed across the wave) + rows_from_tokens_kernel",
                    "roofline": {"bound": "hbm", "achieved": alg / kern_dev / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": alg / kern_dev / 1e9 / hbm_peak,
                                 "note": "text already in HBM; the two featurise kernels timed with CUDA events on the engine's ingest stream (byte work: instruction-bound, not HBM-bound); resident_ms_per_doc adds the call's copies and synchronisations"}})
    return out


def leg_scaling_base(torch, dev, hbm_peak, steps: int):
    """The 100 M-chunk corpus of configs[3] on ONE GPU: the same-corpus origin of the 2/4/8-GPU curve."""
    from rag_foundation_b200 import Engine
    Qh = make_queries(8)
    with Engine(capacity_rows=CFG4_ROWS, device=dev.index or 0) as e:
        s = e.open_store("fileSearchStores/bench")
        e.ingest_synthetic(s, 0, seed=SEED, start_counter=0, n_rows=CFG4_ROWS)
        qd = torch.from_numpy(Qh).to(dev)
        out = torch.zeros((8, K), dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev)
        torch.cuda.synchronize(dev)
        e.set_stream_overlap(stream.cuda_stream, True)
        i = [0]

        def one():
            j = i[0] % 8
            i[0] += 1
            e.search_keys_device(qd[j:j + 1].data_ptr(), 1, [s], K, out[j].data_ptr(), stream.cuda_stream)
        ms = events_ms(torch, stream, one, reps=max(8, min(steps, 50)))
        for j in range(3):
            e.search(Qh[j:j + 1], [[s]], k=K)
        t0 = time.perf_counter()
        for j in range(10):
            e.search(Qh[j % 8:j % 8 + 1], [[s]], k=K)
        e2e_ms = (time.perf_counter() - t0) / 10 * 1e3
        gbs = CFG4_ROWS * BYTES_PER_CHUNK / (ms * 1e-3) / 1e9
        return {"workload": "configs[3] corpus (100M chunks) on ONE B200: same-corpus origin of the 2/4/8-GPU lines", "n_gpus": 1,
                "ms_per_step": ms, "value": CFG4_ROWS / (ms * 1e-3), "unit": "chunks/s",
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                             "traffic": load_traffic(f"shard_{CFG4_ROWS}")},
                "e2e_ms_per_query": e2e_ms, "e2e_value": CFG4_ROWS / (e2e_ms * 1e-3)}


def leg_wide(torch, dev, hbm_peak, steps: int, sample_parity: bool, dim: int = 1024):
    """Wider rows: the configs[1] workload with D = 512 / 1024 features per chunk -- 1 M chunks, one
    query at a time, top-10 -- on the same scan kernel (template on the row width), D + 4 algorithmic bytes per chunk;
    and the configs[2] batch at that width on the streamed-K tensor-core kernel."""
    from rag_foundation_b200 import Engine
    n = CFG2_ROWS
    Qh = make_queries(8, seed=SEED + 6, dim=dim)
    with Engine(capacity_rows=n, device=dev.index or 0, dim=dim) as e:
        s = e.open_store("fileSearchStores/wide")
        e.ingest_synthetic(s, 0, seed=SEED, start_counter=0, n_rows=n)
        qd = torch.from_numpy(Qh).to(dev)
        out = torch.zeros((8, K), dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev)
        torch.cuda.synchronize(dev)
        e.set_stream_overlap(stream.cuda_stream, True)
        i = [0]

        def one():
            j = i[0] % 8
            i[0] += 1
            e.search_keys_device(qd[j:j + 1].data_ptr(), 1, [s], K, out[j].data_ptr(), stream.cuda_stream)
        ms = events_ms(torch, stream, one, reps=max(16, min(steps, 200)), warm=8)
        torch.cuda.synchronize(dev)
        bad = oracle_parity(n, Qh, out.cpu().numpy().view(np.uint64)) if sample_parity else None
        for j in range(3):
            e.search(Qh[j:j + 1], [[s]], k=K)
        t0 = time.perf_counter()
        for j in range(20):
            e.search(Qh[j % 8:j % 8 + 1], [[s]], k=K)
        e2e_ms = (time.perf_counter() - t0) / 20 * 1e3
        bytes_per_launch = n * (dim + 4)
        gbs = bytes_per_launch / (ms * 1e-3) / 1e9
        # the batched form (configs[2] at this width): 1024 queries on the streamed-K tensor-core kernel
        nqb = 1024
        Qb = make_queries(nqb, seed=SEED + 7, dim=dim)
        qbd = torch.from_numpy(Qb).to(dev)
        outb = torch.zeros((nqb, K), dtype=torch.int64, device=dev)
        l0 = e.stats()["kernel_launches"]
        e.search_keys_device(qbd.data_ptr(), nqb, [s], K, outb.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize(dev)
        launches_b = e.stats()["kernel_launches"] - l0
        bad_b = oracle_parity(n, Qb[::64], outb.cpu().numpy().view(np.uint64)[::64]) if sample_parity else None
        ms_b = events_ms(torch, stream, lambda: e.search_keys_device(qbd.data_ptr(), nqb, [s], K, outb.data_ptr(), stream.cuda_stream), reps=10)
        from rag_foundation_b200.engine import probe_int8_peak
        peak_ops, _ = probe_int8_peak(dev.index or 0)
        ops_b = 2.0 * nqb * n * dim
        batched = {"workload": f"configs[2] with wider rows: 1M chunks x {dim} features, batched 1024 queries, top-10 on 1 B200",
                   "kernel": f"score_topk_gemm_wide_kernel (tcgen05.mma.cta_group::2.kind::i8, K = {dim} streamed in {dim // 256} slabs) + floor pass + kth_largest + merge_lists",
                   "ms_per_batch": ms_b, "qps": nqb / (ms_b * 1e-3), "launches_per_batch": launches_b,
                   "roofline": {"bound": "hbm (256 queries per CTA pair: one pass over the features per 256 queries, mostly from L2 after the first)",
                                "achieved": (nqb // 256) * n * dim / (ms_b * 1e-3) / 1e9, "peak": hbm_peak, "unit": f"GB/s of feature reads (4 passes x {n * dim / 1e9:.2f} GB; L2 serves part)",
                                "frac": (nqb // 256) * n * dim / (ms_b * 1e-3) / 1e9 / hbm_peak,
                                "tensor_TOPs": ops_b / (ms_b * 1e-3) / 1e12, "tensor_frac": ops_b / (ms_b * 1e-3) / peak_ops},
                   "parity_mismatches": bad_b, "parity_checked": len(Qb[::64]) if sample_parity else 0}
        return {"batched": batched, "workload": f"configs[1] with wider rows: 1M chunks x {dim} int8 features, 1 query at a time, top-10 on 1 B200", "dim": dim,
                "kernel": "score_topk_scan_tma_kernel<6, 12, %d>" % (dim // 256), "ms_per_query": ms, "qps": 1e3 / ms, "chunks_per_s": n / (ms * 1e-3),
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "frac_of_8TBps": gbs / 8000.0,
                             "traffic": load_traffic("wide1024") if dim == 1024 else None},
                "e2e_ms_per_query": e2e_ms, "api": "rf_search (host buffers)", "parity_mismatches": bad, "parity_checked": 8 if sample_parity else 0}


# ---------------------------------------------------------------------------------------------- the B200 arm
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    from rag_foundation_b200 import Engine, EngineGroup
    from rag_foundation_b200.sharded import FusedShardedSearcher, ShardedSearcher, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    host_pg = None
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_pg = dist.new_group(backend="gloo")     # host-side waits (a NCCL barrier would spin on the GPUs rank 0 is timing)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    n_gpus = max(world, 1)
    if args.gpus != n_gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using {n_gpus}", file=sys.stderr)

    n_total = args.chunks or (CFG2_ROWS if n_gpus == 1 else CFG4_ROWS)
    lo, hi = shard_range(n_total, rank, n_gpus)
    k = K
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
    # N = 1: the device-resident searches alternate between TWO caller streams (what a server with several requests
    # in flight does; `e2e` below does the same through rf_search_begin / rf_search_end).  A stream's searches still
    # follow each other under programmatic dependent launch; the other stream's blocks are already queued on the
    # device and take an SM the moment one of this query's blocks retires, so no SM waits for the slowest block of
    # the query before it.  --streams 1 is the single-stream schedule (reported beside it either way).
    streams = [stream]
    if n_gpus == 1:
        for _ in range(max(1, args.streams) - 1):
            s2 = torch.cuda.Stream(dev)
            eng.set_stream_overlap(s2.cuda_stream, True)
            streams.append(s2)

    # (addresses and stream handles looked up once: inside the timed region a step is the C-ABI call and nothing else)
    q_ptrs = [Qd[qi:qi + 1].data_ptr() for qi in range(N_DISTINCT_QUERIES)]
    out_ptrs = [out_keys[qi].data_ptr() for qi in range(N_DISTINCT_QUERIES)]
    stream_handles = [st.cuda_stream for st in streams]
    scope = [seg]

    def step(i: int, n_streams: int = 0):
        qi = i % N_DISTINCT_QUERIES
        if n_gpus == 1:
            eng.search_keys_device(q_ptrs[qi], 1, scope, k, out_ptrs[qi], stream_handles[i % (n_streams or len(streams))])
        else:
            searcher.search_keys(Qd[qi:qi + 1], scope, k, out=out_keys[qi:qi + 1])

    def timed_region(n: int, n_streams: int = 0) -> float:
        """ms for steps 0 .. n-1: CUDA events on the first stream, which the other streams start behind and which
        joins them again before the closing event."""
        use = streams[:(n_streams or len(streams))]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(use[0])
        for o in use[1:]:
            o.wait_event(e0)
        for i in range(n):
            step(i, len(use))
        for o in use[1:]:
            j = torch.cuda.Event()
            j.record(o)
            use[0].wait_event(j)
        e1.record(use[0])
        e1.synchronize()
        return e0.elapsed_time(e1)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- parity BEFORE anything is timed (whatever --steps is): all 64 bench queries through the timed path
    for i in range(N_DISTINCT_QUERIES):
        step(i)
    sync_all()
    keys_host = out_keys.cpu().numpy().view(np.uint64).copy()
    parity = {"queries": N_DISTINCT_QUERIES}
    if n_gpus == 1:
        if rank == 0 and not args.no_parity:
            parity["mismatches"] = oracle_parity(n_total, Qh, keys_host)
            parity["against"] = f"RF-1 C oracle over all {n_total} chunks (ids, int32 scores, tie order)"
    else:
        # size-independent checks at N > 1: fused NVLink exchange == NCCL all-gather + merge kernel on every
        # query, every rank's LOCAL top-10 == the oracle over that rank's shard for the first 4 queries (the CPU
        # cannot score 100 M chunks per query in bench time), and the merged winners re-scored on the CPU
        nccl_ref = ShardedSearcher.for_engine(eng)
        agree = 1
        for qi in range(N_DISTINCT_QUERIES):
            b_keys = nccl_ref.search_keys(Qd[qi:qi + 1], [seg], k)
            agree &= int(torch.equal(out_keys[qi:qi + 1], b_keys))
        shard_bad = 0
        if not args.no_parity:
            local = torch.zeros((4, k), dtype=torch.int64, device=dev)
            for qi in range(4):
                eng.search_keys_device(Qd[qi:qi + 1].data_ptr(), 1, [seg], k, local[qi].data_ptr(), stream.cuda_stream)
            torch.cuda.synchronize(dev)
            shard_bad = oracle_parity(hi - lo, Qh[:4], local.cpu().numpy().view(np.uint64), id_base=lo, start_counter=lo)
        t_ag = torch.tensor([agree, -shard_bad], device=dev)
        dist.all_reduce(t_ag, op=dist.ReduceOp.MIN)
        parity.update({"fused_equals_nccl": bool(int(t_ag[0].item())), "shard_local_mismatches_max_over_ranks": -int(t_ag[1].item()),
                       "against": "NCCL path on all 64 queries; C oracle over each rank's own shard (4 queries); winners re-scored on the CPU"})
        if rank == 0:
            from oracle import c_oracle as co, rf1
            zb = rf1.zipf_bucket_table()
            ok = True
            for qi in range(4):
                for key in keys_host[qi]:
                    if int(key) == 0:
                        continue
                    gid = 0xFFFFFFFF - (int(key) & 0xFFFFFFFF)
                    row = co.synth_rows(SEED, gid, 1, zb)[0]
                    ok &= int(row.astype(np.int32) @ Qh[qi].astype(np.int32)) == int(key) >> 32
            parity["winner_scores_match_cpu_rescoring"] = bool(ok)
            parity["mismatches"] = (0 if parity["fused_equals_nccl"] and ok and parity["shard_local_mismatches_max_over_ranks"] == 0 else 1)

    steps, warmup = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    e2e_steps = max(200 if n_gpus == 1 else 10, min(steps, 2000))   # e2e.steps says how many; N = 1: at least 200 (8 ms)
    e2e_s, e2e_conc_qps, h2d, d2h, api = None, None, 0, 0, ""
    e2e_seq_s = None
    with sampler:
        for i in range(warmup):
            step(i)
        sync_all()
        launches0 = eng.stats()["kernel_launches"]
        ms = timed_region(steps)
        sync_all()
        launches = eng.stats()["kernel_launches"] - launches0

        # ---- scan kernel alone (the dominant kernel): the same launches without the exchange, CUDA events on the
        # launching stream(s); at N = 1 also the single-stream schedule
        kq = 512 if n_gpus == 1 else min(steps, 512)       # (N = 1: a region long enough that its two ends do not show: 21 ms)
        scratch = torch.zeros((1, k), dtype=torch.int64, device=dev)
        single_stream_ms = None
        if n_gpus == 1:
            kernel_ms = timed_region(kq) / kq
            sync_all()
            if len(streams) > 1:
                single_stream_ms = timed_region(kq, 1) / kq
                sync_all()
        else:
            ek0, ek1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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

        # ---- e2e at N = 1: the public C-ABI call with HOST buffers (query H2D + result D2H inside)
        if n_gpus == 1:
            for i in range(10):
                eng.search(Qh[i:i + 1], [[seg]], k=k)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                qi = i % N_DISTINCT_QUERIES
                eng.search(Qh[qi:qi + 1], [[seg]], k=k)
            e2e_seq_s = time.perf_counter() - t0
            # The same public call in its two halves (rf_search_begin / rf_search_end), two searches in flight from ONE
            # host thread: a query's host-side work and result delivery overlap the next query's scan.  Every step still
            # sends its query from a host buffer and reads its result back; results are taken in order.
            e2e_check = []
            for depth_warm in range(4):
                eng.search_begin(Qh[depth_warm:depth_warm + 1], [[seg]], k).result()
            t0 = time.perf_counter()
            pend = []
            for i in range(e2e_steps):
                qi = i % N_DISTINCT_QUERIES
                pend.append(eng.search_begin(Qh[qi:qi + 1], [[seg]], k))
                if len(pend) == 2:
                    r = pend.pop(0).result()
                    if i < 8:
                        e2e_check.append(r)
            while pend:
                pend.pop(0).result()
            e2e_s = time.perf_counter() - t0
            for j, r in enumerate(e2e_check):          # the pipelined answers are the blocking call's answers
                w = eng.search(Qh[j:j + 1], [[seg]], k=k)
                if not all((x == y).all() for x, y in zip(r, w)):
                    raise SystemExit("bench.py: pipelined rf_search_begin / rf_search_end differs from rf_search")
            h2d = 256 + 80 + 16 + 16 + 16   # query row + scan plan + one extent, all in the kernel's parameter block
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
            api = ("rf_search_begin / rf_search_end (C-ABI, host buffers), 2 searches in flight from one host thread, each on its own context and "
                   "stream: one query's merge tail and result delivery overlap the next query's scan (which is why this can exceed the "
                   "single-stream device-timed value); sequential_ms_per_query = the blocking rf_search, one caller, nothing in flight")

    t = torch.tensor([ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, kernel_ms = (float(x) for x in t.tolist())
    clocks = sampler.summary()
    if world > 1 and isinstance(searcher, FusedShardedSearcher) and searcher.timed_out():
        raise SystemExit("bench.py: a peer's top-k never arrived (fused exchange timed out)")
    shard_rows = hi - lo
    eng.close()
    del searcher
    cfg4_spmd = None
    if world > 1 and not args.no_configs:
        try:
            cfg4_spmd = leg_cfg4_spmd(torch, dist, dev, rank, world, hbm_peak, 10_000, 10_000, not args.no_parity)
        except Exception as exc:   # noqa: BLE001
            cfg4_spmd = {"spmd_error": f"{type(exc).__name__}: {exc}"}

    # ---- N > 1: the host-facing legs run in ONE process (rank 0) that drives all N GPUs through the engine
    # group behind the adapter; the other ranks have released their engines and wait on the host
    configs = {}
    cpu = None
    if world > 1:
        torch.cuda.synchronize(dev)
        dist.barrier(group=host_pg)
    if rank == 0:
        if n_gpus > 1:
            bases = [shard_range(n_total, d, n_gpus)[0] for d in range(n_gpus)]
            cap = max(shard_range(n_total, d, n_gpus)[1] - bases[d] for d in range(n_gpus))
            with EngineGroup(list(range(n_gpus)), capacity_rows=cap, placement="spread", id_bases=bases) as grp:
                gs = grp.open_store("fileSearchStores/bench")
                grp.ingest_synthetic(gs, 0, seed=SEED, start_counter=0, n_rows=n_total)
                g_bad = 0
                for qi in range(N_DISTINCT_QUERIES):   # the group's host-merged answer == the fused SPMD answer, key for key
                    ids, sc, _, cnt = grp.search(Qh[qi:qi + 1], [[gs]], k=k)
                    got = (sc[0].astype(np.int64).astype(np.uint64) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - (ids[0] & np.uint64(0xFFFFFFFF)))
                    g_bad += int(got.tolist() != keys_host[qi].tolist())
                parity["group_vs_spmd_mismatches"] = g_bad
                parity["mismatches"] = int(bool(parity.get("mismatches", 0)) or g_bad > 0)
                for i in range(10):
                    grp.search(Qh[i:i + 1], [[gs]], k=k)
                t0 = time.perf_counter()
                for i in range(e2e_steps):
                    qi = i % N_DISTINCT_QUERIES
                    grp.search(Qh[qi:qi + 1], [[gs]], k=k)
                e2e_s = time.perf_counter() - t0
                h2d, d2h = n_gpus * (256 + 80 + 48), n_gpus * (k * 16 + 4)
                api = f"rf_group_search: one process, {n_gpus} engines (what B200Rag.retrieve calls with RAG_B200_DEVICES), host buffers, host-merged top-k; one caller, nothing in flight (4 callers: qps_4_host_threads)"
                n_thr, per_thr = 4, max(25, e2e_steps // 8)

                def _gclient(tid):
                    for i in range(per_thr):
                        qi = (tid * 17 + i) % N_DISTINCT_QUERIES
                        grp.search(Qh[qi:qi + 1], [[gs]], k=k)
                ths = [threading.Thread(target=_gclient, args=(t,)) for t in range(n_thr)]
                t0 = time.perf_counter()
                [t.start() for t in ths]
                [t.join() for t in ths]
                e2e_conc_qps = n_thr * per_thr / (time.perf_counter() - t0)
        if not args.no_cpu_baseline:
            cpu = cpu_oracle_leg(n_total, args.cpu_budget_s, Qh)
        if not args.no_configs:
            sp = not args.no_parity
            try:
                legs = set(args.legs.split(","))
                if n_gpus == 1:
                    if "cfg2" in legs:
                        with Engine(capacity_rows=CFG2_ROWS, device=local_rank) as e2:
                            s2 = e2.open_store("fileSearchStores/cfg2")
                            e2.ingest_synthetic(s2, 0, seed=SEED, start_counter=0, n_rows=CFG2_ROWS)
                            configs["cfg2"] = leg_cfg2(torch, dev, e2, s2, hbm_peak, sp)
                    if "ingest" in legs:
                        configs["ingest"] = leg_ingest(torch, dev, hbm_peak, sp)
                    if "cfg4" in legs:
                        configs["cfg4"] = leg_cfg4(torch, dev, hbm_peak, 10_000, 10_000, sp)
                    if "wide" in legs:
                        configs["wide"] = leg_wide(torch, dev, hbm_peak, steps, sp)
                        configs["wide512"] = leg_wide(torch, dev, hbm_peak, steps, sp, dim=512)
                    if "scaling_base" in legs:
                        configs["scaling_base"] = leg_scaling_base(torch, dev, hbm_peak, steps)
                else:
                    configs["cfg4"] = leg_cfg4(torch, dev, hbm_peak, 10_000, 10_000, sp, devices=list(range(n_gpus)))
                    if cfg4_spmd:
                        configs["cfg4"].update(cfg4_spmd)
            except Exception as exc:   # noqa: BLE001  (a secondary leg must not lose the headline line)
                configs["error"] = f"{type(exc).__name__}: {exc}"
    if world > 1:
        dist.barrier(group=host_pg)

    if rank == 0:
        ms_per_step = ms / steps
        value = n_total * steps / (ms * 1e-3)
        achieved = shard_rows * BYTES_PER_CHUNK / (kernel_ms * 1e-3) / 1e9
        traffic_key = "cfg2" if (n_gpus == 1 and not args.chunks) else f"shard_{shard_rows}"
        line = {
            "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if n_gpus > 1 else "weak",
            "vs_baseline": None, "dtype": "s8 x s8 -> s32", "data": "synthetic",
            "config": make_config(n_gpus, n_total, custom=bool(args.chunks)),
            "exchange": exchange, "qps": steps / (ms * 1e-3),
            **({"schedule": f"unbatched single-query searches (one scan-kernel launch each), alternating between {len(streams)} caller streams, "
                            "so two searches are in flight on the GPU; roofline.single_stream and e2e.sequential_ms_per_query are the one-stream / one-caller figures"}
               if n_gpus == 1 and len(streams) > 1 else {}),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": load_traffic(traffic_key), "traffic_source": "ncu capture of this kernel at this shard size (profiles/roofline_traffic.json, ncu_traffic_r02c.txt)",
                         "kernel": SCAN_KERNEL, "kernel_ms": kernel_ms, "kernel_launches_timed": kq,
                         "kernel_ms_note": (f"timed region / launches, launches alternating between {len(streams)} caller streams (each stream's searches under programmatic "
                                            "dependent launch): the aggregate rate of the kernel with two queries in flight; one launch alone takes "
                                            "isolated_launch_latency_us" if n_gpus == 1 and len(streams) > 1 else
                                            "local scan + fused top-k per GPU, back-to-back launches on one stream, max over ranks" + ("" if n_gpus == 1 else " (the exchange is in ms_per_step, not here)")),
                         **({"single_stream": {"kernel_ms": single_stream_ms, "achieved": shard_rows * BYTES_PER_CHUNK / (single_stream_ms * 1e-3) / 1e9,
                                               "frac": shard_rows * BYTES_PER_CHUNK / (single_stream_ms * 1e-3) / 1e9 / hbm_peak,
                                               "note": "the same launches back to back on ONE stream (programmatic dependent launch only)"}}
                            if single_stream_ms else {}),
                         "algorithmic_bytes_per_launch": shard_rows * BYTES_PER_CHUNK, "peak_source": peak_src,
                         "frac_of_8TBps": achieved / 8000.0},
            "e2e": {"value": n_total * e2e_steps / e2e_s, "unit": "chunks/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "qps": e2e_steps / e2e_s, "ms_per_query": 1e3 * e2e_s / e2e_steps,
                    "steps": e2e_steps, "qps_4_host_threads": e2e_conc_qps, "api": api,
                    **({"sequential_ms_per_query": 1e3 * e2e_seq_s / e2e_steps} if e2e_seq_s else {})},
            "isolated_launch_latency_us": {"p50": float(np.percentile(lat_us, 50)), "p95": float(np.percentile(lat_us, 95)),
                                           "n": len(lat_us), "note": "one query at a time with a device sync between launches (device-resident query)"},
            "parity": parity, "parity_mismatches": parity.get("mismatches"),
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if configs:
            line["configs"] = configs
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's real stdout (see main)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main() -> None:
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner under torchrun,
    # for one) goes to stderr instead -- file descriptor 1 is pointed at 2 until emit() writes the line
    # watchdog: a hung run (a peer that never arrives, a wedged driver call) dumps every thread's stack to stderr and exits
    # instead of sitting on the GPU box until the caller's limit
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("BENCH_WATCHDOG_S", "900")), exit=True)
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chunks", type=int, default=0, help="override the corpus size (not the headline workload)")
    ap.add_argument("--cpu-budget-s", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[2] / configs[4] / ingest / scaling_base legs")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle checks (profiling runs)")
    ap.add_argument("--streams", type=int, default=2, help="caller streams the N = 1 device-resident searches alternate between")
    ap.add_argument("--legs", default="cfg2,ingest,cfg4,wide,scaling_base", help="which configs legs to run at N = 1 (comma list)")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"], help="multi-GPU top-k exchange")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
