#!/usr/bin/env python3
"""BASELINE.json configs[4] across the GPUs of one box, sharded by STORE: 10 000
stores x 10 000 chunks, store g on rank g % G, 1024 store-scoped queries per batch replicated to
every rank.  Each rank scores the queries whose store it owns in one launch; one all-gather of the
packed keys + merge kernel gives every rank the answers.  Strong scaling: the corpus and the batch
are fixed, per-rank work shrinks with G.

  torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29533 tests/perf/store_sharded_bench.py

One JSON line from rank 0 (CUDA events, max over ranks, >= 3 warm-ups); results are checked against
the C oracle on a sample of queries (regenerating the scoped store from its counters)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle as co, rf1  # noqa: E402  (checker only)
from rag_foundation_b200 import Engine  # noqa: E402
from rag_foundation_b200.sharded import StoreShardedSearcher, unpack_keys_torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stores", type=int, default=10_000)
    ap.add_argument("--per-store", type=int, default=10_000)
    ap.add_argument("--queries", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    per = args.per_store
    mine = [g for g in range(args.stores) if g % world == rank]
    eng = Engine(capacity_rows=max(1, len(mine)) * per, device=local_rank, id_base=StoreShardedSearcher.id_base_for(rank, world))
    s = StoreShardedSearcher.for_engine(eng)
    for g in range(args.stores):
        s.open_store(f"fileSearchStores/mt{g}")
    # the owned stores hold the rows of the single-index corpus: store g = counters [g * per, (g + 1) * per)
    for g in mine:
        eng.ingest_synthetic(s.local_seg[g], 0, seed=5, start_counter=g * per, n_rows=per)
    zb = rf1.zipf_bucket_table()
    rng = np.random.default_rng(5)
    scopes = [[int(rng.integers(0, args.stores))] for _ in range(args.queries)]
    Q = np.stack([co.synth_query(5, i, zb) for i in range(args.queries)])
    qd = torch.from_numpy(Q).to(dev)
    ids, sc, valid = s.search(qd, scopes, 10)
    torch.cuda.synchronize()
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    bad = 0
    n_check = min(args.queries, 24)
    for i in range(n_check):
        g = scopes[i][0]
        F = co.synth_rows(5, g * per, per, zb)
        base = StoreShardedSearcher.id_base_for(g % world, world) + (g // world) * per
        w_ids, w_sc, _ = co.score_topk(F, np.zeros(per, np.uint32), Q[i], [0], id_base=base)
        bad += int(ids[i].tolist() != w_ids.tolist() or sc[i].tolist() != w_sc.tolist())
    prepared = s.prepare(scopes)
    assert torch.equal(s.search_keys(qd, prepared, 10), s.search_keys(qd, scopes, 10))
    for _ in range(3):
        s.search_keys(qd, prepared, 10)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        s.search_keys(qd, prepared, 10)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.reps, float(bad)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    if rank == 0:
        alg = args.queries * per * 260
        print(json.dumps({"config": "configs[4] sharded by store: %d stores x %d chunks, %d store-scoped queries per batch" % (args.stores, per, args.queries),
                          "n_gpus": world, "scaling": "strong", "ms_per_batch": ms, "qps": args.queries / (ms * 1e-3),
                          "chunks_per_s": args.queries * per / (ms * 1e-3), "algorithmic_GBps_aggregate": alg / (ms * 1e-3) / 1e9,
                          "parity_mismatches": int(t[1]), "parity_checked": n_check,
                          "timing": "CUDA events on the launching stream, max over ranks; includes the host-side plan build for 1024 scopes, "
                                    "its H2D, the scan launch, the NCCL all-gather and the merge kernel"}), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
