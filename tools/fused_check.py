#!/usr/bin/env python3
"""torchrun check of the fused top-k exchange (NVLink peer stores + flags) against the NCCL
all-gather + merge path: identical keys over many back-to-back calls, then timing of both.
  torchrun --nproc-per-node N tools/fused_check.py [rows_total]"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rag_foundation_b200 import Engine  # noqa: E402
from rag_foundation_b200.sharded import FusedShardedSearcher, ShardedSearcher, shard_range  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
lo, hi = shard_range(n_total, rank, world)
eng = Engine(capacity_rows=hi - lo, device=local, id_base=lo)
seg = eng.open_store("fileSearchStores/x")
eng.ingest_synthetic(seg, 0, seed=0, start_counter=lo, n_rows=hi - lo)
nccl = ShardedSearcher.for_engine(eng)
fused = FusedShardedSearcher(eng, nq_cap=8, k=10)
Q = torch.from_numpy(bench.make_queries(64)).to(dev)
bad = 0
for it in range(200):
    nq = 1 + it % 8          # 4 and more: FusedShardedSearcher hands the batch to the collective path
    q = Q[(it * 3) % 60:(it * 3) % 60 + nq].contiguous()
    a = nccl.search_keys(q, [seg], 10)
    b = fused.search_keys(q, [seg], 10)
    bad += int(not torch.equal(a, b))
torch.cuda.synchronize()
t = torch.tensor([bad, int(fused.timed_out())], device=dev)
dist.all_reduce(t)
res = {"world": world, "rows_total": n_total, "mismatching_calls": int(t[0]), "timeouts": int(t[1])}


def timeit(fn, n=300):
    for i in range(20):
        fn(i)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    x = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(x, op=dist.ReduceOp.MAX)
    return float(x)


if res["mismatching_calls"] == 0 and res["timeouts"] == 0:
    res["nccl_ms_per_query"] = timeit(lambda i: nccl.search_keys(Q[i % 64:i % 64 + 1], [seg], 10))
    res["fused_ms_per_query"] = timeit(lambda i: fused.search_keys(Q[i % 64:i % 64 + 1], [seg], 10))
    res["local_scan_only_ms"] = timeit(lambda i: nccl.local_search(Q[i % 64:i % 64 + 1], [seg], 10))
    res["batch8_ms"] = timeit(lambda i: fused.search_keys(Q[i % 56:i % 56 + 8], [seg], 10), n=50)
    fused.BATCH_THRESHOLD = 99
    res["batch8_fused_scan_ms"] = timeit(lambda i: fused.search_keys(Q[i % 56:i % 56 + 8], [seg], 10), n=50)
if rank == 0:
    print(json.dumps(res))
eng.close()
dist.destroy_process_group()
