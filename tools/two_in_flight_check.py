"""Store-sharded fused exchange with two batches in flight per rank (alternating caller streams) against the same batches
issued one after the other: where do the answers differ?  torchrun --nproc-per-node 2 tools/two_in_flight_check.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rag_foundation_b200 import Engine  # noqa: E402
from rag_foundation_b200.sharded import FusedStoreShardedSearcher  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    n_stores, per_store, nq, K = 2000, 10_000, 1024, 10
    rng = np.random.default_rng(5)
    Q = bench.make_queries(nq, seed=bench.SEED + 4)
    scopes = [[int(rng.integers(0, n_stores))] for _ in range(nq)]
    owned = (n_stores + world - 1) // world
    eng = Engine(capacity_rows=owned * per_store, device=lr, id_base=rank * ((1 << 32) // world))
    srch = FusedStoreShardedSearcher(eng, nq_cap=nq, k=K)
    for g in range(n_stores):
        srch.open_store(f"fileSearchStores/mt{g}")
    for g in range(rank, n_stores, world):
        eng.ingest_synthetic(srch.local_seg[g], 0, seed=bench.SEED + 4, start_counter=g * per_store, n_rows=per_store)
    qd = torch.from_numpy(Q).to(dev)
    local = srch.prepare_fused(scopes)
    masks = local[4]
    stream = torch.cuda.current_stream(dev)
    s2 = torch.cuda.Stream(dev)
    overlap = os.environ.get("OVERLAP", "1") == "1"
    eng.set_stream_overlap(stream.cuda_stream, overlap)
    eng.set_stream_overlap(s2.cuda_stream, overlap)
    ref = torch.zeros((nq, K), dtype=torch.int64, device=dev)
    srch.search_keys(qd, local, K, out=ref)
    torch.cuda.synchronize(dev)
    dist.barrier()
    keys = ref.cpu().numpy().view(np.uint64)
    outs = [[torch.zeros_like(ref) for _ in range(12)] for _ in range(2)]
    streams = [stream, s2]
    for i in range(24):
        with torch.cuda.stream(streams[i & 1]):
            srch.search_keys(qd, local, K, out=outs[i & 1][i // 2])
    torch.cuda.synchronize(dev)
    dist.barrier()
    report = {"rank": rank, "timed_out": srch.timed_out(), "bad": []}
    for par in range(2):
        for j in range(12):
            got = outs[par][j].cpu().numpy().view(np.uint64)
            bad_q = np.nonzero((got != keys).any(axis=1))[0]
            if len(bad_q):
                mine = [int(q) for q in bad_q[:6]]
                report["bad"].append({"stream": par, "call": 2 * j + par, "n_bad": int(len(bad_q)),
                                      "first": mine, "owner_mask": [int(masks[q]) for q in mine],
                                      "zero_rows": int((got[bad_q] == 0).all(axis=1).sum())})
    print(json.dumps(report), flush=True)
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
