#!/usr/bin/env python3
"""CUDA-graph replay of back-to-back single-query searches (configs[1]): capture 64 launches of
rf_search_keys_device (programmatic dependent launch edges included) and replay them.
Checks the replayed keys against the eager ones and prints us/query for both."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rag_foundation_b200 import Engine
import bench

N, NCAP, k = 1_000_000, 64, 10
Q = bench.make_queries(NCAP)
dev = torch.device("cuda", 0)
with Engine(capacity_rows=N) as e:
    s = e.open_store("fileSearchStores/x"); e.ingest_synthetic(s, 0, 0, 0, N)
    Qd = torch.from_numpy(Q).to(dev)
    eager = torch.zeros((NCAP, k), dtype=torch.int64, device=dev)
    out = torch.zeros((NCAP, k), dtype=torch.int64, device=dev)
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        for i in range(NCAP):   # warm the per-(scope, stream) plan outside the capture
            e.search_keys_device(Qd[i:i + 1].data_ptr(), 1, [s], k, eager[i].data_ptr(), side.cuda_stream)
    side.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        cs = torch.cuda.current_stream(dev).cuda_stream
        assert cs == side.cuda_stream
        for i in range(NCAP):
            e.search_keys_device(Qd[i:i + 1].data_ptr(), 1, [s], k, out[i].data_ptr(), cs)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    same = bool(torch.equal(out, eager))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 40
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    graph_us = e0.elapsed_time(e1) / (reps * NCAP) * 1e3
    with torch.cuda.stream(side):
        e0.record(side)
        for r in range(reps):
            for i in range(NCAP):
                e.search_keys_device(Qd[i:i + 1].data_ptr(), 1, [s], k, out[i].data_ptr(), side.cuda_stream)
        e1.record(side)
    torch.cuda.synchronize()
    eager_us = e0.elapsed_time(e1) / (reps * NCAP) * 1e3
    print(json.dumps({"graph_replay_us_per_query": graph_us, "eager_us_per_query": eager_us, "keys_equal": same,
                      "captured_launches": NCAP, "after_graph_eager_still_equal": bool(torch.equal(out, eager))}))
