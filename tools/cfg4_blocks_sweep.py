#!/usr/bin/env python3
"""Blocks-per-query sweep for small store-scoped batches (what ONE rank of an 8-GPU store-sharded configs[4] runs:
128 queries x 10 k chunks each): RF_SCAN_BLOCKS = 0 (engine's choice) and 1..16, CUDA events, one JSON line each."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import json, os, sys
sys.path.insert(0, %r)
import numpy as np, torch
import bench
from rag_foundation_b200 import Engine
from rag_foundation_b200.engine import scopes_to_csr
nq, n_st, per = int(sys.argv[1]), 1250, 10000
rng = np.random.default_rng(5)
Q = bench.make_queries(nq, seed=4)
scopes = [[int(x)] for x in rng.choice(n_st, size=nq, replace=False)]
csr = scopes_to_csr(scopes)
with Engine(capacity_rows=n_st * per) as e:
    for i in range(n_st):
        e.open_store("s%%d" %% i)
    e.ingest_synthetic(0, per, seed=4, start_counter=0, n_rows=n_st * per)
    qd = torch.from_numpy(Q).cuda()
    out = torch.zeros((nq, 10), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream()
    torch.cuda.synchronize()
    e.set_stream_overlap(st.cuda_stream, True)
    ms = bench.events_ms(torch, st, lambda: e.search_keys_device_scoped(qd.data_ptr(), nq, csr, 10, out.data_ptr(), st.cuda_stream), reps=50, warm=5)
print(json.dumps({"nq": nq, "blocks_per_query": int(os.environ.get("RF_SCAN_BLOCKS", "0")), "us": ms * 1e3, "GBps": nq * per * 260 / (ms * 1e-3) / 1e9}))
""" % ROOT

for nq in (128, 256, 512):
    for x in (0, 1, 2, 3, 4, 5, 6, 8, 10, 12, 16):
        env = dict(os.environ, RF_SCAN_BLOCKS=str(x))
        if x == 0:
            env.pop("RF_SCAN_BLOCKS")
        o = subprocess.run([sys.executable, "-c", CHILD, str(nq)], capture_output=True, text=True, env=env)
        print(o.stdout.strip().splitlines()[-1] if o.returncode == 0 else json.dumps({"error": o.stderr[-300:]}), flush=True)
