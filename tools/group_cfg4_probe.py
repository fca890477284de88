#!/usr/bin/env python3
"""configs[4] through an engine group on the box's GPUs, one process (what bench.py's cfg4 leg does at N > 1), with a
watchdog that dumps the Python stacks: python tools/group_cfg4_probe.py [n_devices] [n_stores]"""
import faulthandler
import sys
import time

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np

import bench
from rag_foundation_b200 import EngineGroup

faulthandler.dump_traceback_later(90, exit=True)
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_stores = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
per_store, nq = 10_000, 1024
rng = np.random.default_rng(5)
Q = bench.make_queries(nq, seed=4)
scopes = [[int(rng.integers(0, n_stores))] for _ in range(nq)]
from rag_foundation_b200.engine import scopes_to_csr
csr = scopes_to_csr(scopes)
t0 = time.time()
with EngineGroup(list(range(G)), capacity_rows=(n_stores + G - 1) // G * per_store, placement="store") as eng:
    for i in range(n_stores):
        eng.open_store(f"fileSearchStores/mt{i}")
    print("stores open", time.time() - t0, flush=True)
    eng.ingest_synthetic(0, per_store, seed=4, start_counter=0, n_rows=n_stores * per_store)
    print("ingested", time.time() - t0, flush=True)
    for it in range(5):
        ids, sc, cs, cnt = eng.search(Q, csr, k=10)
        print("search", it, time.time() - t0, int(cnt.sum()), flush=True)
print("done", flush=True)
