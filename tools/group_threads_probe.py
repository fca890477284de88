#!/usr/bin/env python3
"""Concurrent callers on an engine group, then a second group in the same process (the sequence bench.py runs at N > 1), with a
watchdog that dumps the Python stacks.  python tools/group_threads_probe.py [n_devices]"""
import faulthandler
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from rag_foundation_b200 import EngineGroup
from rag_foundation_b200.engine import scopes_to_csr

faulthandler.dump_traceback_later(120, exit=True)
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
devs = list(range(G)) if G > 1 else [0, 0]
n = 20_000_000
Qh = bench.make_queries(64)
for cycle in range(2):
    with EngineGroup(devs, capacity_rows=n // len(devs), placement="spread", id_bases=[d * (n // len(devs)) for d in range(len(devs))]) as grp:
        gs = grp.open_store("fileSearchStores/bench")
        grp.ingest_synthetic(gs, 0, seed=0, start_counter=0, n_rows=n)
        for i in range(10):
            grp.search(Qh[i:i + 1], [[gs]], k=10)
        print("cycle", cycle, "seq ok", flush=True)

        def cl(tid, m):
            for i in range(m):
                grp.search(Qh[(tid * 31 + i) % 64:(tid * 31 + i) % 64 + 1], [[gs]], k=10)
        for nthr in (2, 4):
            ts = [threading.Thread(target=cl, args=(t, 10 if nthr == 2 else 25)) for t in range(nthr)]
            t0 = time.perf_counter(); [t.start() for t in ts]; [t.join() for t in ts]
            print(nthr, "threads ok", flush=True)
    n_stores, per_store, nq = 2000, 10_000, 1024
    rng = np.random.default_rng(5)
    Q = bench.make_queries(nq, seed=4)
    csr = scopes_to_csr([[int(rng.integers(0, n_stores))] for _ in range(nq)])
    with EngineGroup(devs, capacity_rows=(n_stores + len(devs) - 1) // len(devs) * per_store, placement="store") as eng:
        for i in range(n_stores):
            eng.open_store(f"fileSearchStores/mt{i}")
        eng.ingest_synthetic(0, per_store, seed=4, start_counter=0, n_rows=n_stores * per_store)
        for it in range(3):
            ids, sc, cs, cnt = eng.search(Q, csr, k=10)
        print("cycle", cycle, "store-placed group ok", int(cnt.sum()), flush=True)
print("done", flush=True)
