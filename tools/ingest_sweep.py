#!/usr/bin/env python3
"""Sweep of the ingest staging parameters (copy-chunk size x helper threads) for one 22.8 MB document from
pageable host memory, plus the host's own memcpy rate -- run under gpurun; one JSON line per setting."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r"""
import json, os, sys, time
sys.path.insert(0, %r)
import numpy as np
import bench
from rag_foundation_b200 import Engine
data = bench.make_text(22_800_000)
with Engine(capacity_rows=2_000_000) as e:
    s = e.open_store("fileSearchStores/i")
    for r in range(3):
        e.ingest_text(s, r, data, want_spans=False)
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for r in range(5):
            e.ingest_text(s, 10 + rep * 5 + r, data, want_spans=False)
        best = min(best, (time.perf_counter() - t0) / 5)
print(json.dumps({"chunk_kb": int(os.environ.get("RF_STAGE_CHUNK_KB", "0")), "threads": int(os.environ.get("RF_STAGE_THREADS", "-1")),
                  "ms": best * 1e3, "GBps": len(data) / best / 1e9}))
""" % ROOT


def main():
    import numpy as np
    a = np.frombuffer(os.urandom(1 << 20) * 23, np.uint8).copy()
    b = np.empty_like(a)
    np.copyto(b, a)
    t0 = time.perf_counter()
    for _ in range(10):
        np.copyto(b, a)
    print(json.dumps({"host_memcpy_1_thread_GBps": a.nbytes * 10 / (time.perf_counter() - t0) / 1e9, "cpus": len(os.sched_getaffinity(0))}), flush=True)
    for kb in (1024, 2048):
        for th in (0, 1, 2):
            env = dict(os.environ, RF_STAGE_CHUNK_KB=str(kb), RF_STAGE_THREADS=str(th))
            out = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=env)
            print(out.stdout.strip().splitlines()[-1] if out.returncode == 0 else json.dumps({"error": out.stderr[-300:]}), flush=True)


if __name__ == "__main__":
    main()
