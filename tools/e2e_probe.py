import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["RF_PROFILE"] = "1"
from rag_foundation_b200 import Engine
import bench
N = 1_000_000
Q = bench.make_queries(64)
with Engine(capacity_rows=N) as e:
    s = e.open_store("fileSearchStores/x"); e.ingest_synthetic(s, 0, 0, 0, N)
    for i in range(50): e.search(Q[i % 64][None], [[s]], k=10)
    t = time.perf_counter()
    for i in range(2000): e.search(Q[i % 64][None], [[s]], k=10)
    print("python API us/query", (time.perf_counter() - t) / 2000 * 1e6)
    t = time.perf_counter()
    for i in range(2000): pass
    q = np.ascontiguousarray(Q[0][None]); segs = np.array([s], np.uint32); off = np.array([0, 1], np.uint32)
    ids = np.empty((1, 10), np.uint64); sc = np.empty((1, 10), np.int32); cs = np.empty((1, 10), np.float32); cnt = np.empty(1, np.uint32)
    args = (e._h, q.ctypes.data, 1, segs.ctypes.data, off.ctypes.data, 10, ids.ctypes.data, sc.ctypes.data, cs.ctypes.data, cnt.ctypes.data)
    f = e._L.rf_search
    t = time.perf_counter()
    for i in range(2000): f(*args)
    print("raw ctypes us/query", (time.perf_counter() - t) / 2000 * 1e6)
