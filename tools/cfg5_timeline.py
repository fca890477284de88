import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["RF_SCAN_DEBUG"] = "1"
from rag_foundation_b200 import Engine, _capi
import bench
NS, PER, NQ = 2000, 10_000, int(os.environ.get("NQ", "128"))
X = int(os.environ.get("RF_SCAN_BLOCKS", "7"))
os.environ["RF_SCAN_BLOCKS"] = str(X)
Q = np.concatenate([bench.make_queries(64)] * (NQ // 64 + 1))[:NQ]
rng = np.random.default_rng(1)
with Engine(capacity_rows=NS * PER) as e:
    first = e.open_store("fileSearchStores/m0")
    for i in range(1, NS): e.open_store(f"fileSearchStores/m{i}")
    e.ingest_synthetic(first, PER, 5, 0, NS * PER)
    scopes = [[int(first + rng.integers(0, NS))] for _ in range(NQ)]
    for _ in range(3): e.search(Q, scopes, k=10)
    nb = X * NQ
    buf = np.zeros(nb * 8, np.uint64)
    _capi.check(_capi.lib().rf_debug_timestamps(e.handle, buf.ctypes.data, buf.size, 1))
    e.search(Q, scopes, k=10)
    _capi.check(_capi.lib().rf_debug_timestamps(e.handle, buf.ctypes.data, buf.size, 1))
    t = buf.reshape(nb, 8).astype(np.int64)
    t0 = t[:, 0].min()
    r = (t - t0) / 1e3
    life = r[:, 4] - r[:, 0]
    print(f"X={X} nq={NQ} blocks={nb}: kernel span {r[:, 5].max():.1f} us | block entry: first {r[:,0].min():.1f} median {np.median(r[:,0]):.1f} last {r[:,0].max():.1f}")
    print(f"  per block (us): entry->plan {np.median(r[:,1]-r[:,0]):.2f} | entry->first tile {np.median(r[:,2]-r[:,0]):.2f} | entry->scan done {np.median(r[:,3]-r[:,0]):.2f} | entry->published {np.median(life):.2f}")
