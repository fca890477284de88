import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["RF_SCAN_DEBUG"] = "1"
from rag_foundation_b200 import Engine, _capi
import bench
N, NQ = 1_000_000, int(os.environ.get("NQ", "1024"))
Q = bench.make_queries(64)
Q = np.concatenate([Q] * (NQ // 64 + 1))[:NQ]
with Engine(capacity_rows=N) as e:
    s = e.open_store("fileSearchStores/x"); e.ingest_synthetic(s, 0, 0, 0, N)
    qd = torch.from_numpy(np.ascontiguousarray(Q)).cuda(); out = torch.zeros((NQ, 10), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): e.search_keys_device(qd.data_ptr(), NQ, [s], 10, out.data_ptr(), st)
    torch.cuda.synchronize()
    nb = 148
    buf = np.zeros(4096, np.uint64)
    _capi.check(_capi.lib().rf_debug_timestamps(e.handle, buf.ctypes.data, buf.size, 1))
    e.search_keys_device(qd.data_ptr(), NQ, [s], 10, out.data_ptr(), st)
    _capi.check(_capi.lib().rf_debug_timestamps(e.handle, buf.ctypes.data, buf.size, 1))
    tl = buf[2048:2048 + 16 * 8].reshape(16, 8).astype(np.int64)
    t = buf[:nb * 8].reshape(nb, 8).astype(np.float64)
    t = t[t[:, 3] > 0]
    print("blocks", len(t), "tiles/block", t[:, 3].mean())
    print("MMA thread: total cyc %.0f | wait TMA full %.0f (%.0f%%) | wait tmem_empty %.0f (%.0f%%) | per tile %.0f cyc" % (t[:, 0].mean(), t[:, 1].mean(), 100 * t[:, 1].mean() / t[:, 0].mean(), t[:, 2].mean(), 100 * t[:, 2].mean() / t[:, 0].mean(), (t[:, 0] / t[:, 3]).mean()))
    print("epilogue warp 2: total cyc %.0f | wait tmem_full %.0f (%.0f%%) | candidate path %.0f (%.0f%%), entered %.1f times/tile" % (t[:, 4].mean(), t[:, 5].mean(), 100 * t[:, 5].mean() / t[:, 4].mean(), t[:, 6].mean(), 100 * t[:, 6].mean() / t[:, 4].mean(), (t[:, 7] / t[:, 3]).mean()))
    if tl[:, 0].any():
        base = tl[0, 0]
        print("timeline of the first cluster's leader CTA, tiles 16..23 (cycles from the first stamp): MMAs issued | warp 2: ready, handed back, done | warp 17: same")
        for i in range(16):
            r = [int(x - base) if x else -1 for x in tl[i, :7]]
            print("  tile %2d group %d: issued %6d | w2 %6d %6d %6d | w17 %6d %6d %6d" % (16 + i // 2, i % 2, *r))
