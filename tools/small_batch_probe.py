import os, sys, json
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from rag_foundation_b200 import Engine
import bench
res = {}
for rows in (50_000, 200_000, 1_000_000):
    with Engine(capacity_rows=rows) as e:
        s = e.open_store("fileSearchStores/x"); e.ingest_synthetic(s, 0, 0, 0, rows)
        Q = bench.make_queries(64)
        qd = torch.from_numpy(Q).cuda(); out = torch.zeros((64, 10), dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for nq in (2, 3, 4, 8, 16, 32, 63):
            for _ in range(3): e.search_keys_device(qd.data_ptr(), nq, [s], 10, out.data_ptr(), st)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): e.search_keys_device(qd.data_ptr(), nq, [s], 10, out.data_ptr(), st)
            e1.record(); torch.cuda.synchronize()
            res[(rows, nq)] = e0.elapsed_time(e1) / 20 * 1e3
            ref = out[:nq].cpu().numpy().copy()
            res[(rows, nq, "keys")] = ref
print(os.environ.get("RF_GEMM_MIN_QUERIES"), {k: round(v, 1) for k, v in res.items() if len(k) == 2})
np.save("/root/repo/gpurun_out/small_%s.npy" % os.environ.get("RF_GEMM_MIN_QUERIES", "64"), np.concatenate([res[k].ravel() for k in sorted(k for k in res if len(k) == 3)]))
