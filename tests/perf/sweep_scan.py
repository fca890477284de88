#!/usr/bin/env python3
"""Sweep the scan-kernel variants / grid sizes on one GPU (development tool; run under gpurun).
Each configuration is checked against the C oracle on 4 queries before it is timed."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle as co, rf1  # noqa: E402  (checker)
from rag_foundation_b200 import Engine  # noqa: E402

N = int(os.environ.get("SWEEP_ROWS", "1000000"))
zb = rf1.zipf_bucket_table()
F = co.synth_rows(0, 0, N, zb)
seg0 = np.zeros(N, np.uint32)
Q = np.stack([co.synth_query(0, i, zb) for i in range(64)])
want = [co.score_topk_keys(F, seg0, Q[i], [0]).tolist() for i in range(4)]
names = {0: "ldg8x2", 1: "tma8x24", 2: "tma12x24", 3: "tma8x16", 4: "tma6x12", 5: "tma12x12", 6: "tma4x12", 7: "tma4x8"}
configs = []
for v in [int(x) for x in os.environ.get("SWEEP_VARIANTS", "0,1,2,3,4").split(",")]:
    for mult in [float(x) for x in os.environ.get("SWEEP_MULTS", "1,2,3,4").split(",")]:
        configs.append((v, int(148 * mult)))
results = []
for v, blocks in configs:
    os.environ["RF_SCAN_VARIANT"] = str(v)
    os.environ["RF_SCAN_BLOCKS"] = str(blocks)
    try:
        with Engine(capacity_rows=N) as e:
            s = e.open_store("fileSearchStores/sweep")
            e.ingest_synthetic(s, 0, seed=0, start_counter=0, n_rows=N)
            qd = torch.from_numpy(Q).cuda()
            out = torch.zeros((64, 10), dtype=torch.int64, device="cuda")
            st = torch.cuda.current_stream().cuda_stream
            for i in range(64):
                e.search_keys_device(qd[i:i + 1].data_ptr(), 1, [s], 10, out[i].data_ptr(), st)
            torch.cuda.synchronize()
            got = out.cpu().numpy().view(np.uint64)
            ok = all(got[i].tolist() == want[i] for i in range(4))
            best = 1e9
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(300):
                    e.search_keys_device(qd[i % 64:i % 64 + 1].data_ptr(), 1, [s], 10, out[i % 64].data_ptr(), st)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / 300)
            r = {"variant": names[v], "blocks": blocks, "us": best * 1e3, "GBps": N * 260 / (best * 1e-3) / 1e9, "parity": ok}
    except Exception as ex:  # noqa: BLE001
        r = {"variant": names[v], "blocks": blocks, "error": str(ex)[:200]}
    results.append(r)
    print(json.dumps(r), flush=True)
json.dump(results, open(os.path.join(ROOT, "gpurun_out", "sweep_scan.json"), "w"), indent=1)
