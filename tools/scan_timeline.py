#!/usr/bin/env python3
"""Per-block timeline of the scan kernel from its %globaltimer stamps (RF_SCAN_DEBUG=1)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RF_SCAN_DEBUG"] = "1"
from rag_foundation_b200 import Engine, _capi  # noqa: E402

N = int(os.environ.get("SWEEP_ROWS", "1000000"))
for v, blocks in [(int(x.split(":")[0]), int(x.split(":")[1])) for x in os.environ.get("CONFIGS", "0:148,0:296,1:148").split(",")]:
    os.environ["RF_SCAN_VARIANT"] = str(v)
    os.environ["RF_SCAN_BLOCKS"] = str(blocks)
    with Engine(capacity_rows=N) as e:
        s = e.open_store("fileSearchStores/t")
        e.ingest_synthetic(s, 0, seed=0, start_counter=0, n_rows=N)
        q = torch.randint(0, 3, (64, 256), dtype=torch.int8, device="cuda")
        out = torch.zeros((64, 10), dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for i in range(20):
            e.search_keys_device(q[i:i + 1].data_ptr(), 1, [s], 10, out[i].data_ptr(), st)
        torch.cuda.synchronize()
        buf = np.zeros(blocks * 8, np.uint64)
        _capi.check(_capi.lib().rf_debug_timestamps(e.handle, buf.ctypes.data, buf.size, 1))
        rows = []
        for rep in range(5):
            e.search_keys_device(q[rep:rep + 1].data_ptr(), 1, [s], 10, out[rep].data_ptr(), st)
            _capi.check(_capi.lib().rf_debug_timestamps(e.handle, buf.ctypes.data, buf.size, 1))
            t = buf.reshape(blocks, 8).astype(np.int64)
            t0 = t[:, 0].min()
            rel = (t - t0) / 1e3
            last = t[:, 5].max()
            lb = int(t[:, 5].argmax())
            extra = (t[lb, 4] - t0) / 1e3, (t[lb, 6] - t0) / 1e3, float(np.mean(rel[:, 7]))
            rows.append([rel[:, 0].max(), rel[:, 1].mean(), rel[:, 1].max(), rel[:, 2].mean(), rel[:, 2].max(),
                         np.median(rel[:, 3]), rel[:, 3].min(), rel[:, 3].max(), rel[:, 4].max(), (last - t0) / 1e3])
        r = np.median(np.array(rows), axis=0)
        print(f"variant {v} blocks {blocks} rows {N} (us from first block entry): last-entry {r[0]:.1f} | plan mean {r[1]:.1f} max {r[2]:.1f} | "
              f"first-tile mean {r[3]:.1f} max {r[4]:.1f} | loop-done med {r[5]:.1f} min {r[6]:.1f} max {r[7]:.1f} | "
              f"partial published max {r[8]:.1f} | answer {r[9]:.1f} | last block: published {extra[0]:.1f} level-1 done {extra[1]:.1f} | producer issued first copy at {extra[2]:.2f}", flush=True)
