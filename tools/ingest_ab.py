"""Device time of the featurise kernels (CUDA events inside the engine) and the call's wall time for one
device-resident document; RF_TOKENIZE_SPAN_KB sets the bytes per tokeniser CTA.
Usage: python tools/ingest_ab.py [bytes]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from rag_foundation_b200 import Engine  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 22_800_000
    data = bench.make_text(n)
    reps = 10
    with Engine(capacity_rows=(len(data) // 400 + 64) * (reps + 4) * 2) as e:
        s = e.open_store("fileSearchStores/ab")
        dd = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
        torch.cuda.synchronize()
        first, nc = e.ingest_text_ptr(s, 1, dd.data_ptr(), len(data))
        F0 = e.read_rows(first - e.id_base, nc)[0]
        k0 = e.stats()["ingest_kernel_ns"]
        t0 = time.perf_counter()
        for r in range(reps):
            f, c = e.ingest_text_ptr(s, 2 + r, dd.data_ptr(), len(data))
        wall = (time.perf_counter() - t0) / reps
        kern = (e.stats()["ingest_kernel_ns"] - k0) / reps
        same = bool(c == nc and (e.read_rows(f - e.id_base, c)[0] == F0).all())
        print(json.dumps({"bytes": len(data), "chunks": nc,
                          "kernel_us": kern / 1e3, "wall_us": wall * 1e6, "rows_repeatable": same,
                          "rows_sum": int(F0.astype(np.int64).sum())}))


if __name__ == "__main__":
    main()
