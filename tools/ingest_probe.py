#!/usr/bin/env python3
"""One 22.8 MB document through rf_ingest_text three times with the text already in HBM -- the subject of the
ncu launch list of the featurise kernels (profiles/ncu_launches_ingest_r02.csv)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from rag_foundation_b200 import Engine  # noqa: E402

data = bench.make_text(22_800_000)
dd = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
torch.cuda.synchronize()
with Engine(capacity_rows=400_000) as e:
    s = e.open_store("fileSearchStores/i")
    for r in range(3):
        print(e.ingest_text_ptr(s, r, dd.data_ptr(), len(data)))
