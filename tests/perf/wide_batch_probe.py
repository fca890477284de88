"""1024 queries x 1 M chunks at D = 512 through the streamed-K tensor-core kernel: ms per batch (CUDA events), sampled
parity against the C oracle.  Usage: python tests/perf/wide_batch_probe.py [dim]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import bench  # noqa: E402
from rag_foundation_b200 import Engine  # noqa: E402


def main():
    dim = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    n, nq, k = 1_000_000, 1024, 10
    from oracle import c_oracle as co, rf1
    zb = rf1.zipf_bucket_table(dim=dim)
    Q = np.stack([co.synth_query(bench.SEED + 6, i, zb, dim=dim) for i in range(nq)])
    with Engine(capacity_rows=n, dim=dim) as e:
        s = e.open_store("fileSearchStores/w")
        e.ingest_synthetic(s, 0, seed=bench.SEED, start_counter=0, n_rows=n)
        qd = torch.from_numpy(Q).cuda()
        out = torch.zeros((nq, k), dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream()
        l0 = e.stats()["kernel_launches"]
        e.search_keys_device(qd.data_ptr(), nq, [s], k, out.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        launches = e.stats()["kernel_launches"] - l0
        ms = bench.events_ms(torch, st, lambda: e.search_keys_device(qd.data_ptr(), nq, [s], k, out.data_ptr(), st.cuda_stream), reps=20)
        keys = out.cpu().numpy().view(np.uint64)
        F = co.synth_rows(bench.SEED, 0, n, zb, dim=dim)
        seg = np.full(n, s, np.uint32)
        bad = sum(int(keys[i].tolist() != co.score_topk_keys(F, seg, Q[i], [s], k=k).tolist()) for i in range(0, nq, 64))
        ops = 2.0 * nq * n * dim
        print(json.dumps({"dim": dim, "ms_per_batch": ms, "launches": launches, "TOPs": ops / (ms * 1e-3) / 1e12,
                          "hbm_GBps_one_pass": n * (dim + 4) / (ms * 1e-3) / 1e9, "mismatches": bad, "checked": len(range(0, nq, 64))}))


if __name__ == "__main__":
    main()
