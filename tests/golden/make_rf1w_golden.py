#!/usr/bin/env python3
"""Freeze RF-1w (IDF-weighted variant, oracle/SPEC.md) known-answer vectors from the Python oracle
into rf1w_golden.json.  Self-generated, like rf1_golden.json: the reference has no retrieval
arithmetic, so these pin the C oracle and the CUDA path to the frozen spec."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import rf1  # noqa: E402


def main():
    out = {"weight_cases": [[n, df, int(rf1.idf_weights([df], n)[0])] for n, df in
                            [(0, 0), (1, 0), (1, 1), (3, 1), (10, 3), (1000, 0), (1000, 1), (1000, 99), (1000, 124),
                             (1000, 199), (1000, 249), (1000, 333), (1000, 499), (1000, 999), (1000, 1000),
                             (10 ** 6, 12345), (10 ** 8, 7), (2 ** 32 - 2, 0), (2 ** 32 - 2, 2 ** 32 - 2)]]}
    zb = rf1.zipf_bucket_table()
    n, seed = 20_000, 3
    F = rf1.synth_rows(seed, 0, n, zb)
    seg = np.zeros(n, np.uint32)
    seg[5000:9000] = 1
    seg[17] = seg[6000] = rf1.TOMBSTONE
    cases = []
    for scope in ([0], [1], [0, 1]):
        df, live = rf1.bucket_df(F, seg, scope)
        w = rf1.idf_weights(df, live)
        qs = []
        for qi in range(3):
            q = rf1.synth_query(seed, qi, zb)
            qw = rf1.weight_query(q, w)
            ids, sc = rf1.score_topk(F, seg, qw, scope)
            qs.append({"qi": qi, "qw_sparse": [[int(i), int(qw[i])] for i in np.nonzero(qw)[0]],
                       "ids": ids.tolist(), "scores": sc.tolist()})
        cases.append({"scope": scope, "n": live, "df": df.tolist(), "w": w.tolist(), "queries": qs})
    out["synthetic"] = {"seed": seed, "rows": n, "store1": [5000, 9000], "tombstones": [17, 6000], "cases": cases}
    sat = rf1.weight_query(np.array([0, 1, 4, 5, 127] + [0] * 251, np.int8), np.array([31] * 256, np.uint8))
    out["saturation"] = sat[:5].tolist()
    with open(os.path.join(HERE, "rf1w_golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote rf1w_golden.json", os.path.getsize(os.path.join(HERE, "rf1w_golden.json")), "bytes")


if __name__ == "__main__":
    main()
