#!/usr/bin/env python3
"""Freeze known-answer vectors of the WIDER RF-1 rows (D = 512, 1024; oracle/SPEC.md "Wider rows") from the
Python oracle into rf1_wide_golden.json.  Self-generated like rf1_golden.json (the reference has no
retrieval arithmetic); the document texts are taken from that fixture, so nothing outside the repo is read."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import rf1  # noqa: E402


def sparse(row):
    nz = np.nonzero(row)[0]
    return [[int(i), int(row[i])] for i in nz]


def main():
    base = json.load(open(os.path.join(HERE, "rf1_golden.json")))
    out = {}
    for dim in (512, 1024):
        g = {}
        for which in ("sample_report", "long_doc"):
            text = base[which]["text"].encode("utf-8")
            F, ff, spans, ntok = rf1.featurize_doc(text, dim)
            g[which] = {"n_tokens": ntok, "rows_sparse": [sparse(r) for r in F], "ff": ff.tolist(), "spans": spans.tolist()}
        q = rf1.query_vector(base["sample_report"]["query"].encode(), dim)
        Fd, ffd, _, _ = rf1.featurize_doc(base["sample_report"]["text"].encode("utf-8"), dim)
        g["demo_query"] = {"q_sparse": sparse(q), "scores": rf1.scores(Fd, q).tolist()}
        zb = rf1.zipf_bucket_table(dim=dim)
        g["zipf_bucket_sha256"] = hashlib.sha256(zb.astype("<u2").tobytes()).hexdigest()
        n = 3000
        F = rf1.synth_rows(3, 100, n, zb, dim=dim)
        g["synth_rows_sha256"] = hashlib.sha256(F.tobytes()).hexdigest()
        g["synth_row_0_sparse"] = sparse(F[0])
        seg = np.zeros(n, np.uint32)
        seg[1::5] = 1
        seg[7::11] = rf1.TOMBSTONE
        cases = []
        for qi in range(4):
            qv = rf1.synth_query(3, qi, zb, dim=dim)
            ids, sc = rf1.score_topk(F, seg, qv, [0], k=10, id_base=50)
            cases.append({"qi": qi, "scope": [0], "ids": ids.tolist(), "scores": sc.tolist()})
            ids, sc = rf1.score_topk(F, seg, qv, [0, 1], k=10, id_base=50)
            cases.append({"qi": qi, "scope": [0, 1], "ids": ids.tolist(), "scores": sc.tolist()})
        g["synth_top10"] = {"seed": 3, "start": 100, "n_rows": n, "id_base": 50, "cases": cases}
        df, nn = rf1.bucket_df(F, seg, [0, 1])
        w = rf1.idf_weights(df, nn)
        g["rf1w"] = {"n": nn, "df_sha256": hashlib.sha256(df.astype("<u8").tobytes()).hexdigest(), "weights": w.tolist()}
        out[str(dim)] = g
    with open(os.path.join(HERE, "rf1_wide_golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote rf1_wide_golden.json", os.path.getsize(os.path.join(HERE, "rf1_wide_golden.json")), "bytes")


if __name__ == "__main__":
    main()
