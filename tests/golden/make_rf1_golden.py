#!/usr/bin/env python3
"""Freeze RF-1 known-answer vectors from the Python oracle (oracle/rf1.py) into rf1_golden.json.

These are SELF-generated (the reference has no retrieval arithmetic to generate them from -- see
oracle/SPEC.md), so they pin the C oracle and the CUDA path to the frozen spec, not to upstream.
The sample document text is embedded in the fixture so nothing reads /root/reference at test time.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import rf1  # noqa: E402

DEMO_QUERY = "What does this demo prove about the RAG engineering flow?"


def sparse(row):
    nz = np.nonzero(row)[0]
    return [[int(i), int(row[i])] for i in nz]


def main():
    doc = open("/root/reference/docs/demo/sample-report.md", "rb").read()
    zb = rf1.zipf_bucket_table()
    out = {"fnv1a32": {t: rf1.fnv1a32(t.encode()) for t in ["", "a", "foobar", "rag", "demo", "0", "49999"]}}

    F, ff, spans, ntok = rf1.featurize_doc(doc)
    q = rf1.query_vector(DEMO_QUERY.encode())
    out["sample_report"] = {
        "text": doc.decode("utf-8"), "n_tokens": ntok, "n_chunks": int(F.shape[0]),
        "rows_sparse": [sparse(r) for r in F], "ff": ff.tolist(), "spans": spans.tolist(),
        "query": DEMO_QUERY, "q_sparse": sparse(q), "qq": int((q.astype(np.int32) ** 2).sum()),
        "scores": rf1.scores(F, q).tolist(),
        "cos": [float(x) for x in rf1.cosine(rf1.scores(F, q), (q.astype(np.int32) ** 2).sum(), ff)],
    }

    # chunk-count rule at the window boundaries
    out["n_chunks_for"] = {str(t): rf1.n_chunks_for(t) for t in
                           [0, 1, 127, 128, 129, 143, 144, 145, 240, 241, 352, 353, 1000, 100000]}

    # a long document: 700 kept tokens, every 7th word is a stop-word that must not count
    words = []
    for i in range(700):
        words.append("w%d" % (i % 37))
        if i % 7 == 0:
            words.append(["The", "a", "AN"][i % 3])
    long_doc = " ".join(words).encode()
    F, ff, spans, ntok = rf1.featurize_doc(long_doc)
    out["long_doc"] = {"text": long_doc.decode(), "n_tokens": ntok, "n_chunks": int(F.shape[0]),
                       "rows_sparse": [sparse(r) for r in F], "ff": ff.tolist(), "spans": spans.tolist()}

    # saturation: 200 copies of one token in a single chunk-sized query -> min(tf,127)
    out["saturation"] = {"text": "zz " * 200, "q_sparse": sparse(rf1.query_vector(b"zz " * 200))}

    # hand-built tie / scope / tombstone / k>N cases on a tiny matrix
    Ft = np.zeros((12, 256), np.int8)
    for r in range(12):
        Ft[r, 5] = [3, 3, 3, 1, 0, 3, 2, 2, 9, 3, 3, 3][r]
        Ft[r, 9] = [0, 0, 1, 0, 0, 0, 1, 1, 0, 0, 0, 0][r]
    qt = np.zeros(256, np.int8); qt[5] = 2; qt[9] = 1
    seg = np.array([0, 1, 0, 0, 0, 1, 2, 0, 0xFFFFFFFF, 0, 1, 0], np.uint32)
    cases = []
    for name, scope, k, base in [("ties_one_store", [0], 10, 0), ("ties_two_stores", [0, 1], 10, 0),
                                 ("k_lt_n", [0, 1, 2], 3, 0), ("empty_scope", [7], 10, 0),
                                 ("no_scope", [], 10, 0), ("id_base", [0, 1], 5, 1000),
                                 ("tombstone_excluded", [0, 1, 2, 0xFFFFFFFF], 12, 0)]:
        ids, sc = rf1.score_topk(Ft, seg, qt, scope, k=k, id_base=base)
        cases.append({"name": name, "scope": scope, "k": k, "id_base": base,
                      "ids": [int(x) for x in ids], "scores": [int(x) for x in sc]})
    out["tiny"] = {"F_sparse": [sparse(r) for r in Ft], "q_sparse": sparse(qt), "store_seg": seg.tolist(),
                   "cases": cases}

    # synthetic generator: a few rows/queries and a 20k-row top-10
    out["mix64"] = [[s, a, b, int(rf1.mix64(s, a, b))] for s, a, b in
                    [(0, 0, 0), (1, 2, 3), (7, 10**9, 126), (0xA5, 2**40 + 17, 0), (2**63, 2**63, 2**63)]]
    rows = rf1.synth_rows(0, 0, 4, zb)
    out["synth"] = {"seed": 0, "rows_sparse_0_3": [sparse(r) for r in rows],
                    "rows_at_99999999": [sparse(r) for r in rf1.synth_rows(0, 99_999_999, 1, zb)],
                    "queries_sparse_0_2": [sparse(rf1.synth_query(0, i, zb)) for i in range(3)]}
    N = 20000
    topk = []
    for seed in (0, 1):
        Fs = rf1.synth_rows(seed, 0, N, zb)
        for qi in range(4):
            qs = rf1.synth_query(seed, qi, zb)
            ids, sc = rf1.score_topk(Fs, np.zeros(N, np.uint32), qs, [0])
            topk.append({"seed": seed, "qi": qi, "n_rows": N, "ids": [int(x) for x in ids],
                         "scores": [int(x) for x in sc]})
    out["synth_top10"] = topk
    out["zipf_bucket_sha_first64"] = [int(x) for x in zb[:64]]

    with open(os.path.join(HERE, "rf1_golden.json"), "w") as f:
        json.dump(out, f, indent=None, separators=(",", ":"), ensure_ascii=True)
    print("wrote rf1_golden.json", os.path.getsize(os.path.join(HERE, "rf1_golden.json")), "bytes")


if __name__ == "__main__":
    main()
