#!/usr/bin/env python3
"""Retrieval-quality evaluation of the B200 retriever, RF-1 (tf) against RF-1w (idf), graded the way
the reference's benchmark harness grades citations.

The reference harness (scripts/benchmark/run_benchmark.py:210-215, 365) asks each question of a
questions.jsonl over HTTP and scores `citation_hit(citations, gold_doc_ids)`
(scripts/benchmark/metrics.py:73-92) with gold ids from `extract_gold_doc_ids` (:66-70).  Here the
same record format and the same two functions (restated below, pinned to fixtures generated from
the reference's own module: tests/golden/benchmark_metrics_golden.json) are driven straight
through the adapter (`B200Rag.upload_file` / `ask_stream` / `extract_citations_from_response`), on a
seeded synthetic labelled set -- no network, no dataset: documents are Zipf word streams, a question
is a handful of words of one window of its gold document plus off-topic noise words.

A citation is matched on `doc_id` = the citation's `title`, i.e. the upload's display name, which is
the identity the gold labels use (the harness matches doc_id / sourceId / uri / title in that order).

  python tools/quality_eval.py --docs 2000 --questions 500  > profiles/quality_eval_rNN.json
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import tempfile
import time
from typing import Iterable, Optional, Sequence

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


# ---- the reference's grading functions (scripts/benchmark/metrics.py), restated -------------------
def extract_gold_doc_ids(rec: dict) -> list:
    """metrics.py:66-70: `gold_docs` if present, else the doc_id of every supporting_docs entry."""
    if rec.get("gold_docs"):
        return list(rec["gold_docs"])
    return [d["doc_id"] for d in (rec.get("supporting_docs") or []) if isinstance(d, dict) and d.get("doc_id")]


def citation_hit(citations: Optional[Iterable[dict]], gold_doc_ids: Sequence[str]) -> Optional[int]:
    """metrics.py:73-92: None without gold ids; 1 when any citation's first non-empty field of
    (doc_id, sourceId, uri, title) equals a gold id, case-insensitively; else 0."""
    if not gold_doc_ids:
        return None
    gold = {str(g).lower() for g in gold_doc_ids}
    for c in citations or []:
        cand = str(c.get("doc_id") or c.get("sourceId") or c.get("uri") or c.get("title") or "").lower()
        if cand and cand in gold:
            return 1
    return 0


def p95(values: Sequence[float]) -> float:
    """metrics.py:99-109: linear interpolation at rank 0.95 (n - 1)."""
    if not values:
        return 0.0
    vals = sorted(values)
    pos = (len(vals) - 1) * 0.95
    lo, hi = math.floor(pos), math.ceil(pos)
    return vals[int(pos)] if lo == hi else vals[lo] + (vals[hi] - vals[lo]) * (pos - lo)


# ---- seeded labelled set -------------------------------------------------------------------------
def make_labelled_set(n_docs: int, n_questions: int, seed: int = 0, doc_tokens=(300, 700), q_own: int = 6, q_noise: int = 4):
    """-> (docs {doc_id: bytes}, questions [records in the reference's questions.jsonl shape])."""
    from rag_foundation_b200.engine import load_zipf_vocab
    vocab = load_zipf_vocab().astype(np.int64)            # 65536-entry inverse CDF of Zipf(1.07) over 50 000 ids
    rng = np.random.default_rng(seed)
    docs, toks = {}, {}
    for d in range(n_docs):
        n = int(rng.integers(doc_tokens[0], doc_tokens[1]))
        ids = vocab[rng.integers(0, 65536, n)]
        name = f"doc-{d:05d}.txt"
        toks[name] = ids
        docs[name] = (" ".join(f"w{t}" for t in ids) + "\n").encode()
    names = sorted(docs)
    questions = []
    for i in range(n_questions):
        gold = names[int(rng.integers(0, n_docs))]
        ids = toks[gold]
        start = int(rng.integers(0, max(1, len(ids) - 40)))
        own = rng.choice(ids[start:start + 40], size=q_own, replace=False)
        noise = vocab[rng.integers(0, 65536, q_noise)]
        words = [f"w{t}" for t in np.concatenate([own, noise])]
        rng.shuffle(words)
        questions.append({"id": f"q{i + 1}", "question": "Which report mentions " + " ".join(words) + "?", "answer": "",
                          "store": "bench-quality", "supporting_docs": [{"doc_id": gold, "span_start": start, "span_end": start + 40}]})
    return docs, questions


# ---- what-if model on the CPU: the same RF-1 / RF-1w arithmetic at other feature widths ------------
def model_dim_sweep(docs: dict, questions: list, dims=(256, 1024, 4096), top_k: int = 10) -> dict:
    """NOT the product path: a numpy model of RF-1 (tf) and RF-1w (idf) with D = 256 / 1024 / 4096
    hash buckets on the same labelled set, to show how much of the miss rate is bucket collisions
    (engines run D = 256, 512 or 1024 -- `--dim`; the model's row for the engine's width must equal
    the GPU numbers exactly, tests/test_gpu_wide_rows.py).
    Documents here are space-separated lower-case words without stop-words, so tokenising is split()."""
    def fnv(tok: bytes) -> int:
        h = 0x811C9DC5
        for b in tok:
            h = ((h ^ b) * 0x01000193) & 0xFFFFFFFF
        return h

    cache = {}

    def hashes(words):
        return np.fromiter((cache.setdefault(w, fnv(w)) for w in words), dtype=np.int64, count=len(words))

    chunk_doc, chunk_hash = [], []
    for name in sorted(docs):
        h = hashes(docs[name].split())
        w = 0
        while w == 0 or 112 * w + 16 < len(h):
            chunk_doc.append(name)
            chunk_hash.append(h[112 * w:112 * w + 128])
            w += 1
    stop = {b"a", b"an", b"the"}
    q_hash = []
    for rec in questions:
        toks = "".join(c if c.isalnum() and c.isascii() else " " for c in rec["question"].lower()).encode().split()
        q_hash.append(hashes([t for t in toks if t not in stop]))
    out = {}
    for D in dims:
        F = np.zeros((len(chunk_hash), D), np.int32)
        for i, h in enumerate(chunk_hash):
            np.add.at(F[i], h & (D - 1), 1)
        F = np.minimum(F, 127)
        n = F.shape[0]
        df = (F > 0).sum(axis=0)
        r = ((n + 1) * 256) // (df + 1)
        lg = np.floor(np.log2(r)).astype(np.int64)
        wgt = np.minimum(4 + 4 * (lg - 8) + ((r >> (lg - 2)) & 3), 31)
        res = {}
        for mode in ("tf", "idf"):
            hit = 0
            for rec, h in zip(questions, q_hash):
                q = np.zeros(D, np.int32)
                np.add.at(q, h & (D - 1), 1)
                q = np.minimum(q, 127)
                if mode == "idf":
                    q = np.minimum(q * wgt, 127)
                s = F @ q
                order = np.lexsort((np.arange(n), -s))[:top_k]
                gold = set(extract_gold_doc_ids(rec))
                hit += any(chunk_doc[i] in gold for i in order)
            res[mode] = hit / len(questions)
        out[str(D)] = res
    return out


def run_eval(rag_for, docs: dict, questions: list, top_k: int = 10) -> dict:
    """rag_for(scoring) -> B200Rag over ONE shared registry.  Uploads once, asks every question under
    both scoring rules.  Returns the summary the reference harness would write (per scoring rule)."""
    rag = rag_for("tf")
    store = rag.create_store("bench-quality")
    t0 = time.perf_counter()
    with tempfile.TemporaryDirectory() as tmp:
        for name, data in docs.items():
            path = os.path.join(tmp, name)
            with open(path, "wb") as f:
                f.write(data)
            up = rag.upload_file(store, path, display_name=name)
            st = rag.op_status(up.operation_name)
            assert st["done"] and not st["error"], st
    ingest_s = time.perf_counter() - t0
    out = {"docs": len(docs), "questions": len(questions), "top_k": top_k, "corpus_bytes": sum(map(len, docs.values())),
           "ingest_s": ingest_s, "scoring": {}}
    for scoring in ("tf", "idf"):
        rag = rag_for(scoring)
        hits, hits1, rr, lat = [], [], [], []
        for rec in questions:
            gold = extract_gold_doc_ids(rec)
            contents = [{"role": "user", "parts": [{"text": rec["question"]}]}]
            t0 = time.perf_counter()
            final = None
            for chunk in rag.ask_stream(contents=contents, store_names=[store], metadata_filter=None, model="eval"):
                if chunk.candidates:
                    final = chunk
            cits = rag.extract_citations_from_response(final)[:top_k]
            lat.append((time.perf_counter() - t0) * 1e3)
            graded = [{"doc_id": c["title"], "uri": c["uri"], "title": c["title"]} for c in cits]
            hits.append(citation_hit(graded, gold))
            hits1.append(citation_hit(graded[:1], gold))
            rank = next((i for i, c in enumerate(graded) if citation_hit([c], gold)), None)
            rr.append(0.0 if rank is None else 1.0 / (rank + 1))
        out["scoring"][scoring] = {"citation_hit_rate": float(np.mean(hits)), "hit_at_1": float(np.mean(hits1)),
                                   "mrr": float(np.mean(rr)), "latency_ms_mean": float(np.mean(lat)), "latency_ms_p95": p95(lat)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=2000)
    ap.add_argument("--questions", type=int, default=500)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--top-k", type=int, default=10)
    ap.add_argument("--dim-sweep", action="store_true", help="also run the CPU what-if model at D = 256 / 1024 / 4096")
    ap.add_argument("--dim", type=int, default=256, help="features per chunk row of the engine (256, 512, 1024)")
    args = ap.parse_args()
    from rag_foundation_b200 import Engine
    from rag_foundation_b200 import adapter as ad
    docs, questions = make_labelled_set(args.docs, args.questions, args.seed)
    reg = ad.Registry(Engine(capacity_rows=max(4096, args.docs * 8), dim=args.dim))
    try:
        res = run_eval(lambda scoring: ad.B200Rag(registry=reg, top_k=args.top_k, scoring=scoring), docs, questions, args.top_k)
    finally:
        reg.engine.close()
    res["seed"] = args.seed
    res["dim"] = args.dim
    if args.dim_sweep:
        res["cpu_model_citation_hit_rate_by_dim"] = model_dim_sweep(docs, questions, top_k=args.top_k)
    res["note"] = ("synthetic labelled set: Zipf(1.07) word streams; a question = 6 words of one 40-word window of its gold "
                   "document + 4 off-topic words; graded with the reference harness's citation_hit on doc_id = citation title")
    print(json.dumps(res))


if __name__ == "__main__":
    main()
