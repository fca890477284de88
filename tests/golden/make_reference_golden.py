#!/usr/bin/env python3
"""Generate golden fixtures from the UNMODIFIED reference at /root/reference (run in the build
container only; the GPU box has no /root/reference, tests read the committed JSON).

  normalize_golden.json  <- scripts/benchmark/metrics.py:_normalize  (pins RF-1 step 1)
  config1_wire.json      <- backend/app/services/gemini_rag.py MockGeminiRag + get_rag_client,
                            driven as BASELINE.json configs[0] does: ingest docs/demo/sample-report.md,
                            ask the README.md:34 demo question.  Pins object shapes / wire contract.

The reference imports `sqlalchemy.engine.url.make_url` (backend/app/config.py:10, used only by a
production validator); sqlalchemy is not installed here, so a 3-line stub module is injected.
"""
import json
import os
import re
import sys
import types

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DEMO_QUERY = "What does this demo prove about the RAG engineering flow?"  # README.md:34


def normalize_cases():
    sys.path.insert(0, os.path.join(REF, "scripts", "benchmark"))
    import metrics  # the reference's own module

    doc = open(os.path.join(REF, "docs", "demo", "sample-report.md"), encoding="utf-8").read()
    texts = [
        DEMO_QUERY, doc, "", "   ", "a an the", "The Theory of an Anagram; a-b_c/d", "A", "AN apple, THE END.",
        "tabs\tand\nnewlines\r\nmixed", "snake_case kebab-case camelCase 123abc 4.56e7", "x" * 300,
        "e-mail: someone@example.com (urgent!!) #42", "the" * 5 + " thethe a1 an2 3the",
        "café naïve 中文 text", "trailing the", "the",
    ]
    return [{"text": t, "normalized": metrics._normalize(t)} for t in texts]


def metrics_cases():
    """scripts/benchmark/metrics.py citation_hit / extract_gold_doc_ids / p95 on fixed inputs (pins the
    grading functions tools/quality_eval.py restates)."""
    sys.path.insert(0, os.path.join(REF, "scripts", "benchmark"))
    import metrics
    recs = [{"gold_docs": ["A.md", "b.md"]}, {"supporting_docs": [{"doc_id": "x.txt"}, {"nope": 1}, "str", {"doc_id": ""}]},
            {"gold_docs": [], "supporting_docs": [{"doc_id": "y"}]}, {}, {"supporting_docs": None}]
    cits = [[], None, [{"title": "a.md"}], [{"doc_id": "B.MD", "title": "zzz"}], [{"sourceId": "cit-0", "title": "a.md"}],
            [{"uri": "chunk://s/1/0#5", "title": "a.md"}], [{"title": "q"}, {"title": "A.MD"}], [{"doc_id": "", "sourceId": None, "uri": "", "title": "b.md"}],
            [{"doc_id": 7}], [{}]]
    golds = [["A.md", "b.md"], [], ["7"], ["chunk://S/1/0#5"]]
    series = [[], [5.0], [1.0, 2.0], [3.0, 1.0, 2.0, 10.0], [float(i) for i in range(21)], [0.5] * 7 + [9.25]]
    return {"extract_gold_doc_ids": [{"rec": r, "want": metrics.extract_gold_doc_ids(r)} for r in recs],
            "citation_hit": [{"citations": c, "gold": g, "want": metrics.citation_hit(c, g)} for c in cits for g in golds],
            "p95": [{"values": v, "want": metrics.p95(v)} for v in series]}


def wire_case():
    os.environ.update(ENVIRONMENT="test", GEMINI_MOCK_MODE="true", JWT_SECRET="x" * 64,
                      GEMINI_API_KEY="fake-key-for-tests")
    sa = types.ModuleType("sqlalchemy"); eng = types.ModuleType("sqlalchemy.engine")
    url = types.ModuleType("sqlalchemy.engine.url"); url.make_url = lambda s: s
    sys.modules.update({"sqlalchemy": sa, "sqlalchemy.engine": eng, "sqlalchemy.engine.url": url})
    sys.path.insert(0, os.path.join(REF, "backend"))
    from app.services import gemini_rag as g

    rag = g.get_rag_client()
    store = rag.create_store("demo")
    up = rag.upload_file(store, os.path.join(REF, "docs", "demo", "sample-report.md"), display_name="sample-report.md")
    st = rag.op_status(up.operation_name)
    contents = [{"role": "user", "parts": [{"text": DEMO_QUERY}]}]
    chunks = list(rag.ask_stream(contents=contents, store_names=[store], metadata_filter=None,
                                 model="gemini-2.5-flash", system=None))
    cits = rag.extract_citations_from_response(chunks[-1])
    ids = rag.new_stream_ids()
    hexes = re.compile(r"[0-9a-f]{32}")
    mask = lambda s: hexes.sub("<hex32>", s)
    frames = []
    for c in cits:  # routes/chat.py:576-586 restated (chat.py itself needs sqlalchemy ORM + jose)
        frames.append({"type": "source-document", "sourceId": f"cit-{c['index']}", "mediaType": "file",
                       "title": c.get("title") or c.get("uri") or "Source", "snippet": c.get("snippet")})
    return {
        "adapter_class": type(rag).__name__, "is_mock": rag.is_mock,
        "store_name": mask(store), "store_name_len": len(store),
        "upload": {"operation_name": mask(up.operation_name), "file_id": mask(up.file_id)},
        "op_status": {**st, "name": mask(st["name"])},
        "n_stream_chunks": len(chunks),
        "chunk0": {"text": chunks[0].text, "candidates": chunks[0].candidates,
                   "usage": [chunks[0].usage_metadata.prompt_token_count, chunks[0].usage_metadata.candidates_token_count]},
        "chunk1": {"text": chunks[1].text, "n_candidates": len(chunks[1].candidates),
                   "usage": [chunks[1].usage_metadata.prompt_token_count, chunks[1].usage_metadata.candidates_token_count]},
        "citations": [{**c, "store": mask(c["store"])} for c in cits],
        "citation_keys": sorted(cits[0].keys()),
        "source_document_frames": frames,
        "stream_id_lens": [len(ids[0]), len(ids[1])],
        "contents_to_text": [
            {"contents": c, "text": g.MockGeminiRag._contents_to_text(c)} for c in (
                "plain", contents, [{"role": "user", "parts": [{"text": "  first "}]}, {"role": "model", "parts": [{"text": ""}]}],
                ["a", "  "], [], [{"role": "user", "parts": []}], [{"role": "user", "parts": [{"text": "q1"}]}, "tail "])
        ],
        "extract_edge_cases": {
            "empty_candidates": g.GeminiRag.extract_citations_from_response(types.SimpleNamespace(candidates=[])),
            "no_metadata": g.GeminiRag.extract_citations_from_response(
                types.SimpleNamespace(candidates=[types.SimpleNamespace(grounding_metadata=None)])),
            "no_chunks": g.GeminiRag.extract_citations_from_response(types.SimpleNamespace(
                candidates=[types.SimpleNamespace(grounding_metadata=types.SimpleNamespace(grounding_chunks=None))])),
        },
        "demo_query": DEMO_QUERY,
    }


if __name__ == "__main__":
    with open(os.path.join(HERE, "normalize_golden.json"), "w") as f:
        json.dump(normalize_cases(), f, indent=1, ensure_ascii=True)
    with open(os.path.join(HERE, "config1_wire.json"), "w") as f:
        json.dump(wire_case(), f, indent=1, ensure_ascii=True)
    with open(os.path.join(HERE, "benchmark_metrics_golden.json"), "w") as f:
        json.dump(metrics_cases(), f, indent=1, ensure_ascii=True)
    print("wrote normalize_golden.json, config1_wire.json, benchmark_metrics_golden.json")
