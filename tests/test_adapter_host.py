"""Host-side logic of the B200Rag adapter against the reference's wire contract
(tests/golden/config1_wire.json, generated from the reference's MockGeminiRag) with a scripted
engine double -- the real engine needs a GPU (tests/test_gpu_adapter.py)."""
import json
import os

import numpy as np
import pytest

from rag_foundation_b200 import adapter as ad


class ScriptedEngine:
    """Stands in for rag_foundation_b200.Engine: records calls, returns scripted hits."""

    def __init__(self):
        self.stores = {}
        self.next_row = 0
        self.ingested = []
        self.tombstoned = []
        self.dropped = []
        self.hits = []

    def open_store(self, name):
        return self.stores.setdefault(name, len(self.stores))

    def lookup_store(self, name):
        return self.stores.get(name)

    def drop_store(self, seg):
        self.dropped.append(seg)
        for k in [k for k, v in self.stores.items() if v == seg]:
            del self.stores[k]

    def ingest_text(self, seg, doc_id, data):
        words = data.split()
        n = 0 if not words else 1 + max(0, (len(words) - 128 + 111) // 112)
        spans = np.array([[0, len(data)]] * n, np.int64).reshape(n, 2)
        first = self.next_row
        self.next_row += n
        self.ingested.append((seg, doc_id, first, n))
        return first, n, spans

    def tombstone_doc(self, doc_id):
        self.tombstoned.append(doc_id)

    def search_text(self, text, scope, k, ranges=None):
        self.last_query = (text, list(scope), k)
        self.last_ranges = ranges
        ids = np.array([h[0] for h in self.hits[:k]], np.uint64)
        sc = np.array([h[1] for h in self.hits[:k]], np.int32)
        cs = np.array([0.5] * len(ids), np.float32)
        return ids, sc, cs, np.zeros(256, np.int8)


@pytest.fixture()
def rag():
    reg = ad.Registry(ScriptedEngine())
    return ad.B200Rag(registry=reg)


@pytest.fixture(scope="module")
def wire(golden_dir):
    return json.load(open(os.path.join(golden_dir, "config1_wire.json")))


def test_contents_to_text_matches_reference(wire):
    for case in wire["contents_to_text"]:
        assert ad.contents_to_text(case["contents"]) == case["text"]


def test_store_name_contract(rag, wire):
    name = rag.create_store("demo")
    assert name.startswith("fileSearchStores/")            # routes/stores.py:46
    assert len(name) == wire["store_name_len"]
    assert rag.create_store("demo") != name                # unique (models.py:66)
    assert rag.is_mock is True and rag.list_stores() == []


def test_upload_op_status_and_stream_protocol(rag, wire, tmp_path):
    store = rag.create_store("demo")
    p = tmp_path / "sample-report.md"
    p.write_text("RAG Foundation demo flow proves streaming citations work end to end")
    up = rag.upload_file(store, str(p), display_name="sample-report.md")
    assert up.operation_name.startswith("operations/") and len(up.operation_name) <= 255   # ingestion.py:226-230
    assert up.file_id.startswith("files/") and len(up.file_id) <= 255
    st = rag.op_status(up.operation_name)
    assert set(st) == set(wire["op_status"]) and st["done"] is True and st["error"] is None
    assert rag.op_status({"name": "operations/unknown"})["done"] is True
    with pytest.raises(ValueError):
        rag.op_status({})

    rag._reg.engine.hits = [(0, 11)]
    contents = [{"role": "user", "parts": [{"text": wire["demo_query"]}]}]
    chunks = list(rag.ask_stream(contents=contents, store_names=[store], metadata_filter=None, model="m", system=None))
    assert len(chunks) == wire["n_stream_chunks"]
    assert isinstance(chunks[0].text, str) and chunks[0].candidates is None
    assert chunks[1].text is None and len(chunks[1].candidates) == 1
    for ch in chunks:   # routes/chat.py:667-682 reads these
        assert ch.usage_metadata.prompt_token_count == 0 and ch.usage_metadata.candidates_token_count == 0
    assert rag._reg.engine.last_query[0] == wire["demo_query"].encode()

    cits = rag.extract_citations_from_response(chunks[1])
    assert len(cits) == 1 and sorted(cits[0]) == wire["citation_keys"]
    c = cits[0]
    assert c["index"] == 0 and c["source_type"] == "retrieved_context" and c["store"] == store
    assert c["title"] == "sample-report.md" and c["snippet"].startswith("RAG Foundation demo flow")
    assert c["uri"].startswith("chunk://" + store + "/") and c["uri"].endswith("#0")
    # the caller's frame (routes/chat.py:576-586) keeps its exact key set
    frame = {"type": "source-document", "sourceId": f"cit-{c['index']}", "mediaType": "file",
             "title": c.get("title") or c.get("uri") or "Source", "snippet": c.get("snippet")}
    assert set(frame) == set(wire["source_document_frames"][0])
    assert rag.ask(contents="q", store_names=[store], metadata_filter=None, model="m").candidates


def test_extract_citations_edge_cases_match_reference(wire):
    from types import SimpleNamespace as NS
    ex = ad.B200Rag.extract_citations_from_response
    assert ex(NS(candidates=[])) == wire["extract_edge_cases"]["empty_candidates"] == []
    assert ex(NS(candidates=[NS(grounding_metadata=None)])) == wire["extract_edge_cases"]["no_metadata"]
    assert ex(NS(candidates=[NS(grounding_metadata=NS(grounding_chunks=None))])) == wire["extract_edge_cases"]["no_chunks"]
    assert ex(object()) == []
    web = NS(candidates=[NS(grounding_metadata=NS(grounding_chunks=[NS(retrieved_context=None, web=NS(uri="u", title="t"))]))])
    assert ex(web) == [{"index": 0, "source_type": "web", "uri": "u", "title": "t", "snippet": None, "store": None}]


def test_stream_ids(rag, wire):
    a, b = rag.new_stream_ids()
    assert [len(a), len(b)] == wire["stream_id_lens"] and a != b


def test_rank_order_and_sidecar_lookup(rag, tmp_path):
    store = rag.create_store("s")
    texts = []
    for i in range(3):
        p = tmp_path / f"d{i}.txt"
        p.write_text(" ".join(f"doc{i}w{j}" for j in range(300)))   # 300 tokens -> 3 chunks each
        rag.upload_file(store, str(p), display_name=f"d{i}.txt")
        texts.append(p.read_text())
    rag._reg.engine.hits = [(7, 30), (0, 30), (5, 12), (99, 1)]      # chunk 99 does not exist -> dropped
    g = rag.retrieve("q", [store])
    assert [x["chunk_id"] for x in g] == [7, 0, 5]
    assert [x["title"] for x in g] == ["d2.txt", "d0.txt", "d1.txt"]
    assert g[0]["uri"].split("/")[-1] == "1#7" and g[2]["uri"].split("/")[-1] == "2#5"
    assert g[1]["score"] == 30 and len(g[0]["text"].encode()) <= ad.SNIPPET_MAX_BYTES


def test_unknown_and_empty_scope(rag):
    assert rag.retrieve("q", ["fileSearchStores/nope"]) == []
    chunks = list(rag.ask_stream(contents="q", store_names=[], metadata_filter=None, model="m"))
    assert len(chunks) == 2 and rag.extract_citations_from_response(chunks[1]) == []


def test_delete_paths(rag, tmp_path):
    store = rag.create_store("s")
    p = tmp_path / "d.txt"
    p.write_text("alpha beta gamma")
    up = rag.upload_file(store, str(p))
    rag.delete_document_from_store(store, 123, filename="d.txt", file_id=up.file_id)
    rag.delete_document_from_store(store, 123, filename="d.txt", file_id=up.file_id)   # idempotent
    assert rag._reg.engine.tombstoned == [1]
    rag.delete_document_from_store(store, 5, file_id="files/other")                     # unknown: no-op
    rag.delete_store(store)
    rag.delete_store(store)
    assert rag._reg.engine.dropped == [0]


def test_upload_failure_is_visible_in_op_status(rag, tmp_path):
    store = rag.create_store("s")
    p = tmp_path / "d.txt"
    p.write_text("x")

    def boom(*a, **k):
        raise RuntimeError("arena full")
    rag._reg.engine.ingest_text = boom
    with pytest.raises(RuntimeError):
        rag.upload_file(store, str(p))
    ops = list(rag._reg.ops.items())
    assert len(ops) == 1 and rag.op_status(ops[0][0])["error"] == "arena full"   # ingestion.py:132-133 raises on it


def test_metadata_helpers():
    m = ad.normalize_custom_metadata([{"key": "team", "string_value": "ops"}, {"key": "year", "numeric_value": 2024}, {"bad": 1}])
    assert m == {"team": "ops", "year": 2024}
    assert ad.normalize_custom_metadata({"a": 1}) == {"a": 1} and ad.normalize_custom_metadata(None) == {}
    assert ad.doc_matches(m, None) and ad.doc_matches(m, {})
    assert ad.doc_matches(m, {"team": "ops"}) and ad.doc_matches(m, {"team": ["dev", "ops"], "year": 2024})
    assert ad.doc_matches(m, {"year": "2024"})                      # upstream coerces values to strings or numbers
    assert not ad.doc_matches(m, {"team": "dev"}) and not ad.doc_matches(m, {"missing": "x"})
    assert not ad.doc_matches(m, "team = ops")                       # only the validated dict form is accepted


def test_metadata_filter_narrows_to_matching_documents(rag, tmp_path):
    store = rag.create_store("s")
    for i, team in enumerate(["ops", "dev", "ops"]):
        p = tmp_path / f"d{i}.txt"
        p.write_text(" ".join(f"w{j}" for j in range(200)))          # 2 chunks each
        rag.upload_file(store, str(p), display_name=f"d{i}.txt", custom_metadata=[{"key": "team", "string_value": team}])
    seen = {}

    def search_text(text, scope, k, ranges=None):
        seen["ranges"] = ranges
        return np.zeros(0, np.uint64), np.zeros(0, np.int32), np.zeros(0, np.float32), np.zeros(256, np.int8)
    rag._reg.engine.search_text = search_text
    rag.retrieve("q", [store], metadata_filter={"team": "ops"})
    assert seen["ranges"] == [(0, 2), (4, 6)]
    rag.retrieve("q", [store], metadata_filter=None)
    assert seen["ranges"] is None
    seen.clear()
    assert rag.retrieve("q", [store], metadata_filter={"team": "nobody"}) == [] and "ranges" not in seen
    list(rag.ask_stream(contents="q", store_names=[store], metadata_filter={"team": ["dev"]}, model="m"))
    assert seen["ranges"] == [(2, 4)]


def test_delete_forgets_the_document_and_upload_to_a_deleted_store_fails(rag, tmp_path, monkeypatch):
    store = rag.create_store("s")
    p = tmp_path / "d.txt"
    p.write_text("alpha beta gamma")
    up = rag.upload_file(store, str(p))
    reg = rag._reg
    assert len(reg.docs) == 1 and reg.first_chunks == [0]
    rag.delete_document_from_store(store, 1, file_id=up.file_id)
    assert reg.docs == {} and reg.doc_by_file == {} and reg.first_chunks == [] and reg.first_chunk_doc == []
    reg.engine.hits = [(0, 5)]
    assert rag.retrieve("alpha", [store]) == []                  # a stale hit on a forgotten chunk is dropped, not mis-attributed
    # the engine may hand the freed chunk ids to the next document: the lookup must find the NEW one
    reg.engine.next_row = 0
    up2 = rag.upload_file(store, str(p), display_name="second.txt")
    assert rag.retrieve("alpha", [store])[0]["title"] == "second.txt" and up2.file_id != up.file_id
    rag.delete_store(store)
    assert reg.docs == {} and store in reg.deleted_stores
    with pytest.raises(ValueError, match="unknown or deleted store"):
        rag.upload_file(store, str(p))
    with pytest.raises(ValueError, match="unknown or deleted store"):
        rag.upload_file("fileSearchStores/never-created", str(p))
    monkeypatch.setenv("RAG_B200_AUTOCREATE_STORES", "1")
    rag.upload_file("fileSearchStores/never-created", str(p))     # explicitly allowed
    with pytest.raises(ValueError):
        rag.upload_file(store, str(p))                            # a deleted store stays deleted


def test_metrics_series_of_the_reference_are_fed(rag, tmp_path):
    """backend/app/metrics.py:6-8 -- same series names and operation labels as GeminiRag."""
    from rag_foundation_b200 import metrics
    if metrics.SOURCE == "disabled":
        pytest.skip("prometheus_client is not installed")
    import sys
    ref = sys.modules.get("app.metrics")     # loaded when the reference's own test module ran first in this process
    if ref is not None:
        from prometheus_client import REGISTRY
    else:
        REGISTRY = metrics.REGISTRY

    def val(name, **labels):
        return REGISTRY.get_sample_value(name, labels) or 0.0
    before = {k: val("gemini_api_calls_total", operation=k[0], status=k[1])
              for k in [("upload", "ok"), ("upload", "error"), ("generate_stream", "ok"), ("generate", "ok")]}
    lat0 = val("gemini_api_latency_seconds_count", operation="generate_stream")
    store = rag.create_store("s")
    p = tmp_path / "d.txt"
    p.write_text("alpha beta")
    rag.upload_file(store, str(p))
    with pytest.raises(ValueError):
        rag.upload_file("fileSearchStores/nope", str(p))
    list(rag.ask_stream(contents="alpha", store_names=[store], metadata_filter=None, model="m"))
    rag.ask(contents="alpha", store_names=[store], metadata_filter=None, model="m")
    after = {k: val("gemini_api_calls_total", operation=k[0], status=k[1]) for k in before}
    assert all(after[k] == before[k] + 1 for k in before), (before, after)
    assert val("gemini_api_latency_seconds_count", operation="generate_stream") == lat0 + 1


def test_registry_sidecar_round_trip_is_plain_data(rag, tmp_path):
    store = rag.create_store("s")
    for i in range(3):
        p = tmp_path / f"d{i}.txt"
        p.write_text(" ".join(f"w{j}" for j in range(150 + i)))
        rag.upload_file(store, str(p), display_name=f"d{i}.txt", custom_metadata=[{"key": "n", "numeric_value": i}])
    reg = rag._reg
    reg.engine.save_snapshot = lambda path: open(path, "wb").write(b"snap")
    reg.save(str(tmp_path / "snap"))
    raw = open(tmp_path / "snap" / "sidecar.msgpack", "rb").read()
    import msgpack
    assert msgpack.unpackb(raw, raw=False)["version"] == 2      # not a pickle
    eng2 = ScriptedEngine()
    eng2.load_snapshot = lambda path: None
    reg2 = ad.Registry.load(eng2, str(tmp_path / "snap"))
    assert sorted(reg2.docs) == sorted(reg.docs) and reg2.first_chunks == reg.first_chunks and reg2.next_doc == reg.next_doc
    for k, d in reg.docs.items():
        d2 = reg2.docs[k]
        assert (d2.spans == d.spans).all() and d2.data == d.data and d2.meta == d.meta and d2.display_name == d.display_name
    assert reg2.ops == reg.ops
