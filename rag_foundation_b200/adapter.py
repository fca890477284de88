"""B200Rag -- drop-in for the object `get_rag_client()` returns in the reference.

Mirrors the duck-typed adapter protocol of backend/app/services/gemini_rag.py (paths relative to
/root/reference): GeminiRag :242-599, MockGeminiRag :602-718, get_rag_client :721-725.  Same method
names, argument meaning, return shapes and error behaviour, so the SSE chat route
(routes/chat.py:499-505, 576-586) and the ARQ ingest worker (services/ingestion.py:45-52,106-139)
need no change.  Where the reference calls a remote service (or, in mock mode, returns one canned
citation), this adapter featurises on the GPU at upload and runs the score/top-k kernels at query
time; see include/rf_b200.h for the C-ABI underneath.

The adapter object is re-created on every request (gemini_rag.py:721-725), so it is a stateless
facade over one process-global Registry (engine handle + chunk-text sidecar).
"""
from __future__ import annotations

import bisect
import logging
import os
import threading
import time
import uuid
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Any, Dict, Generator, List, Optional, Sequence, Union

import numpy as np

from .engine import Engine
from .metrics import gemini_calls_total, gemini_latency

TOPK_DEFAULT = int(os.environ.get("RAG_TOPK", "10"))
SNIPPET_MAX_BYTES = 1024


@dataclass
class UploadResult:  # gemini_rag.py:30-33
    operation_name: str
    file_id: str | None = None


@dataclass
class _Doc:
    doc_id: int
    store_name: str
    display_name: str
    file_id: str
    first_chunk: int
    n_chunks: int
    spans: np.ndarray
    data: bytes
    custom_metadata: Any = None
    deleted: bool = False
    meta: Dict[str, Any] = field(default_factory=dict)   # custom_metadata flattened to {key: value}


class Registry:
    """Process-global state behind every B200Rag facade: the engine and the chunk sidecar."""

    SIDECAR = "sidecar.msgpack"

    def __init__(self, engine: Engine):
        self.engine = engine
        self.lock = threading.RLock()
        # an upload (engine ingest + registration here) and a snapshot exclude each other, so the
        # engine file and the sidecar written by save() always describe the same set of documents
        self.ingest_lock = threading.Lock()
        self.docs: Dict[int, _Doc] = {}
        self.doc_by_file: Dict[str, int] = {}
        self.first_chunks: List[int] = []   # sorted, parallel to first_chunk_doc
        self.first_chunk_doc: List[int] = []
        self.ops: Dict[str, dict] = {}
        self.deleted_stores: set = set()
        self.next_doc = 1

    def chunk_to_doc(self, chunk_id: int) -> Optional[tuple]:
        with self.lock:
            i = bisect.bisect_right(self.first_chunks, chunk_id) - 1
            if i < 0:
                return None
            d = self.docs.get(self.first_chunk_doc[i])
            if d is None or not (d.first_chunk <= chunk_id < d.first_chunk + d.n_chunks):
                return None
            return d, chunk_id - d.first_chunk

    def add_doc(self, doc: _Doc) -> None:
        with self.lock:
            self.docs[doc.doc_id] = doc
            self.doc_by_file[doc.file_id] = doc.doc_id
            if doc.n_chunks:
                i = bisect.bisect_right(self.first_chunks, doc.first_chunk)
                self.first_chunks.insert(i, doc.first_chunk)
                self.first_chunk_doc.insert(i, doc.doc_id)

    def forget_doc(self, doc_id: int) -> Optional[_Doc]:
        """Drop a deleted document entirely -- its bytes, spans, metadata and index entries (the
        reference's cleanup path, services/cleanup.py:37,66, means the data to be gone; the engine may hand
        its chunk ids to a later document)."""
        with self.lock:
            doc = self.docs.pop(doc_id, None)
            if doc is None:
                return None
            self.doc_by_file.pop(doc.file_id, None)
            if doc.n_chunks:
                i = bisect.bisect_left(self.first_chunks, doc.first_chunk)
                while i < len(self.first_chunks) and self.first_chunks[i] == doc.first_chunk:
                    if self.first_chunk_doc[i] == doc_id:
                        del self.first_chunks[i]
                        del self.first_chunk_doc[i]
                        break
                    i += 1
            return doc

    # ---- durability: engine snapshot + the chunk-text sidecar ----
    def save(self, directory: str) -> None:
        """Engine snapshot(s) + sidecar, both written to a temporary name, fsync'ed and renamed.  The sidecar
        is msgpack (plain data: nothing executable is read back at load)."""
        import msgpack
        os.makedirs(directory, exist_ok=True)
        with self.ingest_lock, self.lock:
            self.engine.save_snapshot(os.path.join(directory, "index.rfsnap"))
            docs = [{"doc_id": d.doc_id, "store_name": d.store_name, "display_name": d.display_name, "file_id": d.file_id,
                     "first_chunk": d.first_chunk, "n_chunks": d.n_chunks,
                     "spans": np.ascontiguousarray(d.spans, dtype="<i8").tobytes(), "data": d.data,
                     "custom_metadata": d.custom_metadata, "meta": d.meta} for d in self.docs.values()]
            side = {"version": 2, "docs": docs, "ops": self.ops, "next_doc": self.next_doc,
                    "deleted_stores": sorted(self.deleted_stores)}
            tmp = os.path.join(directory, self.SIDECAR + ".tmp")
            with open(tmp, "wb") as f:
                f.write(msgpack.packb(side, use_bin_type=True))
                f.flush()
                os.fsync(f.fileno())
            os.replace(tmp, os.path.join(directory, self.SIDECAR))

    @classmethod
    def load(cls, engine: Engine, directory: str) -> "Registry":
        """`engine` must be freshly created (empty)."""
        import msgpack
        engine.load_snapshot(os.path.join(directory, "index.rfsnap"))
        reg = cls(engine)
        with open(os.path.join(directory, cls.SIDECAR), "rb") as f:
            side = msgpack.unpackb(f.read(), raw=False, strict_map_key=False)
        if side.get("version") != 2:
            raise ValueError("unknown sidecar version")
        for d in side["docs"]:
            spans = np.frombuffer(d["spans"], dtype="<i8").reshape(-1, 2).copy()
            reg.add_doc(_Doc(int(d["doc_id"]), d["store_name"], d["display_name"], d["file_id"], int(d["first_chunk"]),
                             int(d["n_chunks"]), spans, bytes(d["data"]), d.get("custom_metadata"), meta=d.get("meta") or {}))
        reg.ops = dict(side["ops"])
        reg.next_doc = int(side["next_doc"])
        reg.deleted_stores = set(side.get("deleted_stores") or [])
        return reg


_registry: Optional[Registry] = None
_registry_lock = threading.Lock()


def get_registry() -> Registry:
    """Lazily create the process-global engine from env flags (kept beside, not inside, the
    reference's Settings -- SURVEY.md §5): RAG_B200_CAPACITY_ROWS, RAG_B200_DIM, RAG_B200_DEVICE, or RAG_B200_DEVICES=0,1,..
    for an engine group over several GPUs."""
    global _registry
    with _registry_lock:
        if _registry is None:
            cap = int(os.environ.get("RAG_B200_CAPACITY_ROWS", str(4_000_000)))
            n_ctx = int(os.environ.get("RAG_B200_CONTEXTS", "16"))
            dim = int(os.environ.get("RAG_B200_DIM", "256"))     # 256, 512 or 1024 features per chunk
            devices = os.environ.get("RAG_B200_DEVICES", "").strip()
            if devices:
                # several GPUs behind ONE adapter: an engine per device in this process, stores placed by
                # RAG_B200_PLACEMENT ("store": whole stores per GPU; "spread": documents round-robin, so one
                # huge store is sharded by chunk); capacity is per device
                from .engine import EngineGroup
                devs = [int(x) for x in devices.split(",") if x.strip() != ""]
                _registry = Registry(EngineGroup(devs, capacity_rows=cap, n_contexts=n_ctx,
                                                 placement=os.environ.get("RAG_B200_PLACEMENT", "store"), dim=dim))
            else:
                dev = int(os.environ.get("RAG_B200_DEVICE", "0"))
                _registry = Registry(Engine(capacity_rows=cap, device=dev, n_contexts=n_ctx, dim=dim))
        return _registry


def set_registry(reg: Optional[Registry]) -> None:
    global _registry
    with _registry_lock:
        _registry = reg


def contents_to_text(contents: Any) -> str:
    """Last non-empty user turn; same rule as MockGeminiRag._contents_to_text (gemini_rag.py:640-654)."""
    if isinstance(contents, str):
        return contents
    if isinstance(contents, list):
        for item in reversed(contents):
            if isinstance(item, str) and item.strip():
                return item.strip()
            if isinstance(item, dict):
                parts = item.get("parts")
                if isinstance(parts, list) and parts and isinstance(parts[0], dict):
                    text = parts[0].get("text")
                    if isinstance(text, str) and text.strip():
                        return text.strip()
    return str(contents)


def normalize_custom_metadata(custom_metadata: Any) -> Dict[str, Any]:
    """upload_file's `custom_metadata` (gemini_rag.py:314: a list of {"key": k, "string_value" |
    "numeric_value": v} entries, or a plain dict) -> {key: value}."""
    out: Dict[str, Any] = {}
    if isinstance(custom_metadata, dict):
        out.update(custom_metadata)
    elif isinstance(custom_metadata, (list, tuple)):
        for item in custom_metadata:
            if isinstance(item, dict) and "key" in item:
                for vk in ("string_value", "numeric_value", "value"):
                    if vk in item:
                        out[str(item["key"])] = item[vk]
                        break
    return out


def doc_matches(meta: Dict[str, Any], metadata_filter: Optional[Dict[str, Any]]) -> bool:
    """`metadata_filter` as validated upstream (routes/chat.py:295-335): {key: scalar | [scalars]};
    every key must match (a list means any-of); documents without the key do not match."""
    if not metadata_filter:
        return True
    if not isinstance(metadata_filter, dict):
        return False
    for key, want in metadata_filter.items():
        if key not in meta:
            return False
        have = meta[key]
        wants = want if isinstance(want, (list, tuple, set)) else [want]
        if not any(have == w or str(have) == str(w) for w in wants):
            return False
    return True


def _get_response_name(response: Any, *, context: str) -> str:  # gemini_rag.py:96-102
    if isinstance(response, str):
        return response
    name = response.get("name") if isinstance(response, dict) else getattr(response, "name", None)
    if not name:
        raise ValueError(f"Missing name in {context} response")
    return name


def stream_lead(grounding: Sequence[dict]) -> str:
    """Text of the first stream chunk (the mock's is "[mock-mode] " + question, gemini_rag.py:676-683)."""
    lead = grounding[0]["text"].strip().splitlines()[0][:200] if grounding and grounding[0]["text"].strip() else "no matching passages"
    return f"[b200-retrieval] {lead}"


def build_final_response(grounding: Sequence[dict]) -> Any:
    """Final stream chunk, shaped like MockGeminiRag._mock_response (gemini_rag.py:704-718) but
    with one grounding chunk per retrieved citation, in rank order."""
    usage = SimpleNamespace(prompt_token_count=0, candidates_token_count=0)
    chunks = [SimpleNamespace(retrieved_context=SimpleNamespace(uri=g["uri"], title=g["title"], text=g["text"],
                                                                file_search_store=g["file_search_store"],
                                                                score=g.get("score"), cosine=g.get("cosine"),
                                                                chunk_id=g.get("chunk_id")), web=None)
              for g in grounding]
    candidate = SimpleNamespace(grounding_metadata=SimpleNamespace(grounding_chunks=chunks), usage_metadata=usage)
    return SimpleNamespace(text=None, candidates=[candidate], usage_metadata=usage)


class B200Rag:
    """GeminiRag-shaped retriever backed by the B200 engine."""

    def __init__(self, registry: Optional[Registry] = None, top_k: int = TOPK_DEFAULT,
                 scoring: Optional[str] = None) -> None:
        self._reg = registry or get_registry()
        self.is_mock = True          # main.py:385 skips the remote health probe when truthy
        self.is_b200 = True
        self.top_k = top_k
        # "tf" = RF-1, "idf" = RF-1w (query buckets weighted by the scope's inverse document frequency)
        self.scoring = (scoring or os.environ.get("RAG_B200_SCORING", "tf")).lower()
        if self.scoring not in ("tf", "idf"):
            raise ValueError(f"RAG_B200_SCORING must be 'tf' or 'idf', not {self.scoring!r}")

    # -------- Stores (gemini_rag.py:268-304, 607-612, 696-697) --------
    def list_stores(self) -> List[Any]:
        return []

    def create_store(self, display_name: str) -> str:
        name = f"fileSearchStores/b200-{uuid.uuid4().hex}"   # routes/stores.py:46 requires this prefix
        self._reg.engine.open_store(name)
        return name

    def delete_store(self, store_name: str) -> None:
        reg = self._reg
        seg = reg.engine.lookup_store(store_name)
        if seg is None:
            return
        reg.engine.drop_store(seg)
        with reg.lock:
            reg.deleted_stores.add(store_name)
            for doc_id in [d.doc_id for d in reg.docs.values() if d.store_name == store_name]:
                reg.forget_doc(doc_id)

    # -------- Upload & operations (gemini_rag.py:307-352, 426-460, 614-638) --------
    def upload_file(self, store_name: str, file_path: str, *, display_name: Optional[str] = None,
                    custom_metadata: Optional[List[Dict[str, Union[str, float, int]]]] = None,
                    chunking_config: Optional[Dict] = None) -> UploadResult:
        with open(file_path, "rb") as f:
            data = f.read()
        return self.upload_bytes(store_name, data, display_name=display_name or os.path.basename(file_path),
                                 custom_metadata=custom_metadata)

    def upload_bytes(self, store_name: str, data: bytes, *, display_name: str,
                     custom_metadata: Optional[Any] = None) -> UploadResult:
        """upload_file once the document bytes are in memory (what the engine daemon receives)."""
        reg = self._reg
        start = time.perf_counter()
        op_name = f"operations/b200-{uuid.uuid4().hex}"
        try:
            seg = reg.engine.lookup_store(store_name)
            if seg is None:
                # a store this engine has never seen (created before a restart without a snapshot) is an
                # error unless RAG_B200_AUTOCREATE_STORES=1; a store that was deleted here always is
                if store_name in reg.deleted_stores or os.environ.get("RAG_B200_AUTOCREATE_STORES", "0") != "1":
                    raise ValueError(f"unknown or deleted store {store_name!r}")
                seg = reg.engine.open_store(store_name)
            with reg.ingest_lock:
                with reg.lock:
                    doc_id = reg.next_doc
                    reg.next_doc += 1
                file_id = f"files/b200-{doc_id:016x}"
                first, n_chunks, spans = reg.engine.ingest_text(seg, doc_id, data)
                reg.add_doc(_Doc(doc_id, store_name, display_name, file_id, first, n_chunks, spans, data, custom_metadata,
                                 meta=normalize_custom_metadata(custom_metadata)))
                with reg.lock:
                    reg.ops[op_name] = {"done": True, "error": None, "n_chunks": n_chunks, "file_id": file_id}
            gemini_calls_total.labels("upload", "ok").inc()
            return UploadResult(operation_name=op_name, file_id=file_id)
        except Exception as exc:
            with reg.lock:
                reg.ops[op_name] = {"done": True, "error": str(exc)}
            gemini_calls_total.labels("upload", "error").inc()
            raise
        finally:
            gemini_latency.labels("upload").observe(time.perf_counter() - start)

    def op_status(self, op_name: str | dict) -> dict[str, Any]:
        name = _get_response_name(op_name, context="operation status request")
        with self._reg.lock:
            op = self._reg.ops.get(name)
        if op is None:   # unknown operations read as finished, like the mock (gemini_rag.py:631-638)
            return {"name": name, "done": True, "metadata": {}, "error": None}
        return {"name": name, "done": bool(op["done"]), "metadata": {k: v for k, v in op.items() if k in ("n_chunks", "file_id")},
                "error": op["error"]}

    def delete_document_from_store(self, store_name: str, document_id: int, filename: str | None = None,
                                   file_id: str | None = None) -> None:
        reg = self._reg
        with reg.lock:
            doc_id = reg.doc_by_file.get(file_id or "")
            doc = reg.docs.get(doc_id) if doc_id is not None else None
            if doc is None or doc.deleted:
                return
            doc.deleted = True          # no second tombstone from a concurrent delete
        if doc.n_chunks:
            reg.engine.tombstone_doc(doc.doc_id)   # masked in the index first, forgotten here after
        reg.forget_doc(doc.doc_id)

    # -------- Query (gemini_rag.py:472-551, 656-694) --------
    def retrieve(self, text: str, store_names: Sequence[str], k: Optional[int] = None,
                 metadata_filter: Optional[Dict[str, Any]] = None) -> List[dict]:
        """Top-k grounding contexts for `text` within `store_names`, rank order.  A metadata filter
        narrows the scan to the chunk ranges of the matching documents (second mask)."""
        reg = self._reg
        segs = []
        for s in store_names:
            seg = reg.engine.lookup_store(s)
            if seg is not None and seg not in segs:
                segs.append(seg)
        if not segs:
            return []
        ranges = None
        if metadata_filter:
            names = set(store_names)
            with reg.lock:
                ranges = sorted((d.first_chunk, d.first_chunk + d.n_chunks) for d in reg.docs.values()
                                if not d.deleted and d.n_chunks and d.store_name in names and doc_matches(d.meta, metadata_filter))
            if not ranges:
                return []
        weights = reg.engine.scope_weights(segs) if self.scoring == "idf" else None
        if weights is None:   # (keeps the call shape of engines that predate the weighted variant)
            ids, scores, cos, _q = reg.engine.search_text(text.encode("utf-8"), segs, k or self.top_k, ranges=ranges)
        else:
            ids, scores, cos, _q = reg.engine.search_text(text.encode("utf-8"), segs, k or self.top_k, ranges=ranges,
                                                          weights=weights)
        return self.describe_chunks(ids.tolist(), scores.tolist(), cos.tolist(), store_names)

    def describe_chunks(self, ids: Sequence[int], scores: Sequence[int], cos: Sequence[float],
                        store_names: Optional[Sequence[str]] = None) -> List[dict]:
        """Chunk ids (rank order) -> grounding contexts: document, chunk number, snippet.  Also what a process that ran
        the search itself on the shared arena (server.RemoteB200Rag in attach mode) asks the daemon for.
        `store_names`: the request's scope.  Chunk ids are reused after deletes, so between the GPU's tenant test and this
        lookup an id can have passed to a document of ANOTHER store (delete + ingest racing one search): such a hit is
        dropped here -- the last line of the tenant mask (security/tenant.py:28-47 decides the scope)."""
        reg = self._reg
        allowed = None if store_names is None else set(store_names)
        out = []
        for gid, sc, c in zip(ids, scores, cos):
            hit = reg.chunk_to_doc(int(gid))
            if hit is None:
                continue
            doc, chunk_no = hit
            if allowed is not None and doc.store_name not in allowed:
                continue
            b0, b1 = int(doc.spans[chunk_no, 0]), int(doc.spans[chunk_no, 1])
            snippet = doc.data[b0:min(b1, b0 + SNIPPET_MAX_BYTES)].decode("utf-8", "replace")
            out.append({"uri": f"chunk://{doc.store_name}/{doc.doc_id}/{chunk_no}#{gid}", "title": doc.display_name,
                        "text": snippet, "file_search_store": doc.store_name, "score": int(sc), "cosine": float(c),
                        "chunk_id": int(gid)})
        return out

    def ask(self, *, contents: Any, store_names: Sequence[str], metadata_filter: Optional[Any], model: str,
            system: str | None = None) -> Any:
        start = time.perf_counter()
        try:
            resp = build_final_response(self.retrieve(contents_to_text(contents), store_names, metadata_filter=metadata_filter))
            gemini_calls_total.labels("generate", "ok").inc()
            return resp
        except Exception:
            gemini_calls_total.labels("generate", "error").inc()
            raise
        finally:
            gemini_latency.labels("generate").observe(time.perf_counter() - start)

    def ask_stream(self, *, contents: Any, store_names: Sequence[str], metadata_filter: Optional[Any], model: str,
                   system: str | None = None) -> Generator:
        start = time.perf_counter()
        try:
            text = contents_to_text(contents)
            grounding = self.retrieve(text, store_names, metadata_filter=metadata_filter)   # the GPU work; no lock is held across the yields
            yield SimpleNamespace(text=stream_lead(grounding), candidates=None,
                                  usage_metadata=SimpleNamespace(prompt_token_count=0, candidates_token_count=0))
            yield build_final_response(grounding)
            gemini_calls_total.labels("generate_stream", "ok").inc()
        except GeneratorExit:            # the route abandoned the stream (client disconnect, chat.py:562-566)
            raise
        except Exception:
            gemini_calls_total.labels("generate_stream", "error").inc()
            raise
        finally:
            gemini_latency.labels("generate_stream").observe(time.perf_counter() - start)

    # -------- Citations (gemini_rag.py:554-599) --------
    @staticmethod
    def extract_citations_from_response(response: Any) -> List[dict[str, Any]]:
        out: List[dict[str, Any]] = []
        try:
            cand = response.candidates[0]
            gm = getattr(cand, "grounding_metadata", None)
            if not gm:
                return out
            for i, ch in enumerate(list(getattr(gm, "grounding_chunks", []) or [])):
                rc = getattr(ch, "retrieved_context", None)
                if rc:
                    out.append({"index": i, "source_type": "retrieved_context", "uri": getattr(rc, "uri", None),
                                "title": getattr(rc, "title", None), "snippet": getattr(rc, "text", None),
                                "store": getattr(rc, "file_search_store", None)})
                    continue
                web = getattr(ch, "web", None)
                if web:
                    out.append({"index": i, "source_type": "web", "uri": getattr(web, "uri", None),
                                "title": getattr(web, "title", None), "snippet": None, "store": None})
            return out
        except (AttributeError, KeyError, IndexError, TypeError) as e:
            # same observable behaviour as the reference (gemini_rag.py:589-595): warn and return what we have
            logging.warning(f"Failed to extract citations: {e}",
                            extra={"response_type": type(response).__name__, "has_candidates": hasattr(response, "candidates")})
            return out

    @staticmethod
    def new_stream_ids() -> tuple[str, str]:
        return str(uuid.uuid4()), str(uuid.uuid4())


# Inside the backend process the reference's own implementations of the pure wire helpers are used
# (nothing to keep in step); the restatements above serve the daemon and the tests, where `app` is absent.
try:
    from app.services.gemini_rag import GeminiRag as _RefGeminiRag   # type: ignore
    B200Rag.extract_citations_from_response = staticmethod(_RefGeminiRag.extract_citations_from_response)
    B200Rag.new_stream_ids = staticmethod(_RefGeminiRag.new_stream_ids)
except Exception:   # noqa: BLE001
    _RefGeminiRag = None


def get_rag_client():
    """What gemini_rag.get_rag_client() (:721-725) returns when RAG_BACKEND=b200:
    the in-process adapter, or -- when RAG_B200_SOCKET names the engine daemon's socket -- the
    client that forwards to it (server.py), so several API / worker processes share one index."""
    if os.environ.get("RAG_B200_SOCKET"):
        from .server import RemoteB200Rag
        return RemoteB200Rag()
    return B200Rag()
