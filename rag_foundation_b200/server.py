"""Process model: one engine daemon owns the HBM index, every API / worker
process talks to it.

The reference runs 4 API workers (backend/Dockerfile:42) plus an ARQ worker process
(backend/app/worker.py:122-126), and builds a new adapter object per request
(services/gemini_rag.py:721-725).  An HBM-resident index lives in ONE process, so:

  * `serve(socket_path)` runs in the process that owns the GPU: it wraps a Registry (engine +
    chunk sidecar) behind a Unix-domain socket, one thread per connection;
  * `RemoteB200Rag` is the adapter the other processes get from `get_rag_client()` when
    RAG_B200_SOCKET is set.  It is the same duck type as B200Rag; every method that touches the
    index is one request/response on the socket, and `ask_stream` builds its chunks locally from
    the returned grounding (so a generator abandoned mid-stream holds nothing on the server).

Transport: multiprocessing.connection (length-prefixed pickles, HMAC challenge on connect with
RAG_B200_AUTHKEY).  Only the methods in _ALLOWED are callable.
"""
from __future__ import annotations

import os
import threading
from multiprocessing.connection import Client, Listener
from types import SimpleNamespace
from typing import Any, Dict, Generator, List, Optional, Sequence

from .adapter import B200Rag, Registry, UploadResult, build_final_response, contents_to_text, get_registry

_ALLOWED = {"create_store", "delete_store", "upload_bytes", "op_status", "delete_document_from_store", "retrieve",
            "list_stores", "stats", "save"}


def _authkey() -> bytes:
    return os.environ.get("RAG_B200_AUTHKEY", "rag-b200-local").encode()


class _Service:
    """Server-side dispatch target: a B200Rag plus the two calls that only make sense remotely."""

    def __init__(self, registry: Registry):
        self.registry = registry
        self.rag = B200Rag(registry=registry)

    def upload_bytes(self, store_name: str, data: bytes, display_name: str, custom_metadata=None) -> dict:
        import tempfile
        with tempfile.NamedTemporaryFile(prefix="rfb200-", delete=True) as f:
            f.write(data)
            f.flush()
            up = self.rag.upload_file(store_name, f.name, display_name=display_name, custom_metadata=custom_metadata)
        return {"operation_name": up.operation_name, "file_id": up.file_id}

    def stats(self) -> dict:
        return self.registry.engine.stats()

    def save(self, directory: str) -> None:
        self.registry.save(directory)

    def __getattr__(self, name):   # everything else is the adapter's own method
        return getattr(self.rag, name)


def _handle(conn, service: _Service) -> None:
    try:
        while True:
            try:
                method, args, kwargs = conn.recv()
            except (EOFError, OSError):
                return
            try:
                if method not in _ALLOWED:
                    raise AttributeError(f"method {method!r} is not served")
                conn.send(("ok", getattr(service, method)(*args, **kwargs)))
            except Exception as exc:   # noqa: BLE001  (shipped to the caller, re-raised there)
                conn.send(("err", type(exc).__name__, str(exc)))
    finally:
        conn.close()


class Server:
    def __init__(self, socket_path: str, registry: Optional[Registry] = None):
        self.socket_path = socket_path
        self.service = _Service(registry or get_registry())
        if os.path.exists(socket_path):
            os.unlink(socket_path)
        self.listener = Listener(address=socket_path, family="AF_UNIX", authkey=_authkey())
        os.chmod(socket_path, 0o600)
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None

    def serve_forever(self) -> None:
        while not self._stop.is_set():
            try:
                conn = self.listener.accept()
            except OSError:
                break
            except Exception:   # failed handshake from a stranger: keep serving
                continue
            threading.Thread(target=_handle, args=(conn, self.service), daemon=True).start()

    def start(self) -> "Server":
        self._thread = threading.Thread(target=self.serve_forever, daemon=True)
        self._thread.start()
        return self

    def close(self) -> None:
        self._stop.set()
        try:
            Client(self.socket_path, family="AF_UNIX", authkey=_authkey()).close()   # unblock accept()
        except Exception:
            pass
        self.listener.close()
        if os.path.exists(self.socket_path):
            os.unlink(self.socket_path)


def serve(socket_path: Optional[str] = None) -> None:
    """Entry point of the daemon: `python -m rag_foundation_b200.server` (RAG_B200_SOCKET, RAG_B200_*)."""
    Server(socket_path or os.environ.get("RAG_B200_SOCKET", "/tmp/rag-b200.sock")).serve_forever()


_EXC = {"TimeoutError": TimeoutError, "ValueError": ValueError, "FileNotFoundError": FileNotFoundError,
        "NotImplementedError": NotImplementedError, "KeyError": KeyError, "AttributeError": AttributeError}


class RemoteB200Rag:
    """Client-side adapter: same protocol as B200Rag / GeminiRag, index behind the socket."""

    def __init__(self, socket_path: Optional[str] = None, top_k: int = 10):
        self.socket_path = socket_path or os.environ["RAG_B200_SOCKET"]
        self.is_mock = True
        self.is_b200 = True
        self.top_k = top_k
        self._local = threading.local()

    def _call(self, method: str, *args, **kwargs):
        conn = getattr(self._local, "conn", None)
        for attempt in (0, 1):
            try:
                if conn is None:
                    conn = Client(self.socket_path, family="AF_UNIX", authkey=_authkey())
                    self._local.conn = conn
                conn.send((method, args, kwargs))
                reply = conn.recv()
                break
            except (EOFError, OSError, ConnectionError):
                self._local.conn = conn = None
                if attempt:
                    # the daemon is away: retryable for the chat route (gemini_rag.py:22-27)
                    raise TimeoutError(f"rag-b200 daemon unreachable at {self.socket_path}")
        if reply[0] == "ok":
            return reply[1]
        _, name, msg = reply
        raise _EXC.get(name, RuntimeError)(msg)

    # ---- the adapter protocol ----
    def list_stores(self) -> List[Any]:
        return self._call("list_stores")

    def create_store(self, display_name: str) -> str:
        return self._call("create_store", display_name)

    def delete_store(self, store_name: str) -> None:
        self._call("delete_store", store_name)

    def upload_file(self, store_name: str, file_path: str, *, display_name: Optional[str] = None, custom_metadata=None,
                    chunking_config: Optional[Dict] = None) -> UploadResult:
        with open(file_path, "rb") as f:
            data = f.read()
        r = self._call("upload_bytes", store_name, data, display_name or os.path.basename(file_path), custom_metadata)
        return UploadResult(operation_name=r["operation_name"], file_id=r["file_id"])

    def op_status(self, op_name) -> dict:
        return self._call("op_status", op_name)

    def delete_document_from_store(self, store_name: str, document_id: int, filename: str | None = None,
                                   file_id: str | None = None) -> None:
        self._call("delete_document_from_store", store_name, document_id, filename, file_id)

    def retrieve(self, text: str, store_names: Sequence[str], k: Optional[int] = None, metadata_filter=None) -> List[dict]:
        return self._call("retrieve", text, list(store_names), k or self.top_k, metadata_filter)

    def ask(self, *, contents: Any, store_names: Sequence[str], metadata_filter: Optional[Any], model: str,
            system: str | None = None) -> Any:
        return build_final_response(self.retrieve(contents_to_text(contents), store_names, metadata_filter=metadata_filter))

    def ask_stream(self, *, contents: Any, store_names: Sequence[str], metadata_filter: Optional[Any], model: str,
                   system: str | None = None) -> Generator:
        grounding = self.retrieve(contents_to_text(contents), store_names, metadata_filter=metadata_filter)
        lead = grounding[0]["text"].strip().splitlines()[0][:200] if grounding else "no matching passages"
        yield SimpleNamespace(text=f"[b200-retrieval] {lead}", candidates=None,
                              usage_metadata=SimpleNamespace(prompt_token_count=0, candidates_token_count=0))
        yield build_final_response(grounding)

    extract_citations_from_response = staticmethod(B200Rag.extract_citations_from_response)
    new_stream_ids = staticmethod(B200Rag.new_stream_ids)


if __name__ == "__main__":
    serve()
