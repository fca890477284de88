"""Process model: one engine daemon owns the HBM index, every API / worker
process talks to it.

The reference runs 4 API workers (backend/Dockerfile:42) plus an ARQ worker process
(backend/app/worker.py:122-126), and builds a new adapter object per request
(services/gemini_rag.py:721-725).  An HBM-resident index lives in ONE process, so:

  * `serve(socket_path)` runs in the process that owns the GPU(s): it wraps a Registry (engine or
    engine group + chunk sidecar) behind a Unix-domain socket, one thread per connection;
  * `RemoteB200Rag` is the adapter the other processes get from `get_rag_client()` when
    RAG_B200_SOCKET is set.  It is the same duck type as B200Rag; every method that touches the
    index is one request/response on the socket, and `ask_stream` builds its chunks locally from
    the returned grounding (so a generator abandoned mid-stream holds nothing on the server).

Transport.  Nothing executable crosses the socket: a frame is a 4-byte big-endian length and a
msgpack map ({"m": method, "a": [...], "k": {...}} -> {"ok": value} | {"err": [type, message]}),
values are plain data (str / int / float / bool / None / bytes / list / map).  A connection starts with
a mutual HMAC-SHA256 challenge on the shared secret RAG_B200_AUTHKEY, which MUST be set (there is no
default key: the daemon refuses to start and the client refuses to connect without it).  The socket is
created with umask 0177 (mode 0600 from the moment it exists).  Only the methods in _ALLOWED are
callable; `save` writes under RAG_B200_SNAPSHOT_DIR only.  The client retries a dropped connection
once, and only for idempotent methods (an upload is never sent twice).
"""
from __future__ import annotations

import hashlib
import hmac
import os
import re
import socket
import struct
import threading
from types import SimpleNamespace
from typing import Any, Dict, Generator, List, Optional, Sequence

import msgpack

from .adapter import B200Rag, Registry, UploadResult, build_final_response, contents_to_text, get_registry, stream_lead

_ALLOWED = {"create_store", "delete_store", "upload_bytes", "op_status", "delete_document_from_store", "retrieve",
            "list_stores", "stats", "save", "export_index", "describe_chunks"}
_IDEMPOTENT = {"op_status", "retrieve", "list_stores", "stats", "delete_store", "delete_document_from_store", "export_index",
               "describe_chunks"}
_MAX_FRAME = 64 << 20          # uploads are capped at 25 MB upstream (config.py:118)
_NONCE = 32
_NAME_RE = re.compile(r"^[A-Za-z0-9._-]{1,64}$")


def _authkey() -> bytes:
    key = os.environ.get("RAG_B200_AUTHKEY", "")
    if len(key) < 16:
        raise RuntimeError("RAG_B200_AUTHKEY must be set to a secret of at least 16 characters "
                           "(the engine daemon and its clients share it; there is no default)")
    return key.encode()


def _recv_exact(sock: socket.socket, n: int) -> bytes:
    buf = bytearray()
    while len(buf) < n:
        part = sock.recv(min(n - len(buf), 1 << 20))
        if not part:
            raise EOFError("connection closed")
        buf += part
    return bytes(buf)


def _send_frame(sock: socket.socket, obj: Any) -> None:
    body = msgpack.packb(obj, use_bin_type=True)
    if len(body) > _MAX_FRAME:
        raise ValueError(f"frame of {len(body)} bytes exceeds the {_MAX_FRAME}-byte limit")
    sock.sendall(struct.pack(">I", len(body)) + body)


def _recv_frame(sock: socket.socket) -> Any:
    (n,) = struct.unpack(">I", _recv_exact(sock, 4))
    if n > _MAX_FRAME:
        raise ValueError(f"frame of {n} bytes exceeds the {_MAX_FRAME}-byte limit")
    return msgpack.unpackb(_recv_exact(sock, n), raw=False, strict_map_key=False)


def _mac(key: bytes, role: bytes, nonce: bytes) -> bytes:
    return hmac.new(key, role + nonce, hashlib.sha256).digest()


def _handshake_server(sock: socket.socket, key: bytes) -> None:
    nonce = os.urandom(_NONCE)
    sock.sendall(nonce)
    reply = _recv_exact(sock, 32 + _NONCE)
    if not hmac.compare_digest(reply[:32], _mac(key, b"client", nonce)):
        raise PermissionError("client failed the authentication challenge")
    sock.sendall(_mac(key, b"server", reply[32:]))


def _handshake_client(sock: socket.socket, key: bytes) -> None:
    nonce = _recv_exact(sock, _NONCE)
    mine = os.urandom(_NONCE)
    sock.sendall(_mac(key, b"client", nonce) + mine)
    if not hmac.compare_digest(_recv_exact(sock, 32), _mac(key, b"server", mine)):
        raise PermissionError("the daemon failed the authentication challenge")


class _Service:
    """Server-side dispatch target: a B200Rag plus the calls that only make sense remotely."""

    def __init__(self, registry: Registry, snapshot_dir: Optional[str] = None):
        self.registry = registry
        self.rag = B200Rag(registry=registry)
        self.snapshot_dir = snapshot_dir or os.environ.get("RAG_B200_SNAPSHOT_DIR")

    def upload_bytes(self, store_name: str, data: bytes, display_name: str, custom_metadata=None) -> dict:
        up = self.rag.upload_bytes(store_name, bytes(data), display_name=display_name, custom_metadata=custom_metadata)
        return {"operation_name": up.operation_name, "file_id": up.file_id}

    def stats(self) -> dict:
        return self.registry.engine.stats()

    def export_index(self) -> bytes:
        """CUDA IPC description of the arena (rf_engine_export): a client process on the same GPU maps it read-only and
        launches its own searches (RemoteB200Rag in attach mode).  One engine only: a group's devices are searched here."""
        eng = self.registry.engine
        if not hasattr(eng, "export_state"):
            raise NotImplementedError("the daemon serves an engine group: attach mode needs a single engine")
        return eng.export_state()

    def describe_chunks(self, ids, scores, cos, store_names=None) -> list:
        return self.rag.describe_chunks([int(x) for x in ids], [int(x) for x in scores], [float(x) for x in cos],
                                        None if store_names is None else [str(s) for s in store_names])

    def save(self, name: str = "snapshot") -> str:
        """Snapshot into <RAG_B200_SNAPSHOT_DIR>/<name>; the client chooses a plain name, never a path."""
        if not self.snapshot_dir:
            raise ValueError("the daemon was started without RAG_B200_SNAPSHOT_DIR: snapshots are disabled")
        if not isinstance(name, str) or not _NAME_RE.match(name) or name in (".", ".."):
            raise ValueError("snapshot name must match [A-Za-z0-9._-]{1,64}")
        target = os.path.join(self.snapshot_dir, name)
        self.registry.save(target)
        return target

    def __getattr__(self, name):   # everything else is the adapter's own method
        return getattr(self.rag, name)


def _handle(conn: socket.socket, service: _Service, key: bytes) -> None:
    try:
        conn.settimeout(10.0)
        _handshake_server(conn, key)
        conn.settimeout(None)
        while True:
            try:
                req = _recv_frame(conn)
            except (EOFError, OSError):
                return
            try:
                method = req.get("m") if isinstance(req, dict) else None
                if method not in _ALLOWED:
                    raise AttributeError(f"method {method!r} is not served")
                reply = {"ok": getattr(service, method)(*(req.get("a") or []), **(req.get("k") or {}))}
            except Exception as exc:   # noqa: BLE001  (shipped to the caller, re-raised there)
                reply = {"err": [type(exc).__name__, str(exc)]}
            _send_frame(conn, reply)
    except Exception:   # noqa: BLE001  failed handshake, malformed frame, client gone: drop the connection
        pass
    finally:
        try:
            conn.close()
        except OSError:
            pass


class Server:
    def __init__(self, socket_path: str, registry: Optional[Registry] = None, snapshot_dir: Optional[str] = None):
        self._key = _authkey()                                 # refuse to start without a secret
        self.socket_path = socket_path
        self.service = _Service(registry or get_registry(), snapshot_dir)
        if os.path.exists(socket_path):
            os.unlink(socket_path)
        self._sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        old = os.umask(0o177)                                  # the socket is 0600 from the moment bind() creates it
        try:
            self._sock.bind(socket_path)
        finally:
            os.umask(old)
        self._sock.listen(64)
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None

    def serve_forever(self) -> None:
        while not self._stop.is_set():
            try:
                conn, _ = self._sock.accept()
            except OSError:
                break
            threading.Thread(target=_handle, args=(conn, self.service, self._key), daemon=True).start()

    def start(self) -> "Server":
        self._thread = threading.Thread(target=self.serve_forever, daemon=True)
        self._thread.start()
        return self

    def close(self) -> None:
        self._stop.set()
        try:
            self._sock.shutdown(socket.SHUT_RDWR)              # unblocks accept()
        except OSError:
            pass
        self._sock.close()
        if os.path.exists(self.socket_path):
            os.unlink(self.socket_path)


def serve(socket_path: Optional[str] = None) -> None:
    """Entry point of the daemon: `python -m rag_foundation_b200.server` (RAG_B200_SOCKET, RAG_B200_AUTHKEY,
    RAG_B200_DEVICES / RAG_B200_DEVICE, RAG_B200_CAPACITY_ROWS, RAG_B200_SNAPSHOT_DIR)."""
    path = socket_path or os.environ.get("RAG_B200_SOCKET")
    if not path:
        raise RuntimeError("RAG_B200_SOCKET must name the daemon's Unix socket (put it in a directory only the service user can enter)")
    Server(path).serve_forever()


_EXC = {"TimeoutError": TimeoutError, "ValueError": ValueError, "FileNotFoundError": FileNotFoundError,
        "NotImplementedError": NotImplementedError, "KeyError": KeyError, "AttributeError": AttributeError,
        "PermissionError": PermissionError}


# attach mode: one read-only engine per (process, daemon socket), shared by the per-request adapter objects
_attached_lock = threading.Lock()
_attached: Dict[str, Dict[str, Any]] = {}


class RemoteB200Rag:
    """Client-side adapter: same protocol as B200Rag / GeminiRag, index behind the socket.

    Attach mode (`attach=True` or RAG_B200_ATTACH=1; the daemon and this process on the same GPU): the process maps
    the daemon's arena read-only over CUDA IPC (Engine.attach) and runs the search itself, on its own streams --
    the four API workers of the reference deployment (backend/Dockerfile:42) then scan concurrently instead of
    queueing on the daemon; only the small "which document is chunk N" question still goes over the socket.  The
    mapping is refreshed from the daemon every RAG_B200_ATTACH_REFRESH_MS (default 200) and whenever a store name
    is unknown; deletes need no refresh (they mask rows in the shared arena)."""

    def __init__(self, socket_path: Optional[str] = None, top_k: int = 10, attach: Optional[bool] = None, scoring: Optional[str] = None):
        self.socket_path = socket_path or os.environ["RAG_B200_SOCKET"]
        self._key = _authkey()
        self.is_mock = True
        self.is_b200 = True
        self.top_k = top_k
        self.attach = (os.environ.get("RAG_B200_ATTACH", "0") == "1") if attach is None else bool(attach)
        self.scoring = (scoring or os.environ.get("RAG_B200_SCORING", "tf")).lower()
        self._refresh_s = float(os.environ.get("RAG_B200_ATTACH_REFRESH_MS", "200")) / 1e3
        self._local = threading.local()

    def _attached_engine(self, force_refresh: bool = False):
        import time
        from .engine import Engine
        with _attached_lock:
            slot = _attached.get(self.socket_path)
            now = time.monotonic()
            if slot is None:
                slot = {"engine": Engine.attach(bytes(self._call("export_index"))), "at": now}
                _attached[self.socket_path] = slot
            elif force_refresh or now - slot["at"] > self._refresh_s:
                slot["engine"].refresh(bytes(self._call("export_index")))
                slot["at"] = now
            return slot["engine"]

    def _retrieve_attached(self, text: str, store_names: Sequence[str], k: int) -> List[dict]:
        eng = self._attached_engine()
        names = list(dict.fromkeys(store_names))
        segs = [eng.lookup_store(s) for s in names]
        if any(x is None for x in segs):                    # a store made after the last refresh?
            eng = self._attached_engine(force_refresh=True)
            segs = [eng.lookup_store(s) for s in names]
        segs = [x for x in segs if x is not None]
        if not segs:
            return []
        weights = eng.scope_weights(segs) if self.scoring == "idf" else None
        ids, scores, cos, _q = eng.search_text(text.encode("utf-8"), segs, k, weights=weights)
        if len(ids) == 0:
            return []
        return self._call("describe_chunks", ids.tolist(), scores.tolist(), cos.tolist(), names)

    def _connect(self) -> socket.socket:
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        try:
            s.settimeout(10.0)
            s.connect(self.socket_path)
            _handshake_client(s, self._key)
            s.settimeout(None)
        except Exception:
            s.close()
            raise
        return s

    def _call(self, method: str, *args, **kwargs):
        conn = getattr(self._local, "conn", None)
        attempts = (0, 1) if method in _IDEMPOTENT else (1,)
        reply = None
        for attempt in attempts:
            sent = False
            try:
                if conn is None:
                    conn = self._connect()
                    self._local.conn = conn
                sent = True
                _send_frame(conn, {"m": method, "a": list(args), "k": kwargs})
                reply = _recv_frame(conn)
                break
            except PermissionError:
                self._local.conn = None
                raise
            except (EOFError, OSError, ConnectionError) as exc:
                if conn is not None:
                    try:
                        conn.close()
                    except OSError:
                        pass
                self._local.conn = conn = None
                if attempt or (sent and method not in _IDEMPOTENT):
                    # the daemon is away: retryable for the chat route (gemini_rag.py:22-27)
                    raise TimeoutError(f"rag-b200 daemon unreachable at {self.socket_path}") from exc
        if "ok" in reply:
            return reply["ok"]
        name, msg = reply["err"]
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
        if self.attach and not metadata_filter:             # (a metadata filter needs the daemon's document table)
            return self._retrieve_attached(text, store_names, k or self.top_k)
        return self._call("retrieve", text, list(store_names), k or self.top_k, metadata_filter)

    def ask(self, *, contents: Any, store_names: Sequence[str], metadata_filter: Optional[Any], model: str,
            system: str | None = None) -> Any:
        return build_final_response(self.retrieve(contents_to_text(contents), store_names, metadata_filter=metadata_filter))

    def ask_stream(self, *, contents: Any, store_names: Sequence[str], metadata_filter: Optional[Any], model: str,
                   system: str | None = None) -> Generator:
        grounding = self.retrieve(contents_to_text(contents), store_names, metadata_filter=metadata_filter)
        yield SimpleNamespace(text=stream_lead(grounding), candidates=None,
                              usage_metadata=SimpleNamespace(prompt_token_count=0, candidates_token_count=0))
        yield build_final_response(grounding)

    extract_citations_from_response = staticmethod(B200Rag.extract_citations_from_response)
    new_stream_ids = staticmethod(B200Rag.new_stream_ids)


if __name__ == "__main__":
    serve()
