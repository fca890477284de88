"""Process model (rag_foundation_b200/server.py): the engine daemon behind a Unix socket and the
client adapter other processes use.  CPU: the engine is the scripted double of test_adapter_host;
a second PROCESS drives the client to prove the index is shared across processes."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from rag_foundation_b200 import adapter as ad
from rag_foundation_b200.server import RemoteB200Rag, Server
from test_adapter_host import ScriptedEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.timeout(120)


KEY = "test-secret-0123456789abcdef"


@pytest.fixture()
def served(tmp_path, monkeypatch):
    monkeypatch.setenv("RAG_B200_AUTHKEY", KEY)
    reg = ad.Registry(ScriptedEngine())
    sock = str(tmp_path / "rag.sock")
    srv = Server(sock, reg, snapshot_dir=str(tmp_path / "snaps")).start()
    yield sock, reg
    srv.close()


def test_remote_adapter_same_protocol(served, tmp_path):
    sock, reg = served
    rag = RemoteB200Rag(sock)
    store = rag.create_store("demo")
    assert store.startswith("fileSearchStores/") and rag.is_mock
    p = tmp_path / "doc.txt"
    p.write_text("shared index across processes " * 20)
    up = rag.upload_file(store, str(p), display_name="doc.txt", custom_metadata={"team": "ops"})
    assert up.operation_name.startswith("operations/") and rag.op_status(up.operation_name)["done"] is True
    reg.engine.hits = [(0, 9)]
    chunks = list(rag.ask_stream(contents=[{"role": "user", "parts": [{"text": "shared index"}]}], store_names=[store],
                                 metadata_filter=None, model="m"))
    assert len(chunks) == 2 and chunks[0].candidates is None and chunks[1].text is None
    cits = rag.extract_citations_from_response(chunks[1])
    assert cits[0]["title"] == "doc.txt" and cits[0]["store"] == store and cits[0]["snippet"].startswith("shared index")
    assert rag.retrieve("q", [store], metadata_filter={"team": "dev"}) == []
    with pytest.raises(ValueError):
        rag.op_status({})                                    # server-side exception types survive the socket
    with pytest.raises(AttributeError):
        rag._call("engine")                                  # not an allowed method
    rag.delete_document_from_store(store, 1, file_id=up.file_id)
    assert reg.engine.tombstoned == [1]
    a, b = rag.new_stream_ids()
    assert len(a) == 36 and a != b


def test_second_process_sees_the_same_index(served, tmp_path):
    sock, reg = served
    store = RemoteB200Rag(sock).create_store("demo")
    p = tmp_path / "doc.txt"
    p.write_text("written by the ingest worker process")
    code = textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        os.environ["RAG_B200_SOCKET"] = {sock!r}
        os.environ["RAG_B200_AUTHKEY"] = {KEY!r}
        from rag_foundation_b200 import get_rag_client
        rag = get_rag_client()
        up = rag.upload_file({store!r}, {str(p)!r}, display_name="worker.txt")
        print(json.dumps({{"cls": type(rag).__name__, "op": rag.op_status(up.operation_name)["done"], "file": up.file_id}}))
    """)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["cls"] == "RemoteB200Rag" and r["op"] is True
    assert reg.engine.ingested and list(reg.docs.values())[0].display_name == "worker.txt"   # landed in THIS process's registry
    reg.engine.hits = [(0, 3)]
    cits = RemoteB200Rag(sock).retrieve("q", [store])
    assert cits[0]["title"] == "worker.txt"


def test_daemon_away_is_retryable(tmp_path, monkeypatch):
    monkeypatch.setenv("RAG_B200_AUTHKEY", KEY)
    with pytest.raises(TimeoutError):                          # gemini_rag.py:22-27: TimeoutError is retried upstream
        RemoteB200Rag(str(tmp_path / "nobody.sock")).create_store("x")


def test_no_default_secret_and_wrong_secret_is_refused(served, tmp_path, monkeypatch):
    sock, reg = served
    monkeypatch.delenv("RAG_B200_AUTHKEY")
    with pytest.raises(RuntimeError, match="RAG_B200_AUTHKEY"):
        Server(str(tmp_path / "other.sock"), reg)               # the daemon refuses to start without a secret
    with pytest.raises(RuntimeError, match="RAG_B200_AUTHKEY"):
        RemoteB200Rag(sock)
    monkeypatch.setenv("RAG_B200_AUTHKEY", "another-secret-0123456789")
    with pytest.raises((PermissionError, TimeoutError)):
        RemoteB200Rag(sock).list_stores()
    assert oct(os.stat(sock).st_mode & 0o777) == "0o600"


def test_wire_carries_data_only_and_save_stays_in_its_directory(served, tmp_path):
    """A pickle sent to the socket is never unpickled (frames are msgpack), and `save` takes a plain name."""
    import pickle
    import socket as so
    import struct
    from rag_foundation_b200 import server as sv
    sock, reg = served
    marker = tmp_path / "pwned"

    class Evil:
        def __reduce__(self):
            return (os.system, (f"touch {marker}",))
    s = so.socket(so.AF_UNIX, so.SOCK_STREAM)
    s.connect(sock)
    sv._handshake_client(s, KEY.encode())
    body = pickle.dumps(("create_store", (Evil(),), {}))
    s.sendall(struct.pack(">I", len(body)) + body)
    try:
        s.settimeout(5)
        s.recv(16)
    except OSError:
        pass
    s.close()
    assert not marker.exists()
    rag = RemoteB200Rag(sock)
    for bad in ("../escape", "/etc/passwd", "a/b", ""):
        with pytest.raises(ValueError):
            rag._call("save", bad)
    reg.engine.save_snapshot = lambda path: open(path, "wb").write(b"snap")     # the scripted engine has no snapshot
    where = rag._call("save", "nightly")
    assert where == str(tmp_path / "snaps" / "nightly") and os.path.exists(os.path.join(where, "sidecar.msgpack"))


def test_upload_is_never_sent_twice(served, tmp_path):
    """A connection that drops after an upload was sent is NOT retried (the document would be ingested twice);
    idempotent calls are."""
    sock, reg = served
    rag = RemoteB200Rag(sock)
    store = rag.create_store("demo")
    p = tmp_path / "doc.txt"
    p.write_text("one copy only")
    rag.list_stores()
    rag._local.conn.close()                                    # the daemon's side sees EOF; our socket object is dead
    with pytest.raises(TimeoutError):
        rag.upload_file(store, str(p))
    assert reg.engine.ingested == []
    rag._local.conn = None
    rag.upload_file(store, str(p))
    rag._local.conn.close()
    assert rag.list_stores() == []                             # idempotent: reconnects and succeeds
    assert len(reg.engine.ingested) == 1
