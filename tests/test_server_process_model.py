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


@pytest.fixture()
def served(tmp_path):
    reg = ad.Registry(ScriptedEngine())
    sock = str(tmp_path / "rag.sock")
    srv = Server(sock, reg).start()
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


def test_daemon_away_is_retryable(tmp_path):
    with pytest.raises(TimeoutError):                          # gemini_rag.py:22-27: TimeoutError is retried upstream
        RemoteB200Rag(str(tmp_path / "nobody.sock")).create_store("x")
