"""BASELINE.json configs[0] through the real engine: ingest docs/demo/sample-report.md (text embedded
in tests/golden/rf1_golden.json) + the README demo query, citations through the B200Rag adapter;
wire shapes against the reference's MockGeminiRag golden, ranking against the oracle."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_config1_demo_flow(tmp_path, golden_dir):
    from oracle import c_oracle as co
    from rag_foundation_b200 import Engine
    from rag_foundation_b200 import adapter as ad

    wire = json.load(open(os.path.join(golden_dir, "config1_wire.json")))
    g = json.load(open(os.path.join(golden_dir, "rf1_golden.json")))
    reg = ad.Registry(Engine(capacity_rows=4096))
    try:
        rag = ad.B200Rag(registry=reg)
        store = rag.create_store("demo")
        other = rag.create_store("someone-else")
        p = tmp_path / "sample-report.md"
        p.write_bytes(g["sample_report"]["text"].encode("utf-8"))
        long_p = tmp_path / "long.txt"
        long_p.write_bytes(g["long_doc"]["text"].encode("utf-8"))
        up = rag.upload_file(store, str(p), display_name="sample-report.md")
        rag.upload_file(store, str(long_p), display_name="long.txt")
        rag.upload_file(other, str(p), display_name="not-yours.md")          # another tenant's copy
        assert rag.op_status(up.operation_name) == {"name": up.operation_name, "done": True,
                                                    "metadata": {"n_chunks": 1, "file_id": up.file_id}, "error": None}
        contents = [{"role": "user", "parts": [{"text": wire["demo_query"]}]}]
        chunks = list(rag.ask_stream(contents=contents, store_names=[store], metadata_filter=None,
                                     model="gemini-2.5-flash", system=None))
        assert len(chunks) == wire["n_stream_chunks"] and chunks[0].candidates is None and chunks[1].text is None
        cits = rag.extract_citations_from_response(chunks[1])
        assert sorted(cits[0]) == wire["citation_keys"]
        assert all(c["store"] == store for c in cits)                         # tenant mask held
        assert [c["index"] for c in cits] == list(range(len(cits)))

        # ranking == oracle over the same rows
        n_rows = 1 + g["long_doc"]["n_chunks"] + 1
        F, seg, ff = reg.engine.read_rows(0, n_rows)
        q = co.query_vector(wire["demo_query"].encode())
        w_ids, w_sc, _ = co.score_topk(F, seg, q, [reg.engine.lookup_store(store)], ff=ff)
        got = [int(c["uri"].rsplit("#", 1)[1]) for c in cits]
        assert got == w_ids.tolist()
        assert cits[0]["title"] == "sample-report.md" and cits[0]["snippet"].startswith("Demo Source Document")
        assert chunks[0].text.startswith("[b200-retrieval] Demo Source Document")

        # delete the demo document: its chunk disappears from the answer
        rag.delete_document_from_store(store, 1, file_id=up.file_id)
        cits2 = rag.extract_citations_from_response(rag.ask(contents=contents, store_names=[store],
                                                            metadata_filter=None, model="m"))
        assert all(c["title"] == "long.txt" for c in cits2)
        # the other tenant still sees theirs
        cits3 = rag.extract_citations_from_response(rag.ask(contents=contents, store_names=[other],
                                                            metadata_filter=None, model="m"))
        assert [c["title"] for c in cits3] == ["not-yours.md"]
        rag.delete_store(other)
        assert rag.retrieve(wire["demo_query"], [other]) == []
    finally:
        reg.engine.close()


def test_snapshot_round_trip(tmp_path, golden_dir):
    """save the HBM index + sidecar, load into a fresh engine, same answers (including
    a tombstoned document and a dropped store)."""
    from rag_foundation_b200 import Engine
    from rag_foundation_b200 import adapter as ad

    g = json.load(open(os.path.join(golden_dir, "rf1_golden.json")))
    wire = json.load(open(os.path.join(golden_dir, "config1_wire.json")))
    reg = ad.Registry(Engine(capacity_rows=4096))
    try:
        rag = ad.B200Rag(registry=reg)
        a, b, c = rag.create_store("a"), rag.create_store("b"), rag.create_store("c")
        files = {}
        for name, key in (("sample.md", "sample_report"), ("long.txt", "long_doc")):
            p = tmp_path / name
            p.write_bytes(g[key]["text"].encode("utf-8"))
            files[name] = p
        up1 = rag.upload_file(a, str(files["sample.md"]), display_name="sample.md", custom_metadata={"team": "ops"})
        rag.upload_file(a, str(files["long.txt"]), display_name="long.txt")
        up3 = rag.upload_file(b, str(files["long.txt"]), display_name="b-long.txt")
        rag.upload_file(c, str(files["sample.md"]), display_name="c-sample.md")
        rag.delete_document_from_store(b, 3, file_id=up3.file_id)
        rag.delete_store(c)
        before = {s: rag.retrieve(wire["demo_query"], [s]) for s in (a, b, c)}
        before_f = rag.retrieve(wire["demo_query"], [a], metadata_filter={"team": "ops"})
        rows_before = reg.engine.read_rows(0, reg.engine.stats()["n_rows"])
        reg.save(str(tmp_path / "snap"))
    finally:
        reg.engine.close()

    reg2 = ad.Registry.load(Engine(capacity_rows=8192), str(tmp_path / "snap"))
    try:
        rag2 = ad.B200Rag(registry=reg2)
        rows_after = reg2.engine.read_rows(0, reg2.engine.stats()["n_rows"])
        for x, y in zip(rows_before, rows_after):
            assert (x == y).all()
        for s in (a, b, c):
            assert rag2.retrieve(wire["demo_query"], [s]) == before[s]
        assert rag2.retrieve(wire["demo_query"], [a], metadata_filter={"team": "ops"}) == before_f
        assert before[b] == [] and before[c] == [] and len(before[a]) == 1 + g["long_doc"]["n_chunks"] and len(before_f) == 1
        # the restored engine keeps working: new uploads append after the restored rows
        p = tmp_path / "new.txt"; p.write_text("what does this demo prove about engineering flow")
        rag2.upload_file(a, str(p), display_name="new.txt")
        assert rag2.retrieve(wire["demo_query"], [a])[0]["title"] in ("new.txt", "sample.md")
        with pytest.raises(RuntimeError):
            reg2.engine.load_snapshot(str(tmp_path / "snap" / "index.rfsnap"))   # only into an empty engine
    finally:
        reg2.engine.close()
    with pytest.raises(RuntimeError):
        e = Engine(capacity_rows=4)
        try:
            e.load_snapshot(str(tmp_path / "snap" / "index.rfsnap"))             # capacity too small
        finally:
            e.close()


def _run_client(code, env, timeout=120):
    import subprocess
    import sys
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.timeout(300)
def test_daemon_process_ingest_process_query_process(tmp_path, golden_dir):
    """SURVEY 8f-1 on the GPU with REAL processes, the reference's layout (backend/Dockerfile:42: 4 API workers;
    worker.py:122-126: one ARQ ingest worker): the engine daemon is its own process and owns the HBM index, a
    second process (the 'worker') creates a store and uploads, a third (an 'API worker') asks -- all through
    get_rag_client() with RAG_B200_SOCKET set -- and this process hammers it from 8 threads meanwhile.  The
    daemon then snapshots, is killed, and a fresh daemon serves the same answers from the snapshot."""
    import signal
    import subprocess
    import sys
    import textwrap
    import threading
    import time
    from rag_foundation_b200.server import RemoteB200Rag

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = json.load(open(os.path.join(golden_dir, "rf1_golden.json")))
    wire = json.load(open(os.path.join(golden_dir, "config1_wire.json")))
    sock = str(tmp_path / "rag.sock")
    env = dict(os.environ, RAG_B200_SOCKET=sock, RAG_B200_AUTHKEY="cross-process-secret-0123456789", RAG_B200_CAPACITY_ROWS="8192",
               RAG_B200_SNAPSHOT_DIR=str(tmp_path / "snaps"), PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    doc = tmp_path / "sample-report.md"
    doc.write_bytes(g["sample_report"]["text"].encode("utf-8"))

    def start_daemon(extra=""):
        code = f"import sys; sys.path.insert(0, {root!r})\n{extra}\nfrom rag_foundation_b200.server import serve; serve()"
        p = subprocess.Popen([sys.executable, "-c", code], env=env, stderr=subprocess.PIPE, text=True)
        for _ in range(600):
            if os.path.exists(sock):
                return p
            if p.poll() is not None:
                raise AssertionError("daemon died: " + p.stderr.read()[-2000:])
            time.sleep(0.1)
        p.kill()
        raise AssertionError("daemon did not come up")

    daemon = start_daemon()
    try:
        worker = textwrap.dedent(f"""
            import json
            from rag_foundation_b200 import get_rag_client
            rag = get_rag_client()
            store = rag.create_store("demo")
            up = rag.upload_file(store, {str(doc)!r}, display_name="sample-report.md")
            print(json.dumps({{"cls": type(rag).__name__, "store": store, "done": rag.op_status(up.operation_name)["done"], "file": up.file_id}}))
        """)
        w = _run_client(worker, env)
        assert w["cls"] == "RemoteB200Rag" and w["done"] is True
        api = textwrap.dedent(f"""
            import json
            from rag_foundation_b200 import get_rag_client
            rag = get_rag_client()
            chunks = list(rag.ask_stream(contents=[{{"role": "user", "parts": [{{"text": {wire["demo_query"]!r}}}]}}],
                                         store_names=[{w["store"]!r}], metadata_filter=None, model="m"))
            cits = rag.extract_citations_from_response(chunks[1])
            print(json.dumps({{"n": len(chunks), "cits": cits}}))
        """)
        a = _run_client(api, env)
        assert a["n"] == 2 and a["cits"] and a["cits"][0]["title"] == "sample-report.md" and a["cits"][0]["store"] == w["store"]
        os.environ["RAG_B200_AUTHKEY"] = env["RAG_B200_AUTHKEY"]
        try:
            local = RemoteB200Rag(sock).retrieve(wire["demo_query"], [w["store"]])
            assert [c["uri"] for c in a["cits"]] == [h["uri"] for h in local]
            errs = []

            def client():
                try:
                    r = RemoteB200Rag(sock)
                    for _ in range(20):
                        assert r.retrieve(wire["demo_query"], [w["store"]]) == local
                except Exception as ex:   # noqa: BLE001
                    errs.append(ex)
            ts = [threading.Thread(target=client) for _ in range(8)]
            [t.start() for t in ts]; [t.join() for t in ts]
            assert not errs, errs[:1]
            stats = RemoteB200Rag(sock)._call("stats")
            assert stats["n_rows"] >= 1 and stats["searches"] >= 160
            where = RemoteB200Rag(sock)._call("save", "nightly")
            daemon.send_signal(signal.SIGKILL)
            daemon.wait(30)
            os.unlink(sock)
            # a fresh daemon process restores the snapshot before it starts serving
            daemon = start_daemon(extra=("from rag_foundation_b200 import adapter as ad, Engine\n"
                                         f"ad.set_registry(ad.Registry.load(Engine(capacity_rows=8192), {where!r}))"))
            assert RemoteB200Rag(sock).retrieve(wire["demo_query"], [w["store"]]) == local
        finally:
            os.environ.pop("RAG_B200_AUTHKEY", None)
    finally:
        if daemon.poll() is None:
            daemon.kill()
            daemon.wait(30)


@pytest.mark.timeout(300)
def test_api_process_attaches_to_the_daemons_arena_over_cuda_ipc(tmp_path, golden_dir):
    """SURVEY 8f-1, second half: an API process maps the daemon's HBM arena read-only (rf_engine_export ->
    rf_engine_attach over CUDA IPC) and runs the searches ITSELF; only the chunk -> document question goes over the
    socket.  Its answers equal the daemon's own; a store created later is found after a refresh; a delete by the
    daemon is visible at once; the attachment refuses to change the index."""
    import subprocess
    import sys
    import textwrap
    import time

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = json.load(open(os.path.join(golden_dir, "rf1_golden.json")))
    wire = json.load(open(os.path.join(golden_dir, "config1_wire.json")))
    sock = str(tmp_path / "rag.sock")
    env = dict(os.environ, RAG_B200_SOCKET=sock, RAG_B200_AUTHKEY="cuda-ipc-attach-secret-0123456789", RAG_B200_CAPACITY_ROWS="8192",
               PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    doc = tmp_path / "sample-report.md"
    doc.write_bytes(g["sample_report"]["text"].encode("utf-8"))
    doc2 = tmp_path / "long.txt"
    doc2.write_bytes(g["long_doc"]["text"].encode("utf-8"))
    code = f"import sys; sys.path.insert(0, {root!r})\nfrom rag_foundation_b200.server import serve; serve()"
    daemon = subprocess.Popen([sys.executable, "-c", code], env=env, stderr=subprocess.PIPE, text=True)
    try:
        for _ in range(600):
            if os.path.exists(sock):
                break
            assert daemon.poll() is None, "daemon died: " + daemon.stderr.read()[-2000:]
            time.sleep(0.1)
        api = textwrap.dedent(f"""
            import json, os
            from rag_foundation_b200.server import RemoteB200Rag
            rpc = RemoteB200Rag(attach=False)
            att = RemoteB200Rag(attach=True)
            store = rpc.create_store("demo")
            rpc.upload_file(store, {str(doc)!r}, display_name="sample-report.md")
            q = {wire["demo_query"]!r}
            out = {{}}
            out["same_1"] = att.retrieve(q, [store]) == rpc.retrieve(q, [store]) and len(att.retrieve(q, [store])) >= 1
            eng = att._attached_engine()
            out["attached"] = bool(getattr(eng, "attached", False))
            searches_before = rpc._call("stats")["searches"]
            for _ in range(5):
                att.retrieve(q, [store])
            out["daemon_searches_during_attached_queries"] = rpc._call("stats")["searches"] - searches_before
            # a second store and document appear AFTER the attachment was made: found through a refresh
            store2 = rpc.create_store("later")
            up = rpc.upload_file(store2, {str(doc2)!r}, display_name="long.txt")
            out["same_2"] = att.retrieve("w5 w17 w30", [store2]) == rpc.retrieve("w5 w17 w30", [store2]) and len(rpc.retrieve("w5 w17 w30", [store2])) >= 1
            out["same_both"] = att.retrieve("w5 demo flow", [store, store2]) == rpc.retrieve("w5 demo flow", [store, store2])
            # idf scoring runs its statistics kernel in this process too
            out["same_idf"] = RemoteB200Rag(attach=True, scoring="idf").retrieve("w5 w17 w30", [store2]) != [] 
            # a delete by the daemon masks rows in the shared arena: no refresh needed
            rpc.delete_document_from_store(store2, 0, file_id=up.file_id)
            out["after_delete"] = att.retrieve("w5 w17 w30", [store2])
            # the attachment is read-only
            try:
                eng.ingest_text(eng.lookup_store(store), 99, b"not allowed here")
                out["read_only"] = False
            except RuntimeError as exc:
                out["read_only"] = "read-only" in str(exc)
            print(json.dumps(out))
        """)
        r = _run_client(api, env)
        assert r["same_1"] and r["attached"] and r["same_2"] and r["same_both"] and r["same_idf"], r
        assert r["daemon_searches_during_attached_queries"] == 0, r       # the scans ran in the client process
        assert r["after_delete"] == [] and r["read_only"] is True, r
    finally:
        if daemon.poll() is None:
            daemon.kill()
            daemon.wait(30)


def test_idf_scoring_through_the_adapter_and_quality_eval(tmp_path):
    """RAG_B200_SCORING=idf (RF-1w): citations rank as the oracle's weighted ranking says, the plain
    adapter on the same registry is unchanged, and the reference-style citation grading runs end to end."""
    import importlib.util
    from oracle import c_oracle as co
    from rag_foundation_b200 import Engine
    from rag_foundation_b200 import adapter as ad
    spec = importlib.util.spec_from_file_location("quality_eval", os.path.join(os.path.dirname(__file__), "..", "tools", "quality_eval.py"))
    qe = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(qe)

    docs, questions = qe.make_labelled_set(150, 60, seed=4)
    reg = ad.Registry(Engine(capacity_rows=8192))
    try:
        res = qe.run_eval(lambda scoring: ad.B200Rag(registry=reg, scoring=scoring), docs, questions)
        # an independent numpy model of RF-1 / RF-1w over the same labelled set grades identically
        model = qe.model_dim_sweep(docs, questions, dims=(256,))["256"]
        for mode in ("tf", "idf"):
            assert res["scoring"][mode]["citation_hit_rate"] == pytest.approx(model[mode], abs=1e-12), (res, model)
        assert res["scoring"]["idf"]["citation_hit_rate"] > res["scoring"]["tf"]["citation_hit_rate"] > 0
        rag_idf, rag_tf = ad.B200Rag(registry=reg, scoring="idf"), ad.B200Rag(registry=reg, scoring="tf")
        store = next(iter({d.store_name for d in reg.docs.values()}))
        seg = reg.engine.lookup_store(store)
        n_rows = reg.engine.stats()["n_rows"]
        F, sg, ff = reg.engine.read_rows(0, n_rows)
        w = co.idf_weights(*co.bucket_df(F, sg, [seg]))
        assert (reg.engine.scope_weights([seg]) == w).all()
        for rec in questions[:12]:
            q = co.query_vector(rec["question"].encode())
            for rag, qv in ((rag_tf, q), (rag_idf, co.weight_query(q, w))):
                got = rag.retrieve(rec["question"], [store])
                w_ids, w_sc, _ = co.score_topk(F, sg, qv, [seg], ff=ff)
                assert [g["chunk_id"] for g in got] == w_ids.tolist() and [g["score"] for g in got] == w_sc.tolist()
        with pytest.raises(ValueError):
            ad.B200Rag(registry=reg, scoring="bm25")
    finally:
        reg.engine.close()
