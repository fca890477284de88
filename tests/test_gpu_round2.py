"""GPU parity tests added in round 2 (all through the C-ABI): BASELINE configs[2] at its full size, the
fused NVLink exchange under pytest (two engines, peer pointers within one process), the engine group
(several engines behind one index), row reuse after deletes, snapshot integrity, and the overlap promise
of the device-resident searches."""
import os
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

COS_RTOL = 1e-5


@pytest.fixture(scope="module")
def co():
    from oracle import c_oracle
    return c_oracle


@pytest.fixture(scope="module")
def zb():
    from oracle import rf1
    return rf1.zipf_bucket_table()


def _engine(cap, **kw):
    from rag_foundation_b200 import Engine
    return Engine(capacity_rows=cap, **kw)


def _keys(ids, sc):
    """(ids, scores) -> packed RF-1 keys, 0 where there is no result."""
    ids = np.asarray(ids, np.uint64)
    valid = ids != np.uint64(0xFFFFFFFFFFFFFFFF)
    k = (np.asarray(sc).astype(np.int64).astype(np.uint64) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - (ids & np.uint64(0xFFFFFFFF)))
    return np.where(valid, k, np.uint64(0))


# ------------------------------------------------------------------ configs[2] at BASELINE size
@pytest.mark.timeout(600)
def test_config2_full_size_every_query_against_the_oracle(co, zb):
    """BASELINE configs[2] itself: 1 M chunks x 1024 batched queries on the tensor-core path; EVERY query's
    top-10 (ids, scores, tie order) is compared with the C oracle, not a sample."""
    import torch
    n, nq, k = 1_000_000, 1024, 10
    with _engine(n) as e:
        s = e.open_store("fileSearchStores/cfg2")
        e.ingest_synthetic(s, 0, seed=0, start_counter=0, n_rows=n)
        Q = np.stack([co.synth_query(0, i, zb) for i in range(nq)])
        qd = torch.from_numpy(Q).cuda()
        out = torch.zeros((nq, k), dtype=torch.int64, device="cuda")
        l0 = e.stats()["kernel_launches"]
        e.search_keys_device(qd.data_ptr(), nq, [s], k, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert e.stats()["kernel_launches"] - l0 <= 6, "the batch should have taken the tensor-core path"
        keys = out.cpu().numpy().view(np.uint64)
        F = co.synth_rows(0, 0, n, zb)
        seg = np.full(n, s, np.uint32)
        bad = [i for i in range(nq) if keys[i].tolist() != co.score_topk_keys(F, seg, Q[i], [s], k=k).tolist()]
        assert bad == [], f"{len(bad)} of {nq} queries differ from the oracle, first {bad[:5]}"
        # the same batch through the host entry point (rf_search routes it to the same kernels)
        ids, sc, cs, cnt = e.search(Q, (np.full(nq, s, np.uint32), np.arange(nq + 1, dtype=np.uint32)), k=k)
        assert (cnt == k).all() and (_keys(ids, sc) == keys).all()
        # ties at the k-th score: wherever the 10th and 11th best scores are equal the id order decides
        ties = 0
        for i in range(0, nq, 64):
            w = co.score_topk_keys(F, seg, Q[i], [s], k=12)
            ties += int((w[9] >> np.uint64(32)) == (w[10] >> np.uint64(32)))
        assert ties > 0, "the sampled queries should include k-th-score ties (the order under test)"


# ------------------------------------------------------------------ fused exchange under pytest
@pytest.mark.timeout(300)
def test_fused_exchange_two_engines_one_process(co, zb):
    """rf_search_keys_device_fused with world = 2 inside ONE process: two engines (shards) on the same GPU,
    the 'peer' pointers are plain device pointers, the two kernels run concurrently on two streams and each
    finishes by storing into BOTH gather buffers, releasing flags and acquiring the other's.  Result on both
    'ranks' == the NCCL-path merge == the oracle over the whole corpus."""
    import torch
    from rag_foundation_b200.sharded import shard_range
    n, k, nq_cap = 400_000, 10, 8
    world = 2
    engines, streams = [], [torch.cuda.Stream() for _ in range(world)]
    try:
        for r in range(world):
            lo, hi = shard_range(n, r, world)
            e = _engine(hi - lo, id_base=lo)
            engines.append(e)
            s = e.open_store("fileSearchStores/a")
            e.ingest_synthetic(s, 0, seed=5, start_counter=lo, n_rows=hi - lo)
        keys_buf = [torch.zeros(4 * world * nq_cap * k, dtype=torch.int64, device="cuda") for _ in range(world)]
        flag_buf = [torch.zeros(4 * world * nq_cap, dtype=torch.int32, device="cuda") for _ in range(world)]
        timeout = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(world)]
        kp = np.asarray([t.data_ptr() for t in keys_buf], np.uint64)
        fp = np.asarray([t.data_ptr() for t in flag_buf], np.uint64)
        F = co.synth_rows(5, 0, n, zb)
        seg = np.zeros(n, np.uint32)
        # Warm-up (world = 1: no exchange) so every engine's per-stream scratch exists: a device allocation
        # issued between the two launches below would keep them from running concurrently (CUDA's implicit
        # synchronisation rules) -- and here, unlike one process per GPU, each kernel NEEDS the other to run.
        warm = torch.zeros((nq_cap, 256), dtype=torch.int8, device="cuda")
        sink = torch.zeros((nq_cap, k), dtype=torch.int64, device="cuda")
        for r in range(world):
            engines[r].search_keys_device_fused(warm.data_ptr(), nq_cap, [0], k, sink.data_ptr(), streams[r].cuda_stream,
                                                0, 1, nq_cap, 1, kp[r:r + 1], fp[r:r + 1], timeout[r].data_ptr())
        torch.cuda.synchronize()
        for t in flag_buf:
            t.zero_()
        torch.cuda.synchronize()
        seq = 0
        for nq in (1, 3, 1, 8, 1, 1, 2):          # > 4 calls: the four-slot buffers wrap around
            Q = np.stack([co.synth_query(5, 100 * seq + i, zb) for i in range(nq)])
            qd = torch.from_numpy(Q).cuda()
            outs = [torch.zeros((nq, k), dtype=torch.int64, device="cuda") for _ in range(world)]
            torch.cuda.synchronize()
            seq += 1
            for r in range(world):
                engines[r].search_keys_device_fused(qd.data_ptr(), nq, [0], k, outs[r].data_ptr(), streams[r].cuda_stream,
                                                    r, world, nq_cap, seq, kp, fp, timeout[r].data_ptr())
            torch.cuda.synchronize()
            assert not any(int(t.item()) for t in timeout), "a peer's keys never arrived"
            got = [o.cpu().numpy().view(np.uint64) for o in outs]
            assert (got[0] == got[1]).all()
            for i in range(nq):
                assert got[0][i].tolist() == co.score_topk_keys(F, seg, Q[i], [0], k=k).tolist()
    finally:
        for e in engines:
            e.close()


# ------------------------------------------------------------------ engine group
def _group(devices, cap, **kw):
    from rag_foundation_b200 import EngineGroup
    return EngineGroup(devices, capacity_rows=cap, **kw)


def _group_devices(n):
    """n engines: distinct GPUs when the box has them, else all on GPU 0 (the logic under test is the same)."""
    import torch
    have = torch.cuda.device_count()
    return list(range(n)) if have >= n else [0] * n


@pytest.mark.parametrize("n_dev", [2, 8])
def test_group_chunk_sharded_equals_single_engine(co, zb, n_dev):
    """One store spread over n_dev engines (the 100 M-chunk layout in small): ids, scores, cosines and counts
    of rf_group_search == one engine holding everything == the oracle; single queries and batches."""
    n, k = 240_000, 10
    from rag_foundation_b200.sharded import shard_range
    bases = [shard_range(n, d, n_dev)[0] for d in range(n_dev)]
    with _group(_group_devices(n_dev), n // n_dev, placement="spread", id_bases=bases) as g, _engine(n) as e:
        gs = g.open_store("fileSearchStores/big")
        es = e.open_store("fileSearchStores/big")
        g.ingest_synthetic(gs, 0, seed=9, start_counter=0, n_rows=n)
        e.ingest_synthetic(es, 0, seed=9, start_counter=0, n_rows=n)
        tot, per = g.stats(per_device=True)
        assert tot["n_rows"] == n and all(p["n_rows"] == n // n_dev for p in per)
        F, ff = co.synth_rows(9, 0, n, zb, with_ff=True)
        seg = np.zeros(n, np.uint32)
        Q = np.stack([co.synth_query(9, i, zb) for i in range(20)])
        for i in range(4):
            a = g.search(Q[i:i + 1], [[gs]], k=k)
            b = e.search(Q[i:i + 1], [[es]], k=k)
            for x, y in zip(a, b):
                assert (x == y).all()
            w_ids, w_sc, w_cs = co.score_topk(F, seg, Q[i], [0], k=k, ff=ff)
            assert a[0][0].tolist() == w_ids.tolist() and a[1][0].tolist() == w_sc.tolist()
            np.testing.assert_allclose(a[2][0], w_cs, rtol=COS_RTOL)
        a = g.search(Q, [[gs]] * 20, k=k)           # batch: each engine may take the tensor-core route
        b = e.search(Q, [[es]] * 20, k=k)
        for x, y in zip(a, b):
            assert (x == y).all()
        a = g.search(Q[:3], [[gs]] * 3, k=32)        # k = 32 through the host merge
        b = e.search(Q[:3], [[es]] * 3, k=32)
        for x, y in zip(a, b):
            assert (x == y).all()


def test_group_store_placement_multi_tenant_batch(co, zb):
    """Whole stores per engine: 12 stores over 3 engines, a batch of store-scoped queries (each engine gets
    only its own queries), scopes that span engines, an unknown store, a dropped store, a deleted document."""
    rows, n_st, k = 5_000, 12, 10
    with _group(_group_devices(3), 4 * rows + 64, placement="store") as g:
        stores = [g.open_store(f"fileSearchStores/t{i}") for i in range(n_st)]
        assert stores == list(range(n_st))
        g.ingest_synthetic(0, rows, seed=4, start_counter=0, n_rows=rows * n_st)
        tot, per = g.stats(per_device=True)
        assert [p["n_rows"] for p in per] == [4 * rows] * 3
        F = co.synth_rows(4, 0, rows * n_st, zb)
        seg = (np.arange(rows * n_st) // rows).astype(np.uint32)
        # global chunk id of corpus row r: engine d = store % 3 numbers its rows from d * stride
        stride = 0xFFFFFFFE // 3
        store_of = np.arange(rows * n_st) // rows
        gid = (store_of % 3) * stride + (store_of // 3) * rows + np.arange(rows * n_st) % rows

        def want(q, scope):
            # independent numpy restatement of RF-1 steps 6-7 over the group's chunk ids (a scope that spans
            # engines ranks score ties by those ids, not by the corpus row)
            m = np.isin(seg, np.asarray(scope, np.uint32))
            sc = F[m].astype(np.int32) @ q.astype(np.int32)
            ids = gid[m]
            order = np.lexsort((ids, -sc.astype(np.int64)))[:k]
            return ids[order].tolist(), sc[order].tolist()

        rng = np.random.default_rng(1)
        nq = 64
        Q = np.stack([co.synth_query(4, i, zb) for i in range(nq)])
        scopes = [[int(rng.integers(0, n_st))] for _ in range(nq)]
        scopes[5] = [1, 2, 9]                         # spans all three engines
        scopes[6] = [3, 99]                           # an unknown store contributes nothing
        scopes[7] = []                                # empty scope
        ids, sc, cs, cnt = g.search(Q, scopes, k=k)
        for i in range(nq):
            w_ids, w_sc = want(Q[i], [s for s in scopes[i] if s < n_st])
            assert int(cnt[i]) == len(w_ids) and sc[i][:cnt[i]].tolist() == w_sc, i
            assert ids[i][:cnt[i]].tolist() == w_ids, i
            assert (ids[i][cnt[i]:] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
        # text ingest goes to the store's engine; its chunk ids come from that engine's range
        doc = (" ".join(f"tenant seven report word{i}" for i in range(400))).encode()
        first, n_chunks, spans = g.ingest_text(7, 1001, doc)
        assert n_chunks > 0 and (7 % 3) * stride <= first < (7 % 3 + 1) * stride
        h_ids, h_sc, _, q = g.search_text(b"tenant seven report", [7], k)
        assert (q == co.query_vector(b"tenant seven report")).all()
        assert first <= int(h_ids[0]) < first + n_chunks
        g.tombstone_doc(1001)
        h2 = g.search_text(b"tenant seven report", [7], k)
        assert not any(first <= int(x) < first + n_chunks for x in h2[0])
        g.drop_store(3)
        assert g.lookup_store("fileSearchStores/t3") is None
        ids, sc, cs, cnt = g.search(Q[:2], [[3], [3, 4]], k=k)
        assert int(cnt[0]) == 0 and ids[1][:cnt[1]].tolist() == want(Q[1], [4])[0]
        # RF-1w statistic = sum over engines
        df, nn = g.scope_df([1, 2, 9])
        w_df, w_n = co.bucket_df(F, seg, [1, 2, 9])
        assert nn == w_n and (df == w_df).all()


def test_group_behind_the_adapter_equals_single_engine(tmp_path, co):
    """B200Rag over an engine group returns the single-engine citations bit for bit (uri aside: chunk ids are
    global per group): same documents, same questions, tf and idf scoring, metadata filter, delete, snapshot."""
    from rag_foundation_b200 import B200Rag, Engine, EngineGroup
    from rag_foundation_b200.adapter import Registry
    rng = np.random.default_rng(7)
    vocab = [f"term{i}" for i in range(300)]
    docs = []
    for d in range(24):
        words = rng.choice(vocab, size=int(rng.integers(150, 900)))
        p = tmp_path / f"doc{d}.txt"
        p.write_text(" ".join(words))
        docs.append(p)
    single = B200Rag(registry=Registry(Engine(capacity_rows=4096)))
    for placement in ("spread", "store"):
        group = B200Rag(registry=Registry(EngineGroup(_group_devices(2), capacity_rows=4096, placement=placement)))
        names = []
        for rag in (single, group):
            ss = [rag.create_store("a"), rag.create_store("b")]
            names.append(ss)
            for d, p in enumerate(docs):
                rag.upload_file(ss[d % 2], str(p), display_name=p.name, custom_metadata=[{"key": "team", "string_value": "ops" if d % 3 else "dev"}])
        for scoring in ("tf", "idf"):
            single.scoring = group.scoring = scoring
            for qi in range(12):
                text = " ".join(rng.choice(vocab, size=6))
                for scope_pick, mf in ((slice(0, 1), None), (slice(0, 2), None), (slice(0, 2), {"team": "dev"})):
                    a = single.retrieve(text, names[0][scope_pick], metadata_filter=mf)
                    b = group.retrieve(text, names[1][scope_pick], metadata_filter=mf)
                    # chunk ids differ between the two layouts, so rows that TIE on the score may rank (and, at the
                    # k-th score, be selected) differently: the scores must agree position by position, and the
                    # citations above the k-th score as sets
                    sa, sb = [h["score"] for h in a], [h["score"] for h in b]
                    assert sa == sb, (placement, scoring, text)
                    cut = sa[-1] if len(sa) == single.top_k else -1
                    strip = lambda hits: sorted((h["title"], h["text"], h["score"], h["cosine"], h["uri"].split("/")[-1].split("#")[0])   # noqa: E731
                                                for h in hits if h["score"] > cut)
                    assert strip(a) == strip(b), (placement, scoring, text)
        # snapshot round trip of the group registry
        group._reg.save(str(tmp_path / f"snap-{placement}"))
        reg2 = Registry.load(EngineGroup(_group_devices(2), capacity_rows=4096, placement=placement), str(tmp_path / f"snap-{placement}"))
        again = B200Rag(registry=reg2, scoring=group.scoring)
        text = "term1 term2 term3 term4"
        assert again.retrieve(text, names[1]) == group.retrieve(text, names[1])
        reg2.engine.close()
        for rag, ss in ((single, names[0]), (group, names[1])):
            rag.delete_store(ss[0]); rag.delete_store(ss[1])
        group._reg.engine.close()
    single._reg.engine.close()


# ------------------------------------------------------------------ deletes give their rows back
def test_ingest_delete_ingest_at_capacity(co):
    """A delete-heavy tenant does not exhaust the arena: freed rows (and their chunk ids) are reused, searches
    stay exact, and a concurrent reader never sees a half-written reused row as a hit of the wrong store."""
    from rag_foundation_b200._capi import RfError, RF_ECAPACITY
    text = lambda tag, n: (" ".join(f"{tag}{i % 50} filler{i}" for i in range(n))).encode()   # noqa: E731
    with _engine(64) as e:
        a = e.open_store("fileSearchStores/a"); b = e.open_store("fileSearchStores/b")
        first1, n1, _ = e.ingest_text(a, 1, text("alpha", 1100))
        first2, n2, _ = e.ingest_text(b, 2, text("beta", 1100))
        first3, n3, _ = e.ingest_text(a, 3, text("gamma", 1100))
        assert n1 == n2 == n3 == 20 and 3 * n1 <= 64 < 4 * n1            # the arena cannot take a 4th document
        assert e.stats()["free_rows"] == 0
        stop = threading.Event()
        wrong = []

        def reader():
            while not stop.is_set():
                ids, sc, _, _ = e.search_text(b"beta1 beta2 beta3", [b], 10)
                wrong.extend(int(x) for x in ids if not (first2 <= int(x) < first2 + n2))
        th = threading.Thread(target=reader)
        th.start()
        try:
            for rnd in range(12):                     # far more rows than the arena holds, in total
                e.tombstone_doc(1 if rnd == 0 else 100 + rnd - 1)
                assert e.stats()["free_rows"] == n1
                f, n, _ = e.ingest_text(a, 100 + rnd, text(f"r{rnd}x", 1100))
                assert (f, n) == (first1, n1), "the freed run is reused (same chunk ids)"
                assert e.stats()["free_rows"] == 0
                ids, sc, _, q = e.search_text(f"r{rnd}x1 r{rnd}x2".encode(), [a], 10)
                F, seg, ff = e.read_rows(0, e.stats()["n_rows"])
                w_ids, w_sc, _ = co.score_topk(F, seg, q, [a], k=10, ff=ff)
                assert ids.tolist() == w_ids.tolist() and sc.tolist() == w_sc.tolist()
                assert all(first1 <= int(x) < first1 + n1 or first3 <= int(x) < first3 + n3 for x in ids)
        finally:
            stop.set()
            th.join()
        assert wrong == []
        # a document larger than any freed run still fails loudly
        e.tombstone_doc(111)
        with pytest.raises(RfError) as err:
            e.ingest_text(a, 999, text("big", 3300))
        assert err.value.code == RF_ECAPACITY
        # dropping a store frees its rows too, and a deleted document's rows leave its store's extents
        e.drop_store(b)
        assert e.stats()["free_rows"] == n1 + n2
        c = e.open_store("fileSearchStores/c")
        f, n, _ = e.ingest_text(c, 500, text("delta", 2000))          # 36 chunks: only the coalesced run of both freed documents takes it
        assert f == first1 and n1 < n <= n1 + n2
        ids, _, _, _ = e.search_text(b"delta1 delta2", [c], 10)
        assert len(ids) and all(f <= int(x) < f + n for x in ids)
        ids, _, _, _ = e.search_text(b"delta1 delta2", [a], 10)      # store a no longer reaches the rows it gave back
        assert all(first3 <= int(x) < first3 + n3 for x in ids)


def test_snapshot_is_atomic_and_checksummed(tmp_path, co, zb):
    from rag_foundation_b200._capi import RfError
    n = 20_000
    path = str(tmp_path / "index.rfsnap")
    with _engine(n + 100) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=2, start_counter=0, n_rows=n)
        e.ingest_text(s, 1, b"some words to delete later " * 100)
        e.tombstone_doc(1)
        e.save_snapshot(path)
        good = open(path, "rb").read()
        assert not os.path.exists(path + ".tmp")
        # a failed save (unwritable temporary) leaves the previous file untouched
        os.mkdir(path + ".tmp")
        with pytest.raises(RfError):
            e.save_snapshot(path)
        os.rmdir(path + ".tmp")
        assert open(path, "rb").read() == good
        q = co.synth_query(2, 0, zb)
        want = e.search(q[None], [[s]], k=10)
        free_rows = e.stats()["free_rows"]
    with _engine(n + 100) as e2:
        e2.load_snapshot(path)
        got = e2.search(q[None], [[0]], k=10)
        for x, y in zip(want, got):
            assert (x == y).all()
        assert e2.stats()["free_rows"] == free_rows
    for what, blob in (("flipped byte", good[:len(good) // 2] + bytes([good[len(good) // 2] ^ 0x40]) + good[len(good) // 2 + 1:]),
                       ("truncated", good[:-4096]), ("no trailer", good[:-16])):
        bad = str(tmp_path / "bad.rfsnap")
        open(bad, "wb").write(blob)
        with _engine(n + 100) as e3:
            with pytest.raises(RfError):
                e3.load_snapshot(bad)
            assert e3.stats()["n_rows"] == 0 and e3.lookup_store("fileSearchStores/a") is None, what
            s3 = e3.open_store("fileSearchStores/fresh")          # the engine is still usable and empty
            assert e3.search(q[None], [[s3]], k=10)[3][0] == 0


def test_device_searches_with_and_without_the_overlap_promise(co, zb):
    """Back-to-back device-resident searches: (a) default = fully serialised launches, the query vector may be
    produced by a kernel right before each search; (b) with rf_stream_set_overlap the launches overlap (PDL)
    over a resident query batch.  Both equal the oracle."""
    import torch
    n, k = 300_000, 10
    with _engine(n) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=6, start_counter=0, n_rows=n)
        F = co.synth_rows(6, 0, n, zb)
        seg = np.full(n, s, np.uint32)
        Q = np.stack([co.synth_query(6, i, zb) for i in range(48)])
        want = [co.score_topk_keys(F, seg, Q[i], [s], k=k).tolist() for i in range(48)]
        Qd = torch.from_numpy(Q).cuda()
        stream = torch.cuda.current_stream().cuda_stream
        # (a) the query is written by a torch kernel immediately before every search, into the SAME buffer
        qbuf = torch.zeros((1, 256), dtype=torch.int8, device="cuda")
        out = torch.zeros((48, k), dtype=torch.int64, device="cuda")
        for rep in range(3):
            for i in range(48):
                torch.add(Qd[i:i + 1], 0, out=qbuf)          # a kernel, not a copy
                e.search_keys_device(qbuf.data_ptr(), 1, [s], k, out[i].data_ptr(), stream)
            torch.cuda.synchronize()
            got = out.cpu().numpy().view(np.uint64)
            assert [g.tolist() for g in got] == want
        # (b) resident batch + promise
        e.set_stream_overlap(stream, True)
        out.zero_()
        for i in range(48):
            e.search_keys_device(Qd[i:i + 1].data_ptr(), 1, [s], k, out[i].data_ptr(), stream)
        torch.cuda.synchronize()
        assert [g.tolist() for g in out.cpu().numpy().view(np.uint64)] == want
        e.set_stream_overlap(stream, False)


def test_many_scopes_hold_no_device_state(co, zb):
    """Thousands of distinct scopes through the device-resident entry point on one stream: the engine keeps per-STREAM
    scratch only, so its device footprint does not grow with the number of tenants."""
    import torch
    rows, n_st = 64, 2_000
    with _engine(rows * n_st) as e:
        for i in range(n_st):
            e.open_store(f"fileSearchStores/t{i}")
        e.ingest_synthetic(0, rows, seed=1, start_counter=0, n_rows=rows * n_st)
        q = torch.from_numpy(co.synth_query(1, 0, zb)[None]).cuda()
        out = torch.zeros((1, 10), dtype=torch.int64, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        e.search_keys_device(q.data_ptr(), 1, [0], 10, out.data_ptr(), stream)
        torch.cuda.synchronize()
        free0 = torch.cuda.mem_get_info()[0]
        for i in range(n_st):
            e.search_keys_device(q.data_ptr(), 1, [i], 10, out.data_ptr(), stream)
        torch.cuda.synchronize()
        assert free0 - torch.cuda.mem_get_info()[0] < (8 << 20), "device memory grew with the number of scopes"
        F = co.synth_rows(1, (n_st - 1) * rows, rows, zb)
        want = co.score_topk_keys(F, np.full(rows, n_st - 1, np.uint32), co.synth_query(1, 0, zb), [n_st - 1], k=10, id_base=(n_st - 1) * rows)
        assert out.cpu().numpy().view(np.uint64)[0].tolist() == want.tolist()


def test_first_search_on_a_fresh_non_blocking_stream(co, zb):
    """A caller stream's scratch (tickets, floors, tile counters) is zeroed on THAT stream.  It used to be zeroed with
    cudaMemset -- the legacy default stream, which does not order with a non-blocking stream: with two store-sharded
    batches in flight per rank the zeroing landed inside the new stream's first batch (queries never answered, one block
    ticket left off for good; `tools/two_in_flight_check.py` under torchrun on 2 GPUs reproduces it with the old library
    and is clean with this one).  This single-GPU test walks the same path -- first searches on fresh non-blocking
    streams while the default stream is busy, through both device-resident entry points -- but the race window is too
    narrow here for it to fail reliably on the old code."""
    import torch
    n, k = 1_000_000, 10
    with _engine(n) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=3, start_counter=0, n_rows=n)
        Q = np.stack([co.synth_query(3, i, zb) for i in range(16)])
        qd = torch.from_numpy(Q).cuda()
        want = torch.zeros((16, k), dtype=torch.int64, device="cuda")
        for i in range(16):
            e.search_keys_device(qd[i:i + 1].data_ptr(), 1, [s], k, want[i].data_ptr(), torch.cuda.current_stream().cuda_stream)
        got = torch.zeros((256, k), dtype=torch.int64, device="cuda")
        gotb = torch.zeros((40, 16, k), dtype=torch.int64, device="cuda")
        csr = (np.full(16, s, np.uint32), np.arange(17, dtype=np.uint32))
        torch.cuda.synchronize()                           # (the result buffers are zeroed before the default stream gets busy)
        torch.cuda._sleep(int(4e-3 * 1.9e9))               # ~4 ms of work on the default stream
        fresh = torch.cuda.Stream()
        for i in range(256):                               # ~11 ms of searches: the default stream wakes up half-way
            e.search_keys_device(qd[i % 16:i % 16 + 1].data_ptr(), 1, [s], k, got[i].data_ptr(), fresh.cuda_stream)
        # ... and a store-scoped batch, whose scratch is set up by the other entry point, on another fresh stream
        torch.cuda._sleep(int(4e-3 * 1.9e9))
        fresh2 = torch.cuda.Stream()
        for i in range(40):
            e.search_keys_device_scoped(qd.data_ptr(), 16, csr, k, gotb[i].data_ptr(), fresh2.cuda_stream)
        torch.cuda.synchronize()
        assert torch.equal(got.view(16, 16, k), want.expand(16, 16, k))
        assert torch.equal(gotb, want.expand(40, 16, k))


# ------------------------------------------------------------------ staged, chunk-pipelined ingest
def _check_ingest(e, seg, doc_id, data, co, ptr=None):
    if ptr is None:
        first, n, spans = e.ingest_text(seg, doc_id, data)
    else:
        first, n = e.ingest_text_ptr(seg, doc_id, ptr, len(data))
        spans = None
    wF, wff, wsp, _ = co.featurize_doc(data)
    assert n == len(wF), (n, len(wF))
    F, sg, ff = e.read_rows(first - e.id_base, n)
    assert (F == wF).all() and (ff == wff).all() and (sg == seg).all()
    if spans is not None:
        assert (spans == wsp).all()
    return first, n


@pytest.mark.timeout(600)
def test_pipelined_ingest_sources_and_chunk_boundaries(co):
    """Documents larger than one 2 MB copy chunk go through the staging ring (helper threads + chunk-by-chunk
    DMA + one tokeniser launch per chunk, the running token count handed from launch to launch): rows, norms and spans
    equal the oracle for pageable, pinned (rf_host_alloc) and device-resident sources; tokens that straddle a
    chunk boundary, stop words at the boundary, and tokens LONGER than a whole chunk (hashed after the last
    copy) included."""
    import torch
    from rag_foundation_b200.engine import PinnedBuffer
    rng = np.random.default_rng(11)
    CH = 2 << 20
    words = [b"alpha", b"the", b"Beta9", b"a", b"an", b"gamma", b"x", b"\xc3\xa9t\xc3\xa9", b"Zeta_zeta", b"0042"]

    def running_text(n):
        parts, size = [], 0
        while size < n:
            w = words[int(rng.integers(0, len(words)))]
            parts.append(w)
            size += len(w) + 1
        return b" ".join(parts)[:n]

    docs = {
        "5 MB running text": running_text(5 * CH // 2 + 12345),
        "token across the first chunk boundary": running_text(CH - 3) + b"straddlingtoken" + b" " + running_text(CH),
        "stop word ends exactly at the boundary": running_text(CH - 4)[:CH - 4] + b" the" + b" next words here " + running_text(CH // 2),
        "stop-word prefix continues into the next chunk": running_text(CH - 2)[:CH - 2] + b" t" + b"heory of chunks " + running_text(CH // 2),
        "a 5 MB token": b"lead in " + b"q" * (5 * CH // 2) + b" tail words follow " + running_text(100_000),
        "document is one giant token": b"z" * (3 * CH + 17),
        "exactly two chunks": running_text(2 * CH),
        "one byte over a chunk": running_text(CH + 1),
    }
    with _engine(400_000) as e:
        s = e.open_store("fileSearchStores/ingest")
        doc_id = 1
        for name, data in docs.items():
            _check_ingest(e, s, doc_id, data, co)
            doc_id += 1
            pb = PinnedBuffer(len(data))
            pb.array[:] = np.frombuffer(data, np.uint8)
            _check_ingest(e, s, doc_id, data, co, ptr=pb.ptr)
            pb.close()
            doc_id += 1
            dd = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
            off = 1 if len(data) > 4 else 0                      # a device pointer need not be aligned
            torch.cuda.synchronize()
            _check_ingest(e, s, doc_id, data[off:], co, ptr=dd.data_ptr() + off)
            doc_id += 1
        # concurrent uploads from several threads (the ARQ worker runs up to 10, worker.py:125) serialise cleanly
        errs = []

        def up(i):
            try:
                _check_ingest(e, s, 1000 + i, docs["5 MB running text"][i * 1000:], co)
            except Exception as ex:   # noqa: BLE001
                errs.append(ex)
        ths = [threading.Thread(target=up, args=(i,)) for i in range(4)]
        [t.start() for t in ths]; [t.join() for t in ths]
        assert not errs, errs[:1]


@pytest.mark.timeout(600)
def test_ingest_spans_and_launches_of_the_tokeniser(co):
    """The tokeniser gives every CTA of a launch one contiguous span of the text (up to 8 KB, a multiple of 16
    bytes), keeps the span's token records in shared memory and places them once it has summed the token counts of
    the CTAs before it; a range longer than 4096 spans takes several launches that hand the running token count
    on.  Device-resident documents take ONE range: sizes around one span, a few spans, and the 32 MB where the
    second launch begins, with tokens / stop words straddling span edges by construction (running text), the
    densest token stream there is ("a b c ...": the record buffers are sized for it) and one token that runs
    from the first span to the last."""
    import torch
    rng = np.random.default_rng(23)
    words = [b"alpha", b"the", b"Beta9", b"a", b"an", b"gamma", b"x", b"\xc3\xa9t\xc3\xa9", b"Zeta_zeta", b"0042", b"The", b"theory"]

    def running_text(n):
        idx = rng.integers(0, len(words), n // 3 + 8)
        return b" ".join(words[int(i)] for i in idx)[:n]

    per_launch = 4096 * 8192
    sizes = [8191, 8192, 8192 + 16, 3 * 8192 + 4097, (1 << 20) + 13, per_launch - 1, per_launch, per_launch + 1, per_launch + 70_001]
    base = running_text(max(sizes))
    dense = b" ".join(bytes([98 + i % 24]) for i in range(40_000))                  # b c d ... : one kept token per 2 bytes
    giant = b"lead in " + b"q" * 100_000 + b" tail words follow " + running_text(5_000)
    with _engine(120_000) as e:
        s = e.open_store("fileSearchStores/spans")
        dd = torch.frombuffer(bytearray(base), dtype=torch.uint8).cuda()
        torch.cuda.synchronize()
        doc = 1
        for n in sizes:
            first, nc = _check_ingest(e, s, doc, base[:n], co, ptr=dd.data_ptr())
            assert nc > 0
            e.tombstone_doc(doc)      # the rows go back to the free list: the next document reuses them
            doc += 1
        for data in (dense, giant):
            for off in (0, 1, 7):
                _check_ingest(e, s, doc, data[off:], co)
                doc += 1


# ------------------------------------------------------------------ store table + store-sharded fused exchange
def test_store_table_batches_equal_host_built_plans(co, zb):
    """Batches of differently-scoped queries read their plans from the device-resident store table: same answers
    as host-built plans (RF_STORE_TABLE=0) and the oracle -- interleaved ingests (many extents per store), scopes
    of several stores, duplicates and unknown stores in a scope, a store with more extents than a plan holds
    (falls back), deletes and drops between batches (the table is rebuilt)."""
    import subprocess
    import sys
    import torch
    rows = 700
    n_st = 40
    with _engine(rows * n_st * 3 + 4096) as e:
        stores = [e.open_store(f"fileSearchStores/t{i}") for i in range(n_st)]
        parts, segs = [], []
        doc = 0
        for rnd in range(3):                       # interleaved: every store ends up with three extents
            for s in stores:
                blk = co.synth_rows(40 + rnd, s * rows, rows, zb)
                doc += 1
                e.ingest_features(s, doc, blk)
                parts.append(blk); segs.append(np.full(rows, s, np.uint32))
        for i in range(70):                        # one store with more extents than a plan holds
            for s in (stores[3], stores[4]):
                blk = co.synth_rows(50, (i * 2 + s) * 16, 16, zb)
                doc += 1
                e.ingest_features(s, doc, blk)
                parts.append(blk); segs.append(np.full(16, s, np.uint32))
        F = np.concatenate(parts); seg = np.concatenate(segs)
        rng = np.random.default_rng(3)
        nq = 96
        Q = np.stack([co.synth_query(41, i, zb) for i in range(nq)])

        def scopes_for(with_big):
            sc = []
            for i in range(nq):
                n = int(rng.integers(1, 5))
                pick = [int(x) for x in rng.choice([s for s in stores if with_big or s not in (3, 4)], size=n, replace=False)]
                if i % 7 == 0:
                    pick.append(pick[0])           # duplicate
                if i % 11 == 0:
                    pick.append(9999)              # unknown store
                sc.append(pick)
            return sc

        def check(scopes):
            l0 = e.stats()["kernel_launches"]
            ids, sc, cs, cnt = e.search(Q, scopes, k=10)
            assert e.stats()["kernel_launches"] - l0 == 1
            qd = torch.from_numpy(Q).cuda()
            out = torch.zeros((nq, 10), dtype=torch.int64, device="cuda")
            e.search_keys_device_scoped(qd.data_ptr(), nq, scopes, 10, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert (_keys(ids, sc) == out.cpu().numpy().view(np.uint64)).all()
            Fh, segh, ffh = e.read_rows(0, e.stats()["n_rows"])
            for i in range(nq):
                w_ids, w_sc, _ = co.score_topk(Fh, segh, Q[i], [s for s in scopes[i] if s < n_st], k=10, ff=ffh)
                m = len(w_ids)
                assert int(cnt[i]) == m and ids[i][:m].tolist() == w_ids.tolist() and sc[i][:m].tolist() == w_sc.tolist(), i
        check(scopes_for(False))                   # every scope fits: the table path
        check(scopes_for(True))                    # some scopes exceed 64 extents: the whole batch falls back, same answers
        e.tombstone_doc(5); e.tombstone_doc(45); e.drop_store(stores[7])
        check(scopes_for(False))                   # the table was rebuilt after the deletes
        first = e.ingest_features(stores[9], 9000, co.synth_rows(60, 0, 300, zb))
        sc9 = [[9]] * nq
        ids, sc, _, cnt = e.search(Q, sc9, k=10)
        Fh, segh, ffh = e.read_rows(0, e.stats()["n_rows"])
        w_ids, w_sc, _ = co.score_topk(Fh, segh, Q[0], [9], k=10, ff=ffh)
        assert ids[0].tolist() == w_ids.tolist() and first >= 0
    # A/B: the same first batch with host-built plans (the table disabled) gives the same keys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import numpy as np\n"
            "from oracle import c_oracle as co, rf1\n"
            "from rag_foundation_b200 import Engine\n"
            "zb = rf1.zipf_bucket_table()\n"
            "e = Engine(capacity_rows=20000)\n"
            "ss = [e.open_store('s%%d' %% i) for i in range(8)]\n"
            "e.ingest_synthetic(0, 2000, seed=1, start_counter=0, n_rows=16000)\n"
            "Q = np.stack([co.synth_query(1, i, zb) for i in range(32)])\n"
            "ids, sc, _, _ = e.search(Q, [[i %% 8, (i * 3) %% 8] for i in range(32)], k=10)\n"
            "print(ids.tobytes().hex()[:64], int(ids.astype(np.int64).sum()), int(sc.sum()))\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = [subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, RF_STORE_TABLE=v)) for v in ("1", "0")]
    assert all(o.returncode == 0 for o in outs), outs[0].stderr[-1500:] + outs[1].stderr[-1500:]
    assert outs[0].stdout == outs[1].stdout and outs[0].stdout.strip()


@pytest.mark.timeout(300)
def test_store_sharded_fused_exchange_two_engines_one_process(co, zb):
    """rf_search_keys_device_scoped_fused, world = 2 in one process (two engines on one GPU, whole stores per
    'rank'): a 256-query batch, most queries owned by one rank, some scopes spanning both; the publish-only scan
    + merge_wait pair on each rank returns the single-index answer on BOTH ranks."""
    import torch
    from rag_foundation_b200.engine import scopes_to_csr
    rows, n_st, k, nq, nq_cap, world = 3000, 16, 10, 256, 256, 2
    engines, streams = [], [torch.cuda.Stream() for _ in range(world)]
    try:
        for r in range(world):
            e = _engine(rows * n_st // world + 64, id_base=r * (1 << 30))
            engines.append(e)
            for g in range(n_st):
                e.open_store(f"fileSearchStores/t{g}")            # same numbering on both ranks; rows only on the owner
            for g in range(r, n_st, world):
                e.ingest_synthetic(g, 0, seed=70, start_counter=g * rows, n_rows=rows)
        F = co.synth_rows(70, 0, rows * n_st, zb)
        seg = (np.arange(rows * n_st) // rows).astype(np.uint32)
        store_of = np.arange(rows * n_st) // rows
        gid = (store_of % world) * (1 << 30) + (store_of // world) * rows + np.arange(rows * n_st) % rows
        rng = np.random.default_rng(9)
        Q = np.stack([co.synth_query(70, i, zb) for i in range(nq)])
        scopes = [[int(rng.integers(0, n_st))] for _ in range(nq)]
        for i in range(0, nq, 9):
            scopes[i] = [int(x) for x in rng.choice(n_st, size=3, replace=False)]      # spans both ranks
        scopes[1] = []
        csr = scopes_to_csr(scopes)
        keys_buf = [torch.zeros(4 * world * nq_cap * k, dtype=torch.int64, device="cuda") for _ in range(world)]
        flag_buf = [torch.zeros(4 * world * nq_cap, dtype=torch.int32, device="cuda") for _ in range(world)]
        timeout = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(world)]
        kp = np.asarray([t.data_ptr() for t in keys_buf], np.uint64)
        fp = np.asarray([t.data_ptr() for t in flag_buf], np.uint64)
        qd = torch.from_numpy(Q).cuda()
        outs = [torch.zeros((nq, k), dtype=torch.int64, device="cuda") for _ in range(world)]
        # what each 'rank' launches: the queries it owns stores of, and for every query the ranks its merge waits for
        masks = np.zeros(nq, np.uint8)
        per_rank = []
        for r in range(world):
            q_index, local = [], []
            for i, sc in enumerate(scopes):
                mine = [g for g in sc if g % world == r]
                if mine:
                    masks[i] |= 1 << r
                    q_index.append(i)
                    local.append(mine)
            per_rank.append((np.asarray(q_index, np.uint32), scopes_to_csr(local)))
        assert int((masks == 3).sum()) > 5 and int((masks == 0).sum()) == 1
        # warm-up without an exchange so the per-stream scratch exists (see the chunk-sharded fused test)
        for r in range(world):
            engines[r].search_keys_device_scoped_fused(qd.data_ptr(), nq, csr, k, outs[r].data_ptr(), streams[r].cuda_stream,
                                                       0, 1, nq_cap, 1, kp[r:r + 1], fp[r:r + 1], timeout[r].data_ptr())
        torch.cuda.synchronize()
        for t in flag_buf:
            t.zero_()
        torch.cuda.synchronize()
        for seq in range(1, 7):
            for o in outs:
                o.zero_()
            torch.cuda.synchronize()
            for r in range(world):
                q_index, local_csr = per_rank[r]
                engines[r].search_keys_device_scoped_fused(qd.data_ptr(), int(q_index.size), local_csr, k, outs[r].data_ptr(), streams[r].cuda_stream,
                                                           r, world, nq_cap, seq, kp, fp, timeout[r].data_ptr(),
                                                           nq_total=nq, q_index=q_index, owner_masks=masks)
            torch.cuda.synchronize()
            assert not any(int(t.item()) for t in timeout)
            got = [o.cpu().numpy().view(np.uint64) for o in outs]
            assert (got[0] == got[1]).all()
        # two batches in flight per 'rank' (each rank alternates between two caller streams, both ranks identically, nothing
        # synchronises in between): a gather slot -- four per rank, indexed by the call number -- is free again when it comes
        # round, and every call returns the answer above
        streams_b = [torch.cuda.Stream() for _ in range(world)]
        for r in range(world):                                     # the second streams' scratch, set up outside an exchange
            engines[r].search_keys_device_scoped_fused(qd.data_ptr(), nq, csr, k, outs[r].data_ptr(), streams_b[r].cuda_stream,
                                                       0, 1, nq_cap, 1, kp[r:r + 1], fp[r:r + 1], timeout[r].data_ptr())
        torch.cuda.synchronize()
        n_calls = 10
        outs2 = [[torch.zeros((nq, k), dtype=torch.int64, device="cuda") for _ in range(n_calls)] for _ in range(world)]
        torch.cuda.synchronize()
        for c in range(n_calls):
            for r in range(world):
                q_index, local_csr = per_rank[r]
                st = (streams[r], streams_b[r])[c & 1]
                engines[r].search_keys_device_scoped_fused(qd.data_ptr(), int(q_index.size), local_csr, k, outs2[r][c].data_ptr(), st.cuda_stream,
                                                           r, world, nq_cap, 7 + c, kp, fp, timeout[r].data_ptr(),
                                                           nq_total=nq, q_index=q_index, owner_masks=masks)
        torch.cuda.synchronize()
        assert not any(int(t.item()) for t in timeout)
        for r in range(world):
            for c in range(n_calls):
                assert (outs2[r][c].cpu().numpy().view(np.uint64) == got[0]).all(), (r, c)
        for i in range(nq):
            m = np.isin(seg, np.asarray(scopes[i], np.uint32))
            sc = F[m].astype(np.int32) @ Q[i].astype(np.int32)
            ids = gid[m]
            order = np.lexsort((ids, -sc.astype(np.int64)))[:k]
            want = (sc[order].astype(np.int64).astype(np.uint64) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - ids[order].astype(np.uint64))
            assert got[0][i][:len(want)].tolist() == want.tolist() and (got[0][i][len(want):] == 0).all(), i
    finally:
        for e in engines:
            e.close()
