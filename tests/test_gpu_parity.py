"""Parity of the CUDA path (through the C-ABI, include/rf_b200.h) with the CPU oracle.
Bit-exact for chunk ids, int32 scores, int8 features and byte spans; cosine within 1e-5 relative
(the tolerance BASELINE.json states).  Run on the GPU box: pytest -m gpu."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

COS_RTOL = 1e-5


@pytest.fixture(scope="module")
def co():
    from oracle import c_oracle
    return c_oracle


@pytest.fixture(scope="module")
def rf1():
    from oracle import rf1 as m
    return m


@pytest.fixture(scope="module")
def zb(rf1):
    return rf1.zipf_bucket_table()


@pytest.fixture(scope="module")
def golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, "rf1_golden.json")))


def _dense(sparse, n=256):
    row = np.zeros(n, np.int8)
    for i, v in sparse:
        row[i] = v
    return row


def _engine(cap, **kw):
    from rag_foundation_b200 import Engine
    return Engine(capacity_rows=cap, **kw)


def _check(co, eng_out, F, seg, q, scope, k, id_base, ff):
    ids, sc, cs, cnt = eng_out
    w_ids, w_sc, w_cs = co.score_topk(F, seg, q, scope, k=k, id_base=id_base, ff=ff)
    m = len(w_ids)
    assert int(cnt) == m
    assert ids[:m].tolist() == w_ids.tolist()
    assert sc[:m].tolist() == w_sc.tolist()
    np.testing.assert_allclose(cs[:m], w_cs, rtol=COS_RTOL, atol=0)
    assert (ids[m:] == np.uint64(0xFFFFFFFFFFFFFFFF)).all() and (sc[m:] == 0).all()


# ------------------------------------------------------------------ synthetic corpus generator
def test_synthetic_rows_bit_exact(co, zb):
    with _engine(70_000) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=5, start_counter=123_456, n_rows=65_537)
        F, seg, ff = e.read_rows(0, 65_537)
        wF, wff = co.synth_rows(5, 123_456, 65_537, zb, with_ff=True)
        assert (F == wF).all() and (ff == wff).all() and (seg == s).all()


# ------------------------------------------------------------------ single query scans
@pytest.mark.parametrize("n_rows", [1, 31, 32, 33, 1000, 20_000, 262_144 + 17])
def test_single_query_parity_sizes(co, zb, n_rows):
    with _engine(n_rows + 64, id_base=1000) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=1, start_counter=0, n_rows=n_rows)
        F, ff = co.synth_rows(1, 0, n_rows, zb, with_ff=True)
        seg = np.full(n_rows, s, np.uint32)
        for qi in range(4):
            q = co.synth_query(1, qi, zb)
            for k in (1, 10, 32):
                ids, sc, cs, cnt = e.search(q[None, :], [[s]], k=k)
                _check(co, (ids[0], sc[0], cs[0], cnt[0]), F, seg, q, [s], k, 1000, ff)


def test_golden_top10_through_cuda(golden):
    for case in golden["synth_top10"]:
        with _engine(case["n_rows"]) as e:
            s = e.open_store("fileSearchStores/a")
            e.ingest_synthetic(s, 0, seed=case["seed"], start_counter=0, n_rows=case["n_rows"])
            from oracle import c_oracle, rf1
            q = c_oracle.synth_query(case["seed"], case["qi"], rf1.zipf_bucket_table())
            ids, sc, _, cnt = e.search(q[None, :], [[s]], k=10)
            assert ids[0].tolist() == case["ids"] and sc[0].tolist() == case["scores"]


@pytest.mark.parametrize("seed", list(range(8)))
def test_one_million_chunks_parity(co, zb, seed):
    """BASELINE.json configs[1]: synthetic single store, 1M chunks, 1 query at a time, top-10."""
    n = 1_000_000
    with _engine(n) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=seed, start_counter=0, n_rows=n)
        F, ff = co.synth_rows(seed, 0, n, zb, with_ff=True)
        seg = np.full(n, s, np.uint32)
        for qi in range(3):
            q = co.synth_query(seed, qi, zb)
            ids, sc, cs, cnt = e.search(q[None, :], [[s]], k=10)
            _check(co, (ids[0], sc[0], cs[0], cnt[0]), F, seg, q, [s], 10, 0, ff)


def test_tiny_golden_cases_via_ingest_features(golden):
    g = golden["tiny"]
    F = np.stack([_dense(r) for r in g["F_sparse"]])
    q = _dense(g["q_sparse"])
    segs = g["store_seg"]
    for case in g["cases"]:
        with _engine(64, id_base=case["id_base"]) as e:
            for i in range(3):
                assert e.open_store(f"fileSearchStores/s{i}") == i
            doc = 1
            for r, sg in enumerate(segs):   # one document per row so the tombstoned row can be deleted
                e.ingest_features(sg if sg != 0xFFFFFFFF else 0, doc + r, F[r:r + 1])
            e.tombstone_doc(doc + segs.index(0xFFFFFFFF))
            scope = [x for x in case["scope"] if x != 0xFFFFFFFF]
            ids, sc, cs, cnt = e.search(q[None, :], [scope], k=case["k"])
            m = len(case["ids"])
            assert int(cnt[0]) == m, case["name"]
            assert ids[0, :m].tolist() == case["ids"] and sc[0, :m].tolist() == case["scores"], case["name"]


def test_adversarial_orderings(co):
    """Ascending scores (every row beats the running threshold), all-equal scores (pure id
    tie-break) and all-zero scores."""
    n = 40_000
    rng = np.random.default_rng(0)
    q = np.zeros(256, np.int8); q[3] = 1; q[200] = 2
    for name in ("ascending", "all_equal", "all_zero", "descending"):
        F = np.zeros((n, 256), np.int8)
        if name == "ascending":
            F[:, 3] = (np.arange(n) * 127 // n).astype(np.int8); F[:, 200] = (np.arange(n) % 100).astype(np.int8)
        elif name == "descending":
            F[:, 3] = ((n - 1 - np.arange(n)) * 127 // n).astype(np.int8)
        elif name == "all_equal":
            F[:, 3] = 5
        F[:, 17] = rng.integers(0, 127, n).astype(np.int8)   # does not touch the score
        with _engine(n) as e:
            s = e.open_store("fileSearchStores/a")
            e.ingest_features(s, 1, F)
            ff = (F.astype(np.int32) ** 2).sum(1).astype(np.int32)
            seg = np.full(n, s, np.uint32)
            ids, sc, cs, cnt = e.search(q[None, :], [[s]], k=10)
            _check(co, (ids[0], sc[0], cs[0], cnt[0]), F, seg, q, [s], 10, 0, ff)
            if name == "all_equal":
                assert ids[0].tolist() == list(range(10))


# ------------------------------------------------------------------ tenant mask, tombstones, extents
def test_store_mask_interleaved_and_multi_store_scope(co, zb):
    rng = np.random.default_rng(3)
    with _engine(60_000) as e:
        stores = [e.open_store(f"fileSearchStores/t{i}") for i in range(5)]
        F_all, seg_all = [], []
        doc = 0
        start = 0
        for _ in range(120):   # many small documents, stores interleaved -> many extents per store
            s = int(rng.integers(0, 5))
            n = int(rng.integers(1, 400))
            rows = co.synth_rows(9, start, n, zb)
            doc += 1
            e.ingest_features(stores[s], doc, rows)
            F_all.append(rows); seg_all.append(np.full(n, stores[s], np.uint32))
            start += n
        F = np.concatenate(F_all); seg = np.concatenate(seg_all)
        ff = (F.astype(np.int32) ** 2).sum(1).astype(np.int32)
        scopes = [[stores[0]], [stores[1], stores[3]], stores, [stores[4], stores[4]], [77], []]
        for qi in range(3):
            q = co.synth_query(9, qi, zb)
            ids, sc, cs, cnt = e.search(np.stack([q] * len(scopes)), scopes, k=10)
            for i, scope in enumerate(scopes):
                _check(co, (ids[i], sc[i], cs[i], cnt[i]), F, seg, q, scope, 10, 0, ff)
        # delete a few documents, then a whole store
        for d in (3, 40, 77):
            e.tombstone_doc(d)
        lo = 0
        for d, rows in enumerate(F_all, start=1):
            if d in (3, 40, 77):
                seg[lo:lo + len(rows)] = 0xFFFFFFFF
            lo += len(rows)
        e.drop_store(stores[1])
        seg[seg == stores[1]] = 0xFFFFFFFF
        q = co.synth_query(9, 0, zb)
        for scope in ([stores[0]], [stores[1]], stores):
            ids, sc, cs, cnt = e.search(q[None, :], [scope], k=10)
            _check(co, (ids[0], sc[0], cs[0], cnt[0]), F, seg, q, scope, 10, 0, ff)
        with pytest.raises(RuntimeError):
            e.tombstone_doc(3)   # already gone


def test_many_extents_are_coalesced_but_exact(co, zb):
    """> 64 extents for one store: the engine widens the scan ranges, the row mask keeps it exact."""
    with _engine(20_000) as e:
        a = e.open_store("fileSearchStores/a"); b = e.open_store("fileSearchStores/b")
        F_all, seg_all = [], []
        start = 0
        for d in range(300):
            n = 7 + d % 13
            rows = co.synth_rows(4, start, n, zb)
            s = a if d % 2 == 0 else b
            e.ingest_features(s, d + 1, rows)
            F_all.append(rows); seg_all.append(np.full(n, s, np.uint32)); start += n
        F = np.concatenate(F_all); seg = np.concatenate(seg_all)
        ff = (F.astype(np.int32) ** 2).sum(1).astype(np.int32)
        q = co.synth_query(4, 1, zb)
        for scope in ([a], [b], [a, b]):
            ids, sc, cs, cnt = e.search(q[None, :], [scope], k=10)
            _check(co, (ids[0], sc[0], cs[0], cnt[0]), F, seg, q, scope, 10, 0, ff)


def test_multi_tenant_batch_store_scoped(co, zb):
    """BASELINE.json configs[4] in miniature: store-sorted corpus, each query scoped to one store."""
    n_stores, per = 200, 1000
    with _engine(n_stores * per) as e:
        first = e.open_store("fileSearchStores/m0")
        for i in range(1, n_stores):
            e.open_store(f"fileSearchStores/m{i}")
        e.ingest_synthetic(first, per, seed=6, start_counter=0, n_rows=n_stores * per)
        F, ff = co.synth_rows(6, 0, n_stores * per, zb, with_ff=True)
        seg = (np.arange(n_stores * per) // per).astype(np.uint32) + first
        rng = np.random.default_rng(1)
        nq = 64
        scopes = [[int(first + rng.integers(0, n_stores))] for _ in range(nq)]
        Q = np.stack([co.synth_query(6, i, zb) for i in range(nq)])
        ids, sc, cs, cnt = e.search(Q, scopes, k=10)
        for i in range(nq):
            _check(co, (ids[i], sc[i], cs[i], cnt[i]), F, seg, Q[i], scopes[i], 10, 0, ff)


def test_batched_queries_shared_store(co, zb):
    n = 50_000
    with _engine(n) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=2, start_counter=0, n_rows=n)
        F, ff = co.synth_rows(2, 0, n, zb, with_ff=True)
        seg = np.full(n, s, np.uint32)
        Q = np.stack([co.synth_query(2, i, zb) for i in range(33)])
        ids, sc, cs, cnt = e.search(Q, [[s]] * 33, k=10)
        for i in range(33):
            _check(co, (ids[i], sc[i], cs[i], cnt[i]), F, seg, Q[i], [s], 10, 0, ff)


# ------------------------------------------------------------------ device-resident path + merge
def test_device_keys_path_and_merge(co, zb):
    import torch
    from rag_foundation_b200 import unpack_keys
    n = 100_000
    with _engine(n, id_base=500) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=3, start_counter=0, n_rows=n)
        F = co.synth_rows(3, 0, n, zb)
        seg = np.full(n, s, np.uint32)
        Q = np.stack([co.synth_query(3, i, zb) for i in range(6)])
        qd = torch.from_numpy(Q).cuda()
        out = torch.zeros((6, 10), dtype=torch.int64, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        for _ in range(3):   # cached plan reuse
            e.search_keys_device(qd.data_ptr(), 6, [s], 10, out.data_ptr(), stream)
        torch.cuda.synchronize()
        keys = out.cpu().numpy().view(np.uint64)
        for i in range(6):
            want = co.score_topk_keys(F, seg, Q[i], [s], k=10, id_base=500)
            assert keys[i].tolist() == want.tolist()
        # merge: split each list in 3 interleaved parts + a duplicate list, shuffle -> same answer
        parts = np.zeros((4, 6, 10), np.uint64)
        for i in range(6):
            for j in range(10):
                parts[j % 3, i, j // 3] = keys[i, j]
            parts[3, i, :4] = keys[i, :4]
        pd = torch.from_numpy(parts.view(np.int64)).cuda()
        merged = torch.zeros((6, 10), dtype=torch.int64, device="cuda")
        e.merge_topk_device(pd.data_ptr(), 4, 6, 10, merged.data_ptr(), stream)
        torch.cuda.synchronize()
        assert (merged.cpu().numpy().view(np.uint64) == keys).all()
        ids, sc, valid = unpack_keys(keys)
        assert valid.all() and (ids >= 500).all()


def test_sharded_searcher_single_process_two_engines(co, zb):
    """Two shards on one GPU, merged with the CUDA merge kernel, equals the unsharded scan."""
    import torch
    from rag_foundation_b200.sharded import ShardedSearcher, shard_range, unpack_keys_torch
    n = 60_001
    engines = []
    try:
        local = []
        Q = np.stack([co.synth_query(8, i, zb) for i in range(4)])
        qd = torch.from_numpy(Q).cuda()
        for r in range(2):
            lo, hi = shard_range(n, r, 2)
            e = _engine(hi - lo, id_base=lo); engines.append(e)
            s = e.open_store("fileSearchStores/a")
            e.ingest_synthetic(s, 0, seed=8, start_counter=lo, n_rows=hi - lo)
            srch = ShardedSearcher.for_engine(e)
            local.append(srch.local_search(qd, [s], 10))
        gathered = torch.stack(local).contiguous()
        merged = ShardedSearcher.for_engine(engines[0]).merge(gathered, 10)
        torch.cuda.synchronize()
        ids, sc, valid = unpack_keys_torch(merged.cpu())
        F = co.synth_rows(8, 0, n, zb)
        for i in range(4):
            w_ids, w_sc, _ = co.score_topk(F, np.zeros(n, np.uint32), Q[i], [0])
            assert ids[i].tolist() == w_ids.tolist() and sc[i].tolist() == w_sc.tolist()
        # RF-1w: the shards' device-side statistics add up to the whole corpus's
        from rag_foundation_b200.sharded import engine_local_df
        stat = sum(engine_local_df(e)([0]) for e in engines).cpu().numpy()
        w_df, w_n = co.bucket_df(F, np.zeros(n, np.uint32), [0])
        assert int(stat[-1]) == w_n and (stat[:-1].astype(np.uint64) == w_df).all()
    finally:
        for e in engines:
            e.close()


# ------------------------------------------------------------------ featurisation (ingest + query)
def _check_doc(e, seg, doc_id, data, oracle_mod, first_expected):
    first, n, spans = e.ingest_text(seg, doc_id, data)
    wF, wff, wsp, _ = oracle_mod.featurize_doc(data)
    assert first == first_expected and n == len(wF)
    if n:
        F, sg, ff = e.read_rows(first - e.id_base, n)
        assert (F == wF).all() and (ff == wff).all() and (sg == seg).all()
        assert (spans == wsp).all()
    return n


def test_ingest_text_golden_documents(golden, co):
    with _engine(4096) as e:
        s = e.open_store("fileSearchStores/a")
        nxt = 0
        for name in ("sample_report", "long_doc"):
            data = golden[name]["text"].encode("utf-8")
            n = _check_doc(e, s, hash(name) & 0xFFFF, data, co, nxt)
            assert n == golden[name]["n_chunks"]
            nxt += n
        sat = golden["saturation"]
        assert (e.featurize_query(sat["text"].encode()) == _dense(sat["q_sparse"])).all()


def test_ingest_text_edge_cases(co):
    cases = [b"", b"   \n\t ", b"the a an The AN", b"x", b"a", b"the", b"thee", b"ana", b"a1 an2 3the",
             b"The" * 2000, b"z" * 10_000, ("café naïve 中文 text " * 50).encode(),
             b" ".join(b"w%d" % i for i in range(128)), b" ".join(b"w%d" % i for i in range(129)),
             b" ".join(b"w%d" % i for i in range(144)), b" ".join(b"w%d" % i for i in range(145)),
             b"A" + b" " * 4095 + b"the" + b" " * 4093 + b"an bc",       # tokens straddling 4 KB block edges
             b"q" * 4095 + b" " + b"r" * 4097 + b" the", b"ab" * 2047 + b"the" + b" x",
             # ... and the 16 KB edges of a tokeniser block / the 4 KB rounds inside it
             b"A" + b" " * 16383 + b"the" + b" " * 16381 + b"an bc", b"w " * 8191 + b"the" + b" x y", b"q" * 16383 + b" " + b"r" * 16385 + b" the",
             (b"tok " * 4096)[:16383] + b"the an a tail", b" " * 16380 + b"edge" + b"word more" + b" " * 4090 + b"lastthe",
             # every token of a window in ONE bucket (count 128 saturates to 127), and in two buckets of the same 4-byte word
             b"zq " * 400, b"zq " * 127 + b"other " + b"zq " * 300, b" ".join(b"w%d" % (i % 4) for i in range(1000))]
    with _engine(8192) as e:
        s = e.open_store("fileSearchStores/a")
        nxt = 0
        for i, data in enumerate(cases):
            nxt += _check_doc(e, s, i + 1, data, co, nxt)
            assert (e.featurize_query(data) == co.query_vector(data)).all(), i


def test_ingest_text_fuzz(co):
    rng = np.random.default_rng(11)
    pieces = [b"the", b"a", b"an", b"The", b"AN", b" ", b"  ", b"\n", b",", b"-", b"_", b"x", b"9", b"ab", b"Zq7",
              "é".encode(), "中".encode(), b"theory", b"anagram", b"a1"]
    with _engine(40_000) as e:
        s = e.open_store("fileSearchStores/a")
        nxt = 0
        for i in range(40):
            n_p = int(rng.integers(0, 6000))
            data = b"".join(pieces[j] for j in rng.integers(0, len(pieces), n_p))
            nxt += _check_doc(e, s, i + 1, data, co, nxt)
            assert (e.featurize_query(data) == co.query_vector(data)).all()


def test_ingest_large_document(co, rf1):
    data = rf1.synth_text(0, 700_000)   # ~2.6 MB, ~6200 chunks
    assert len(data) > 2_000_000
    with _engine(8192) as e:
        s = e.open_store("fileSearchStores/a")
        n = _check_doc(e, s, 1, data, co, 0)
        assert n > 6000


def test_ingest_maximum_upload_size(co, rf1):
    """The reference caps uploads at 25 MB (backend/app/config.py:118): one document of that size,
    rows / norms / byte spans bit-exact, then the arena is full for anything more."""
    data = rf1.synth_text(3, 6_910_000)[:25 * 1024 * 1024]
    assert len(data) == 25 * 1024 * 1024
    wF, wff, wsp, ntok = co.featurize_doc(data)
    with _engine(len(wF)) as e:
        s = e.open_store("fileSearchStores/a")
        first, n, spans = e.ingest_text(s, 1, data)
        assert first == 0 and n == len(wF) and (spans == wsp).all()
        F, sg, ff = e.read_rows(0, n)
        assert (F == wF).all() and (ff == wff).all() and (sg == s).all()
        with pytest.raises(RuntimeError) as ei:
            e.ingest_text(s, 2, b"one more chunk")
        assert "arena full" in str(ei.value)


def test_search_text_equals_search_of_oracle_vector(co, golden):
    g = golden["sample_report"]
    with _engine(1024) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_text(s, 1, g["text"].encode())
        e.ingest_text(s, 2, golden["long_doc"]["text"].encode())
        ids, sc, cs, q = e.search_text(g["query"].encode(), [s], 10)
        assert (q == _dense(g["q_sparse"])).all()
        n = 1 + golden["long_doc"]["n_chunks"]
        F, seg, ff = e.read_rows(0, n)
        w_ids, w_sc, w_cs = co.score_topk(F, seg, q, [s], ff=ff)
        assert ids.tolist() == w_ids.tolist() and sc.tolist() == w_sc.tolist()
        np.testing.assert_allclose(cs, w_cs, rtol=COS_RTOL)
        assert ids[0] == 0 and sc[0] == g["scores"][0]
        np.testing.assert_allclose(cs[0], g["cos"][0], rtol=COS_RTOL)


# ------------------------------------------------------------------ errors and limits
def test_errors_are_loud():
    from rag_foundation_b200 import Engine
    with _engine(100) as e:
        s = e.open_store("fileSearchStores/a")
        assert e.open_store("fileSearchStores/a") == s and e.lookup_store("nope") is None
        with pytest.raises(RuntimeError):
            e.ingest_features(s, 1, np.zeros((101, 256), np.int8))       # capacity
        with pytest.raises(RuntimeError):
            e.ingest_features(99, 1, np.zeros((1, 256), np.int8))        # unknown store
        with pytest.raises(RuntimeError):
            e.search(np.zeros((1, 256), np.int8), [[s]], k=33)           # k > RF_TOPK_MAX
        with pytest.raises(ValueError):
            e.search(np.zeros((1, 256), np.int8), [list(range(17))], k=10)
        ids, sc, cs, cnt = e.search(np.zeros((1, 256), np.int8), [[s]], k=10)   # empty store
        assert cnt[0] == 0
    with pytest.raises(RuntimeError):
        Engine(capacity_rows=0)
    st_ok = _engine(10)
    st = st_ok.stats(); st_ok.close()
    assert st["capacity_rows"] == 10 and st["hbm_bytes"] >= 2640


def test_concurrent_searches_from_threads(co, zb):
    """routes/chat.py:520 runs ask_stream on one daemon thread per request (<= 50 per process)."""
    import threading
    n = 200_000
    with _engine(n, n_contexts=4) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=12, start_counter=0, n_rows=n)
        F = co.synth_rows(12, 0, n, zb)
        seg = np.full(n, s, np.uint32)
        Q = [co.synth_query(12, i, zb) for i in range(16)]
        want = [co.score_topk(F, seg, q, [s])[0].tolist() for q in Q]
        errs = []

        def work(i):
            try:
                for _ in range(5):
                    ids, _, _, _ = e.search(Q[i][None, :], [[s]], k=10)
                    assert ids[0].tolist() == want[i]
            except Exception as ex:   # noqa: BLE001
                errs.append(ex)
        ts = [threading.Thread(target=work, args=(i,)) for i in range(16)]
        [t.start() for t in ts]; [t.join() for t in ts]
        assert not errs, errs[:1]


# ------------------------------------------------------------------ batched tensor-core path (configs[2])
def _device_batch(e, Q, scope, k):
    import torch
    qd = torch.from_numpy(np.ascontiguousarray(Q)).cuda()
    out = torch.zeros((Q.shape[0], k), dtype=torch.int64, device="cuda")
    e.search_keys_device(qd.data_ptr(), Q.shape[0], scope, k, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("n_rows,nq,k", [(100_001, 300, 10), (65_536, 128, 10), (200_000, 1024, 10), (150_000, 513, 3), (70_000, 64, 1)])
def test_gemm_path_parity(co, zb, n_rows, nq, k):
    """tcgen05 int8 GEMM + fused top-k == oracle, including ragged M-tiles, a ragged last chunk tile
    and k < 10."""
    with _engine(n_rows + 200, id_base=7) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=21, start_counter=0, n_rows=n_rows)
        launches0 = e.stats()["kernel_launches"]
        F = co.synth_rows(21, 0, n_rows, zb)
        seg = np.full(n_rows, s, np.uint32)
        Q = np.stack([co.synth_query(21, i, zb) for i in range(nq)])
        keys = _device_batch(e, Q, [s], k)
        assert e.stats()["kernel_launches"] - launches0 in (4, 5), "the batched search should have taken the GEMM path"   # (pair kernel: + the tile-purity pass)
        step = max(1, nq // 48)
        for i in list(range(0, nq, step)) + [nq - 1]:
            want = co.score_topk_keys(F, seg, Q[i], [s], k=k, id_base=7)
            assert keys[i].tolist() == want.tolist(), i


@pytest.mark.parametrize("nq", [256, 600])      # single-CTA kernel / CTA-pair kernel
def test_gemm_path_mask_and_dense_queries(co, zb, nq):
    """Tombstoned rows inside the extent, a second store after it, and dense high-magnitude queries
    (many candidates, many ties)."""
    rng = np.random.default_rng(4)
    n = 50_000
    with _engine(3 * n) as e:
        a = e.open_store("fileSearchStores/a"); b = e.open_store("fileSearchStores/b")
        parts = [co.synth_rows(22, i * n // 2, n // 2, zb) for i in range(4)]
        e.ingest_features(a, 1, parts[0]); e.ingest_features(a, 2, parts[1]); e.ingest_features(a, 3, parts[2])
        e.ingest_features(b, 4, parts[3])
        e.tombstone_doc(2)
        F = np.concatenate(parts)
        seg = np.concatenate([np.full(n // 2, a), np.full(n // 2, 0xFFFFFFFF), np.full(n // 2, a), np.full(n // 2, b)]).astype(np.uint32)
        Q = rng.integers(0, 4, (nq, 256)).astype(np.int8)
        Q[:8] = 127                                   # saturating queries: huge scores
        Q[8:16] = 0                                   # all-zero queries: every score ties at 0
        Q[nq - 3:] = 127
        keys = _device_batch(e, Q, [a], 10)
        for i in list(range(0, nq, 9)) + [0, 8, nq - 1]:
            want = co.score_topk_keys(F, seg, Q[i], [a], k=10)
            assert keys[i].tolist() == want.tolist(), i


# ------------------------------------------------------------------ doc-level row restriction (metadata filters)
def test_search_text_in_ranges(co, zb):
    n = 30_000
    with _engine(n) as e:
        a = e.open_store("fileSearchStores/a"); b = e.open_store("fileSearchStores/b")
        F = co.synth_rows(31, 0, n, zb)
        e.ingest_features(a, 1, F[:10_000]); e.ingest_features(b, 2, F[10_000:20_000]); e.ingest_features(a, 3, F[20_000:])
        seg = np.concatenate([np.full(10_000, a), np.full(10_000, b), np.full(10_000, a)]).astype(np.uint32)
        text = b"1786 23 4479 313 12 7 1318 21"
        q = co.query_vector(text)
        for ranges in ([(100, 5_000)], [(0, 33), (9_990, 20_010), (29_999, 30_000)], [(10_000, 20_000)],
                       [(i * 500, i * 500 + 37) for i in range(60)]):           # 60 ranges: split over several launches
            ids, sc, cs, q_gpu = e.search_text(text, [a], 10, ranges=ranges)
            assert (q_gpu == q).all()
            keys = np.concatenate([co.score_topk_keys(F, seg, q, [a], k=10, row_lo=lo, row_hi=hi) for lo, hi in ranges])
            want = co.merge_topk(keys, 10)
            want = want[want != 0]
            from rag_foundation_b200 import unpack_keys
            w_ids, w_sc, _ = unpack_keys(want)
            assert ids.tolist() == w_ids.tolist() and sc.tolist() == w_sc.tolist(), ranges[:2]
        ids, sc, cs, _ = e.search_text(text, [a], 10, ranges=[])
        assert len(ids) == 0
        with pytest.raises(RuntimeError):
            e.search_text(text, [a], 10, ranges=[(50, 60), (10, 20)])          # unsorted


def test_back_to_back_queries_overlap_safely(co, zb):
    """64 different queries enqueued back to back on one stream with no synchronisation: under
    programmatic dependent launch each query's scan overlaps the previous query's merge tail, and
    the two alternating sync sets must keep them apart."""
    import torch
    n = 300_000
    with _engine(n) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=17, start_counter=0, n_rows=n)
        F = co.synth_rows(17, 0, n, zb)
        seg = np.full(n, s, np.uint32)
        Q = np.stack([co.synth_query(17, i, zb) for i in range(64)])
        qd = torch.from_numpy(Q).cuda()
        out = torch.zeros((64, 10), dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for rep in range(3):
            for i in range(64):
                e.search_keys_device(qd[i:i + 1].data_ptr(), 1, [s], 10, out[i].data_ptr(), st)
        torch.cuda.synchronize()
        keys = out.cpu().numpy().view(np.uint64)
        for i in range(64):
            assert keys[i].tolist() == co.score_topk_keys(F, seg, Q[i], [s]).tolist(), i


# ------------------------------------------------------------------ BASELINE.json full sizes, size-independent properties
@pytest.mark.timeout(900)
def test_full_size_100m_chunks_properties(co, zb):
    """configs[3]'s corpus (100 M chunks, 26 GB) on ONE GPU.  The CPU oracle cannot scan it in
    seconds, so the checks are size-independent: (a) every winner's score equals a CPU re-scoring
    of exactly that row (rows are regenerated from their counters); (b) the list is strictly ordered
    by the RF-1 key; (c) the same corpus split over four engines with id bases, merged by the
    merge kernel, gives the identical keys (two different decompositions of the scan agree);
    (d) the first million rows, which the oracle can scan, contribute exactly the oracle's top-10
    when the scope is narrowed to them by tombstoning nothing and searching a 1 M-row twin."""
    import torch
    from rag_foundation_b200 import unpack_keys
    n = 100_000_000
    free, _total = torch.cuda.mem_get_info()
    if free < 60 * (1 << 30):
        pytest.skip("needs ~56 GB of free HBM")
    Q = np.stack([co.synth_query(0, i, zb) for i in range(4)])
    qd = torch.from_numpy(Q).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    with _engine(n) as e:
        s = e.open_store("fileSearchStores/all")
        e.ingest_synthetic(s, 0, seed=0, start_counter=0, n_rows=n)
        out = torch.zeros((4, 10), dtype=torch.int64, device="cuda")
        for i in range(4):
            e.search_keys_device(qd[i:i + 1].data_ptr(), 1, [s], 10, out[i].data_ptr(), stream)
        torch.cuda.synchronize()
        whole = out.cpu().numpy().view(np.uint64)
        ids_h, sc_h, cs_h, cnt_h = e.search(Q, [[s]] * 4, k=10)            # host path agrees with the device path
    for i in range(4):
        ids, sc, valid = unpack_keys(whole[i])
        assert valid.all() and ids_h[i].tolist() == ids.tolist() and sc_h[i].tolist() == sc.tolist()
        assert all(int(whole[i][j]) > int(whole[i][j + 1]) for j in range(9))                       # (b)
        for gid, score in zip(ids.tolist(), sc.tolist()):                                            # (a)
            row = co.synth_rows(0, int(gid), 1, zb)[0]
            assert int(row.astype(np.int32) @ Q[i].astype(np.int32)) == score
    # (c) four shards + merge kernel
    parts = torch.zeros((4, 4, 10), dtype=torch.int64, device="cuda")
    engines = []
    try:
        for r in range(4):
            lo, hi = r * (n // 4), (r + 1) * (n // 4)
            er = _engine(hi - lo, id_base=lo)
            engines.append(er)
            sr = er.open_store("fileSearchStores/all")
            er.ingest_synthetic(sr, 0, seed=0, start_counter=lo, n_rows=hi - lo)
            for i in range(4):
                er.search_keys_device(qd[i:i + 1].data_ptr(), 1, [sr], 10, parts[r, i].data_ptr(), stream)
        merged = torch.zeros((4, 10), dtype=torch.int64, device="cuda")
        engines[0].merge_topk_device(parts.data_ptr(), 4, 4, 10, merged.data_ptr(), stream)
        torch.cuda.synchronize()
        assert (merged.cpu().numpy().view(np.uint64) == whole).all()
    finally:
        for er in engines:
            er.close()
    # (d) the oracle's own scan of the first 1 M rows vs a 1 M-row engine of the same counters
    F = co.synth_rows(0, 0, 1_000_000, zb)
    with _engine(1_000_000) as e1:
        s1 = e1.open_store("fileSearchStores/all")
        e1.ingest_synthetic(s1, 0, seed=0, start_counter=0, n_rows=1_000_000)
        ids1, sc1, _, _ = e1.search(Q, [[s1]] * 4, k=10)
    for i in range(4):
        w_ids, w_sc, _ = co.score_topk(F, np.zeros(1_000_000, np.uint32), Q[i], [0])
        assert ids1[i].tolist() == w_ids.tolist() and sc1[i].tolist() == w_sc.tolist()
        # every whole-corpus winner that lives in the first million must be in the 1 M top-10
        ids, _, _ = unpack_keys(whole[i])
        assert set(int(x) for x in ids if x < 1_000_000) <= set(w_ids.tolist())


@pytest.mark.timeout(900)
def test_full_size_multi_tenant_batch(co, zb):
    """configs[4] at full size: 10 000 stores x 10 000 chunks (26 GB), 1024 store-scoped queries in
    one call; every 8th answer is checked against the oracle's scan of just that store (10 k rows,
    regenerated from its counters), and every answer must stay inside its store (tenant mask)."""
    import torch
    from rag_foundation_b200.engine import scopes_to_csr
    free, _total = torch.cuda.mem_get_info()
    if free < 30 * (1 << 30):
        pytest.skip("needs ~28 GB of free HBM")
    n_stores, per, nq = 10_000, 10_000, 1024
    rng = np.random.default_rng(9)
    with _engine(n_stores * per) as e:
        first = e.open_store("fileSearchStores/t0")
        for i in range(1, n_stores):
            e.open_store(f"fileSearchStores/t{i}")
        e.ingest_synthetic(first, per, seed=5, start_counter=0, n_rows=n_stores * per)
        stores = rng.integers(0, n_stores, nq)
        scopes = [[int(first + st)] for st in stores]
        Q = np.stack([co.synth_query(5, i, zb) for i in range(nq)])
        ids, sc, cs, cnt = e.search(Q, scopes_to_csr(scopes), k=10)
        assert (cnt == 10).all()
        lo = (stores * per)[:, None]
        assert ((ids >= lo) & (ids < lo + per)).all(), "a citation left its tenant's store"
        for i in range(0, nq, 8):
            st = int(stores[i])
            F = co.synth_rows(5, st * per, per, zb)
            w_ids, w_sc, _ = co.score_topk(F, np.zeros(per, np.uint32), Q[i], [0], id_base=st * per)
            assert ids[i].tolist() == w_ids.tolist() and sc[i].tolist() == w_sc.tolist(), i


def test_host_batch_same_scope_takes_the_tensor_core_route(co, zb):
    """rf_search (host buffers) with >= 64 queries that share one scope: one upload of the queries,
    the batched GEMM search, an unpack kernel, one download -- same ids / scores / cosines as the
    oracle, and a mixed-scope batch of the same size still goes through the scan kernel."""
    n, nq = 120_000, 200
    with _engine(n + 8000) as e:
        a = e.open_store("fileSearchStores/a"); b = e.open_store("fileSearchStores/b")
        e.ingest_synthetic(a, 0, seed=23, start_counter=0, n_rows=n)
        e.ingest_synthetic(b, 0, seed=24, start_counter=0, n_rows=5000)
        F, ff = co.synth_rows(23, 0, n, zb, with_ff=True)
        Fb, ffb = co.synth_rows(24, 0, 5000, zb, with_ff=True)
        Fall = np.concatenate([F, Fb]); ffall = np.concatenate([ff, ffb])
        seg = np.concatenate([np.full(n, a), np.full(5000, b)]).astype(np.uint32)
        Q = np.stack([co.synth_query(23, i, zb) for i in range(nq)])
        l0 = e.stats()["kernel_launches"]
        ids, sc, cs, cnt = e.search(Q, [[a]] * nq, k=10)
        assert e.stats()["kernel_launches"] - l0 == 5, "expected 4 GEMM-path launches + the unpack kernel"
        for i in range(0, nq, 7):
            _check(co, (ids[i], sc[i], cs[i], cnt[i]), Fall, seg, Q[i], [a], 10, 0, ffall)
        scopes = [[a] if i % 2 else [b] for i in range(nq)]
        l0 = e.stats()["kernel_launches"]
        ids, sc, cs, cnt = e.search(Q, scopes, k=10)
        assert e.stats()["kernel_launches"] - l0 == 1
        for i in range(0, nq, 11):
            _check(co, (ids[i], sc[i], cs[i], cnt[i]), Fall, seg, Q[i], scopes[i], 10, 0, ffall)
        for scope in ([77], []):                                           # unknown store / empty scope, large same-scope batch
            ids, sc, cs, cnt = e.search(Q[:70], [scope] * 70, k=10)
            assert (cnt == 0).all() and (ids == np.uint64(0xFFFFFFFFFFFFFFFF)).all() and (sc == 0).all()
        # same scope for every query, but the store now has two extents with another store's rows in between
        e.ingest_features(a, 9, F[:3000])                                  # store a now has a second extent after b's rows
        Fall2 = np.concatenate([Fall, F[:3000]]); ff2 = np.concatenate([ffall, ff[:3000]])
        seg2 = np.concatenate([seg, np.full(3000, a, np.uint32)])
        l0 = e.stats()["kernel_launches"]
        ids, sc, cs, cnt = e.search(Q[:80], [[a]] * 80, k=10)
        # two extents: scored over their bounding range by the tensor-core kernel, store b's rows in between masked out
        assert e.stats()["kernel_launches"] - l0 == 5
        for i in range(0, 80, 9):
            _check(co, (ids[i], sc[i], cs[i], cnt[i]), Fall2, seg2, Q[i], [a], 10, 0, ff2)


# ------------------------------------------------------------------ RF-1w (IDF-weighted variant)
def test_scope_df_and_weights_match_oracle(co, zb):
    """Document frequencies over interleaved stores (many extents), after tombstones and a dropped
    store, with the cache refreshed by every change; weights are the C-ABI's own host function."""
    rng = np.random.default_rng(8)
    with _engine(80_000) as e:
        stores = [e.open_store(f"fileSearchStores/t{i}") for i in range(4)]
        F_all, seg_all = [], []
        start = 0
        for d in range(150):
            s = int(rng.integers(0, 4)); n = int(rng.integers(1, 500))
            rows = co.synth_rows(12, start, n, zb)
            e.ingest_features(stores[s], d + 1, rows)
            F_all.append(rows); seg_all.append(np.full(n, stores[s], np.uint32)); start += n
        F = np.concatenate(F_all); seg = np.concatenate(seg_all)

        def check_all():
            for scope in ([stores[0]], [stores[1], stores[3]], stores, [stores[2], stores[2]], [99], []):
                df, n = e.scope_df(scope)
                w_df, w_n = co.bucket_df(F, seg, scope)
                assert n == w_n and (df == w_df).all(), scope
                assert (e.idf_weights(df, n) == co.idf_weights(w_df, w_n)).all()
                df2, n2 = e.scope_df(scope)                     # second call: served from the cache
                assert n2 == n and (df2 == df).all()
        check_all()
        lo = 0
        for d, rows in enumerate(F_all, start=1):
            if d in (5, 60, 149):
                e.tombstone_doc(d)
                seg[lo:lo + len(rows)] = 0xFFFFFFFF
            lo += len(rows)
        check_all()
        e.drop_store(stores[3]); seg[seg == stores[3]] = 0xFFFFFFFF
        extra = co.synth_rows(12, start, 3000, zb)
        e.ingest_features(stores[0], 1000, extra)
        F = np.concatenate([F, extra]); seg = np.concatenate([seg, np.full(3000, stores[0], np.uint32)])
        check_all()


def test_scope_df_one_million_rows(co, zb):
    n = 1_000_000
    with _engine(n) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=0, start_counter=0, n_rows=n)
        df, live = e.scope_df([s])
        F = co.synth_rows(0, 0, n, zb)
        w_df, w_n = co.bucket_df(F, np.zeros(n, np.uint32), [0])
        assert live == w_n == n and (df == w_df).all()


def test_weighted_text_search_matches_oracle(co, zb, rf1):
    """rf_search_text_w: the query is tokenised, hashed AND weighted on the GPU; ids / scores / cosines
    equal the oracle's RF-1w ranking, also inside a metadata range restriction."""
    n = 40_000
    with _engine(n) as e:
        a = e.open_store("fileSearchStores/a"); b = e.open_store("fileSearchStores/b")
        F, ff = co.synth_rows(41, 0, n, zb, with_ff=True)
        e.ingest_features(a, 1, F[:25_000]); e.ingest_features(b, 2, F[25_000:])
        seg = np.concatenate([np.full(25_000, a), np.full(15_000, b)]).astype(np.uint32)
        for scope in ([a], [b], [a, b]):
            w = e.scope_weights(scope)
            assert (w == co.idf_weights(*co.bucket_df(F, seg, scope))).all()
            for text in (b"1786 23 4479 313 12 7 1318 21", b"7 7 7 7 7 7 7 7 12 the 12", b"", b"zzz"):
                qw = co.weight_query(co.query_vector(text), w)
                ids, sc, cs, q_gpu = e.search_text(text, scope, 10, weights=w)
                assert (q_gpu == qw).all()
                assert (e.weight_query(co.query_vector(text), w) == qw).all()
                w_ids, w_sc, w_cs = co.score_topk(F, seg, qw, scope, k=10, ff=ff)
                assert ids.tolist() == w_ids.tolist() and sc.tolist() == w_sc.tolist()
                np.testing.assert_allclose(cs, w_cs, rtol=COS_RTOL, atol=0)
        w = e.scope_weights([a])
        text = b"1786 23 4479 313 12 7 1318 21"
        qw = co.weight_query(co.query_vector(text), w)
        ids, sc, _, _ = e.search_text(text, [a], 10, ranges=[(500, 9_000), (20_000, 30_000)], weights=w)
        keys = np.concatenate([co.score_topk_keys(F, seg, qw, [a], k=10, row_lo=lo, row_hi=hi) for lo, hi in [(500, 9_000), (20_000, 30_000)]])
        from rag_foundation_b200 import unpack_keys
        w_ids, w_sc, _ = unpack_keys(co.merge_topk(keys, 10))
        assert ids.tolist() == w_ids.tolist() and sc.tolist() == w_sc.tolist()


# ------------------------------------------------------------------ whole stores per rank (configs[4], multi-GPU by store)
def test_device_scoped_search_and_store_sharded_two_engines(co, zb):
    """rf_search_keys_device_scoped (one scope per query, device-resident) against the oracle, then
    two engines as two ranks of a StoreShardedSearcher (stores g % 2), merged with the CUDA merge
    kernel: equals the brute-force ranking over all stores, also for scopes that span both."""
    import torch
    from rag_foundation_b200.sharded import StoreShardedSearcher, unpack_keys_torch
    n_stores, per = 7, 3000
    world = 2
    engines, searchers, held = [], [], []
    try:
        for r in range(world):
            e = _engine(n_stores * per, id_base=StoreShardedSearcher.id_base_for(r, world)); engines.append(e)
            s = StoreShardedSearcher.for_engine(e)
            s.world, s.rank = world, r              # single process standing in for two ranks
            searchers.append(s); held.append([])
        for g in range(n_stores):
            for r, s in enumerate(searchers):
                assert s.open_store(f"fileSearchStores/s{g}") == g
                if s.owner(g) == r:
                    rows = co.synth_rows(14, g * per, per, zb)
                    engines[r].ingest_features(s.local_seg[g], g + 1, rows)
                    held[r].append((g, rows))
        scopes = [[0], [1], [6], [0, 1], [1, 2, 3, 4, 5], [3, 3], [], [2, 4, 6], [5]] * 3
        nq = len(scopes)
        Q = np.stack([co.synth_query(14, i, zb) for i in range(nq)])
        qd = torch.from_numpy(Q).cuda()
        local = []
        for r, s in enumerate(searchers):
            keys = s.local_search(qd, s.local_scopes(scopes), 10)
            torch.cuda.synchronize()
            # the rank's own answer == oracle over its rows with its local segments
            F = np.concatenate([rows for _, rows in held[r]])
            sg = np.concatenate([np.full(per, s.local_seg[g], np.uint32) for g, _ in held[r]])
            got = keys.cpu().numpy().view(np.uint64)
            for i, sc in enumerate(s.local_scopes(scopes)):
                want = co.score_topk_keys(F, sg, Q[i], sc, k=10, id_base=engines[r].id_base) if sc else np.zeros(10, np.uint64)
                assert got[i].tolist() == want.tolist(), (r, i)
            local.append(keys)
        merged = searchers[0].merge(torch.stack(local).contiguous(), 10)
        torch.cuda.synchronize()
        ids, sc, valid = unpack_keys_torch(merged.cpu())
        F = co.synth_rows(14, 0, n_stores * per, zb)
        store = np.repeat(np.arange(n_stores), per)
        gid = np.zeros(n_stores * per, np.int64)
        for g in range(n_stores):
            gid[g * per:(g + 1) * per] = StoreShardedSearcher.id_base_for(g % world, world) + (g // world) * per + np.arange(per)
        for i, scope in enumerate(scopes):
            s_all = F.astype(np.int64) @ Q[i].astype(np.int64)
            rows = np.nonzero(np.isin(store, scope))[0]
            order = rows[np.lexsort((gid[rows], -s_all[rows]))][:10]
            m = len(order)
            assert ids[i][:m].tolist() == gid[order].tolist() and sc[i][:m].tolist() == s_all[order].tolist(), scope
            assert not valid[i][m:].any()
    finally:
        for e in engines:
            e.close()


@pytest.mark.parametrize("k", [1, 2, 17, 32])
def test_other_k_values_single_batch_and_device_paths(co, zb, k):
    """k from 1 to RF_TOPK_MAX: host single query, host batch, device-resident keys and the merge kernel
    (k = 32 with the default grid takes the final merge's fallback path, the others the one-sweep path)."""
    import torch
    n = 150_000
    with _engine(n + 10, id_base=3) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=17, start_counter=0, n_rows=n)
        F, ff = co.synth_rows(17, 0, n, zb, with_ff=True)
        seg = np.full(n, s, np.uint32)
        Q = np.stack([co.synth_query(17, i, zb) for i in range(5)])
        ids, sc, cs, cnt = e.search(Q[:1], [[s]], k=k)
        _check(co, (ids[0], sc[0], cs[0], cnt[0]), F, seg, Q[0], [s], k, 3, ff)
        ids, sc, cs, cnt = e.search(Q, [[s]] * 5, k=k)
        for i in range(5):
            _check(co, (ids[i], sc[i], cs[i], cnt[i]), F, seg, Q[i], [s], k, 3, ff)
        keys = _device_batch(e, Q, [s], k)
        for i in range(5):
            assert keys[i].tolist() == co.score_topk_keys(F, seg, Q[i], [s], k=k, id_base=3).tolist()
        # merge kernel: three copies of the lists with disjoint halves zeroed must merge back
        parts = np.stack([np.where((np.arange(k)[None, :] % 3) == r, keys, 0) for r in range(3)]).astype(np.uint64)
        parts = -np.sort(-parts.view(np.int64), axis=2)          # each list sorted descending, zero padded (keys < 2^63)
        gathered = torch.from_numpy(np.ascontiguousarray(parts)).cuda()
        out = torch.zeros((5, k), dtype=torch.int64, device="cuda")
        e.merge_topk_device(gathered.data_ptr(), 3, 5, k, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert (out.cpu().numpy().view(np.uint64) == keys).all()


def test_small_batches_take_the_tensor_core_path_when_it_pays(co, zb):
    """A handful of queries over a large store: the cost model routes 3+ queries x 500 k rows (host and
    device-resident calls) to the batched kernel -- same keys, ids, scores and cosines as the oracle --
    and leaves 2 queries, or a small store, on the scan kernel."""
    n = 500_000
    with _engine(n + 50_000) as e:
        a = e.open_store("fileSearchStores/a"); b = e.open_store("fileSearchStores/b")
        e.ingest_synthetic(a, 0, seed=27, start_counter=0, n_rows=n)
        e.ingest_synthetic(b, 0, seed=28, start_counter=0, n_rows=40_000)
        F, ff = co.synth_rows(27, 0, n, zb, with_ff=True)
        Fb, ffb = co.synth_rows(28, 0, 40_000, zb, with_ff=True)
        Fall = np.concatenate([F, Fb]); ffall = np.concatenate([ff, ffb])
        seg = np.concatenate([np.full(n, a), np.full(40_000, b)]).astype(np.uint32)
        Q = np.stack([co.synth_query(27, i, zb) for i in range(63)])
        for nq, gemm in [(2, False), (3, True), (5, True), (8, True), (17, True), (63, True)]:
            l0 = e.stats()["kernel_launches"]
            keys = _device_batch(e, Q[:nq], [a], 10)
            assert (e.stats()["kernel_launches"] - l0 == 4) == gemm, nq
            for i in range(nq):
                assert keys[i].tolist() == co.score_topk_keys(Fall, seg, Q[i], [a], k=10).tolist(), (nq, i)
            l0 = e.stats()["kernel_launches"]
            ids, sc, cs, cnt = e.search(Q[:nq], [[a]] * nq, k=10)
            # the host call adds its own copies + unpack kernel (20 us in the model): 3 x 500 k stays on the scan kernel
            assert (e.stats()["kernel_launches"] - l0 == 5) == (gemm and nq >= 5), nq
            for i in range(nq):
                _check(co, (ids[i], sc[i], cs[i], cnt[i]), Fall, seg, Q[i], [a], 10, 0, ffall)
        l0 = e.stats()["kernel_launches"]
        keys = _device_batch(e, Q[:8], [b], 10)                 # 40 k rows: eight scans are cheaper than the fixed cost
        assert e.stats()["kernel_launches"] - l0 == 1
        for i in range(8):
            assert keys[i].tolist() == co.score_topk_keys(Fall, seg, Q[i], [b], k=10).tolist()


def test_interleaved_stores_batch_over_the_bounding_range(co, zb):
    """Two stores ingested alternately (every store has many extents): a same-scope batch is scored by
    the tensor-core kernel over the scope's bounding range, the other store's rows (half of the range)
    and a tombstoned document dropped by the per-row mask; a sparse scope stays on the scan kernel."""
    per, docs = 20_000, 12
    with _engine(per * docs + 1000) as e:
        a = e.open_store("fileSearchStores/a"); b = e.open_store("fileSearchStores/b"); c = e.open_store("fileSearchStores/c")
        parts, segs = [], []
        for d in range(docs):
            rows = co.synth_rows(29, d * per, per, zb)
            st = a if d % 2 == 0 else b
            e.ingest_features(st, d + 1, rows)
            parts.append(rows); segs.append(np.full(per, st, np.uint32))
        e.ingest_features(c, 99, parts[0][:1000])
        e.tombstone_doc(5)                                      # one of store a's documents
        segs[4][:] = 0xFFFFFFFF
        F = np.concatenate(parts + [parts[0][:1000]]); seg = np.concatenate(segs + [np.full(1000, c, np.uint32)])
        Q = np.stack([co.synth_query(29, i, zb) for i in range(32)])
        for scope, gemm in (([a], True), ([a, b], True), ([c], False)):
            l0 = e.stats()["kernel_launches"]
            keys = _device_batch(e, Q, scope, 10)
            assert (e.stats()["kernel_launches"] - l0 == 4) == gemm, scope
            for i in range(0, 32, 3):
                assert keys[i].tolist() == co.score_topk_keys(F, seg, Q[i], scope, k=10).tolist(), (scope, i)


@pytest.mark.timeout(300)
def test_searches_stay_consistent_while_ingest_and_deletes_run(co, zb):
    """Readers (single queries, text queries, weighted queries, batches) on 6 threads while a writer
    appends documents to a second store, tombstones some and drops a store.  What a reader sees depends
    on timing, so the check is by invariants: every hit belongs to a scoped store at read time or
    earlier (never the other tenant's), ids ascend within equal scores, every score is the score of a row
    that some document really held (rows of deleted documents are reused, so a row re-read after the
    search may already belong to a later document: hits in the static store are re-scored exactly, hits in
    the churning stores must match one of the rows the writer ever ingests), and the store nobody writes
    to always returns its oracle answer."""
    import threading
    n = 120_000
    with _engine(n + 200_000, n_contexts=6) as e:
        fixed = e.open_store("fileSearchStores/fixed"); live = e.open_store("fileSearchStores/live")
        gone = e.open_store("fileSearchStores/gone")
        e.ingest_synthetic(fixed, 0, seed=31, start_counter=0, n_rows=n)
        F = co.synth_rows(31, 0, n, zb)
        Q = np.stack([co.synth_query(31, i, zb) for i in range(8)])
        want = [co.score_topk(F, np.full(n, fixed, np.uint32), Q[i], [fixed])[0].tolist() for i in range(8)]
        doc_rows = co.synth_rows(32, 0, 4000, zb)
        possible = [set((doc_rows.astype(np.int32) @ Q[j].astype(np.int32)).tolist()) for j in range(4)]
        e.ingest_features(gone, 500, doc_rows)
        stop = threading.Event()
        errs = []
        epoch = [0]          # bumped by the writer around every change: a reader's re-read of the rows it was given is only
        checked = [0]        # meaningful when nothing changed in between (freed rows are handed to later documents)

        def writer():
            import time
            try:
                for d in range(40):
                    epoch[0] += 1
                    e.ingest_features(live, 1000 + d, doc_rows[(d % 4) * 1000:(d % 4) * 1000 + 1000])
                    if d % 5 == 4:
                        e.tombstone_doc(1000 + d - 2)
                    if d == 20:
                        e.drop_store(gone)
                    epoch[0] += 1
                    time.sleep(0.004)
            except Exception as ex:   # noqa: BLE001
                errs.append(ex)
            finally:
                stop.set()

        def reader(i):
            try:
                it = 0
                while not stop.is_set() or it < 3:
                    it += 1
                    ids, sc, cs, cnt = e.search(Q[i][None, :], [[fixed]], k=10)
                    assert ids[0].tolist() == want[i]
                    ep0 = epoch[0]
                    ids, sc, cs, cnt = e.search(Q[:4], [[live], [fixed, live], [gone], [live, gone]], k=10)
                    for j in range(4):
                        m = int(cnt[j])
                        got_ids, got_sc = ids[j][:m].astype(np.int64), sc[j][:m].astype(np.int64)
                        assert all((got_sc[x] > got_sc[x + 1]) or (got_sc[x] == got_sc[x + 1] and got_ids[x] < got_ids[x + 1]) for x in range(m - 1))
                        static = got_ids < n
                        assert all(int(x) in possible[j] for x in got_sc[~static])
                        if m:
                            got = [e.read_rows(int(g), 1) for g in got_ids]
                            if ep0 % 2 == 0 and epoch[0] == ep0:     # the index did not change between the search and the re-read
                                rows = np.concatenate([g[0] for g in got]); sg = np.concatenate([g[1] for g in got])
                                assert (rows.astype(np.int32) @ Q[j].astype(np.int32) == got_sc).all()
                                allowed = {0: {live}, 1: {fixed, live}, 2: {gone}, 3: {live, gone}}[j]
                                assert set(sg.tolist()) <= allowed, (j, set(sg.tolist()))
                                checked[0] += 1
                    w = e.scope_weights([fixed, live])
                    ids2, _, _, _ = e.search_text(b"1786 23 4479 313 12 7 1318 21", [fixed, live], 10, weights=w)
                    assert len(ids2) == 10
            except Exception as ex:   # noqa: BLE001
                errs.append(ex)

        ts = [threading.Thread(target=writer)] + [threading.Thread(target=reader, args=(i,)) for i in range(5)]
        [t.start() for t in ts]; [t.join() for t in ts]
        assert not errs, errs[:2]
        assert checked[0] >= 20, f"only {checked[0]} re-reads fell between two writes"
        assert e.lookup_store("fileSearchStores/gone") is None
