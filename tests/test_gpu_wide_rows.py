"""Wider feature rows (D = 512, 1024; SURVEY.md 8f-4, oracle/SPEC.md "Wider rows") through the C-ABI: the same
parity bar as the 256-feature suite -- bit-exact ids / int32 scores / int8 rows / spans against the CPU oracle,
cosine within 1e-5 relative -- for the generator, ingest featurisation, the scan kernel (single query, batches,
store-scoped batches, tenant mask, tombstones, ragged tiles), RF-1w, snapshots, the engine group and the
adapter.  Run on the GPU box: pytest -m gpu."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

COS_RTOL = 1e-5
DIMS = (512, 1024)


@pytest.fixture(scope="module")
def co():
    from oracle import c_oracle
    return c_oracle


@pytest.fixture(scope="module")
def zbs():
    from oracle import rf1
    return {d: rf1.zipf_bucket_table(dim=d) for d in DIMS}


def _engine(cap, dim, **kw):
    from rag_foundation_b200 import Engine
    return Engine(capacity_rows=cap, dim=dim, **kw)


def _check(co, eng_out, F, seg, q, scope, k, id_base, ff):
    ids, sc, cs, cnt = eng_out
    w_ids, w_sc, w_cs = co.score_topk(F, seg, q, scope, k=k, id_base=id_base, ff=ff)
    m = len(w_ids)
    assert int(cnt) == m
    assert ids[:m].tolist() == w_ids.tolist()
    assert sc[:m].tolist() == w_sc.tolist()
    np.testing.assert_allclose(cs[:m], w_cs, rtol=COS_RTOL, atol=0)
    assert (ids[m:] == np.uint64(0xFFFFFFFFFFFFFFFF)).all() and (sc[m:] == 0).all()


@pytest.mark.parametrize("dim", DIMS)
def test_synthetic_rows_and_golden_top10(co, zbs, dim, golden_dir):
    g = json.load(open(os.path.join(golden_dir, "rf1_wide_golden.json")))[str(dim)]["synth_top10"]
    zb = zbs[dim]
    with _engine(70_000, dim) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=5, start_counter=123_456, n_rows=65_537)
        F, seg, ff = e.read_rows(0, 65_537)
        wF, wff = co.synth_rows(5, 123_456, 65_537, zb, with_ff=True, dim=dim)
        assert F.shape == (65_537, dim) and (F == wF).all() and (ff == wff).all() and (seg == s).all()
    # the frozen vectors (Python oracle): stores 0 / 1 interleaved, tombstones, id_base 50
    with _engine(g["n_rows"], dim, id_base=g["id_base"]) as e:
        s0 = e.open_store("fileSearchStores/a"); s1 = e.open_store("fileSearchStores/b")
        F = co.synth_rows(g["seed"], g["start"], g["n_rows"], zb, dim=dim)
        seg = np.zeros(g["n_rows"], np.uint32)
        seg[1::5] = 1
        tomb = np.zeros(g["n_rows"], bool)
        tomb[7::11] = True
        doc = 0
        for r in range(g["n_rows"]):      # one document per row: every row can be deleted on its own
            doc += 1
            e.ingest_features(s1 if seg[r] == 1 else s0, doc, F[r:r + 1])
        for r in np.nonzero(tomb)[0]:
            e.tombstone_doc(int(r) + 1)
        for case in g["cases"]:
            q = co.synth_query(g["seed"], case["qi"], zb, dim=dim)
            ids, sc, _, cnt = e.search(q[None, :], [[s0] if case["scope"] == [0] else [s0, s1]], k=10)
            assert ids[0][:cnt[0]].tolist() == case["ids"] and sc[0][:cnt[0]].tolist() == case["scores"]


@pytest.mark.parametrize("dim", DIMS)
@pytest.mark.parametrize("n_rows", [1, 7, 8, 9, 15, 16, 17, 1000, 20_001, 262_144 + 5])
def test_single_query_parity_sizes(co, zbs, dim, n_rows):
    """Row counts around the tile height (32 sub-rows = 16 / 8 rows) and a multi-block corpus."""
    zb = zbs[dim]
    with _engine(n_rows + 64, dim, id_base=1000) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=1, start_counter=0, n_rows=n_rows)
        F, ff = co.synth_rows(1, 0, n_rows, zb, with_ff=True, dim=dim)
        seg = np.full(n_rows, s, np.uint32)
        for qi in range(3):
            q = co.synth_query(1, qi, zb, dim=dim)
            for k in (1, 10, 32):
                ids, sc, cs, cnt = e.search(q[None, :], [[s]], k=k)
                _check(co, (ids[0], sc[0], cs[0], cnt[0]), F, seg, q, [s], k, 1000, ff)


@pytest.mark.parametrize("dim", DIMS)
def test_every_feature_position_counts(co, dim):
    """One-hot rows and queries: feature d of the query must meet feature d of the row, for every d --
    catches a lane reading the wrong 256-feature part of its row or query."""
    n = dim
    F = np.zeros((n, dim), np.int8)
    F[np.arange(n), np.arange(n)] = np.arange(n) % 100 + 1
    with _engine(n + 8, dim) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_features(s, 1, F)
        ff = (F.astype(np.int32) ** 2).sum(1).astype(np.int32)
        seg = np.full(n, s, np.uint32)
        rng = np.random.default_rng(0)
        for d in list(range(0, dim, 37)) + [255, 256, dim - 1]:
            q = np.zeros(dim, np.int8)
            q[d] = 127                      # 127 * F[d, d] >= 127 > F[o, o] <= 100: row d wins
            q[(d * 7 + 11) % dim] = 1
            ids, sc, cs, cnt = e.search(q[None, :], [[s]], k=10)
            assert ids[0][0] == d and int(sc[0][0]) == 127 * int(F[d, d])
            _check(co, (ids[0], sc[0], cs[0], cnt[0]), F, seg, q, [s], 10, 0, ff)
        q = rng.integers(0, 5, dim).astype(np.int8)
        ids, sc, cs, cnt = e.search(q[None, :], [[s]], k=32)
        _check(co, (ids[0], sc[0], cs[0], cnt[0]), F, seg, q, [s], 32, 0, ff)


@pytest.mark.parametrize("dim", DIMS)
def test_store_mask_tombstones_and_batches(co, zbs, dim):
    zb = zbs[dim]
    rng = np.random.default_rng(3)
    with _engine(40_000, dim) as e:
        stores = [e.open_store(f"fileSearchStores/t{i}") for i in range(5)]
        F_all, seg_all = [], []
        start = 0
        for d in range(100):      # many small documents, stores interleaved -> many extents per store, ragged tiles
            s = int(rng.integers(0, 5))
            n = int(rng.integers(1, 300))
            rows = co.synth_rows(9, start, n, zb, dim=dim)
            e.ingest_features(stores[s], d + 1, rows)
            F_all.append(rows); seg_all.append(np.full(n, stores[s], np.uint32))
            start += n
        F = np.concatenate(F_all); seg = np.concatenate(seg_all)
        ff = (F.astype(np.int32) ** 2).sum(1).astype(np.int32)
        scopes = [[stores[0]], [stores[1], stores[3]], stores, [stores[4], stores[4]], [77], []]
        for qi in range(2):
            q = co.synth_query(9, qi, zb, dim=dim)
            ids, sc, cs, cnt = e.search(np.stack([q] * len(scopes)), scopes, k=10)     # per-query scopes (store table)
            for i, scope in enumerate(scopes):
                _check(co, (ids[i], sc[i], cs[i], cnt[i]), F, seg, q, scope, 10, 0, ff)
        # a same-scope batch (shared plan: nq scans in one launch, or the K = 1024 tensor-core kernel when it pays)
        Q = np.stack([co.synth_query(9, 10 + i, zb, dim=dim) for i in range(70)])
        ids, sc, cs, cnt = e.search(Q, [stores] * 70, k=10)
        for i in range(0, 70, 9):
            _check(co, (ids[i], sc[i], cs[i], cnt[i]), F, seg, Q[i], stores, 10, 0, ff)
        for d in (3, 40, 77):
            e.tombstone_doc(d)
        lo = 0
        for d, rows in enumerate(F_all, start=1):
            if d in (3, 40, 77):
                seg[lo:lo + len(rows)] = 0xFFFFFFFF
            lo += len(rows)
        e.drop_store(stores[1])
        seg[seg == stores[1]] = 0xFFFFFFFF
        q = co.synth_query(9, 0, zb, dim=dim)
        for scope in ([stores[0]], [stores[1]], stores):
            ids, sc, cs, cnt = e.search(q[None, :], [scope], k=10)
            _check(co, (ids[0], sc[0], cs[0], cnt[0]), F, seg, q, scope, 10, 0, ff)
        # RF-1w statistics over what is left
        df, n_live = e.scope_df(stores)
        w_df, w_n = co.bucket_df(F, seg, stores)
        assert n_live == w_n and df.shape == (dim,) and (df == w_df).all()
        w = e.idf_weights(df, n_live)
        assert (w == co.idf_weights(w_df, w_n)).all()


@pytest.mark.parametrize("dim", DIMS)
def test_ingest_text_and_text_search(co, dim, golden_dir):
    base = json.load(open(os.path.join(golden_dir, "rf1_golden.json")))
    rng = np.random.default_rng(dim)
    words = [bytes(rng.integers(97, 123, rng.integers(1, 10)).astype(np.uint8)) for _ in range(3000)]
    docs = [base["sample_report"]["text"].encode("utf-8"), base["long_doc"]["text"].encode(),
            b"", b"the a an", b"x" * 5000 + b" tail",
            b" ".join(words[i] for i in rng.integers(0, 3000, 200_000))]            # ~1.1 MB: the chunked copy pipeline
    with _engine(40_000, dim) as e:
        s = e.open_store("fileSearchStores/a")
        F_all, ff_all = [], []
        for d, text in enumerate(docs):
            first, nc, spans = e.ingest_text(s, d + 1, text)
            wF, wff, wsp, _ = co.featurize_doc(text, dim)
            assert nc == len(wF) and (spans == wsp).all()
            if nc:
                F, seg, ff = e.read_rows(first, nc)
                assert (F == wF).all() and (ff == wff).all() and (seg == s).all()
            F_all.append(wF); ff_all.append(wff)
        F = np.concatenate(F_all); ff = np.concatenate(ff_all)
        seg = np.full(len(F), s, np.uint32)
        for text in (base["sample_report"]["query"].encode(), b"tail x", words[5] + b" " + words[77] + b" THE " + words[5]):
            ids, sc, cs, q = e.search_text(text, [s], 10)
            wq = co.query_vector(text, dim)
            assert (q == wq).all() and (e.featurize_query(text) == wq).all()
            w_ids, w_sc, w_cs = co.score_topk(F, seg, wq, [s], ff=ff)
            assert ids.tolist() == w_ids.tolist() and sc.tolist() == w_sc.tolist()
            np.testing.assert_allclose(cs, w_cs, rtol=COS_RTOL)
            # RF-1w: the query weighted on the GPU
            w = e.scope_weights([s])
            ids, sc, cs, qw = e.search_text(text, [s], 10, weights=w)
            want_q = co.weight_query(wq, w)
            assert (qw == want_q).all()
            w_ids, w_sc, _ = co.score_topk(F, seg, want_q, [s], ff=ff)
            assert ids.tolist() == w_ids.tolist() and sc.tolist() == w_sc.tolist()


@pytest.mark.parametrize("dim", DIMS)
def test_snapshot_round_trip_and_width_mismatch(tmp_path, co, zbs, dim):
    zb = zbs[dim]
    path = str(tmp_path / "idx.rfsnap")
    n = 5000
    with _engine(n + 100, dim) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=2, start_counter=0, n_rows=n)
        e.ingest_text(s, 9, b"one more document about snapshots")
        q = co.synth_query(2, 1, zb, dim=dim)
        want = e.search(q[None, :], [[s]], k=10)
        e.save_snapshot(path)
    with _engine(n + 100, dim) as e2:
        e2.load_snapshot(path)
        s2 = e2.lookup_store("fileSearchStores/a")
        got = e2.search(q[None, :], [[s2]], k=10)
        for a, b in zip(want, got):
            assert (a == b).all()
    with _engine(n + 100, 256) as e3:
        with pytest.raises(RuntimeError, match="features"):
            e3.load_snapshot(path)


@pytest.mark.parametrize("dim", DIMS)
def test_group_of_two_engines_equals_single_engine(co, zbs, dim):
    from rag_foundation_b200 import EngineGroup
    zb = zbs[dim]
    n = 40_000
    with _engine(n, dim) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=4, start_counter=0, n_rows=n)
        Q = np.stack([co.synth_query(4, i, zb, dim=dim) for i in range(5)])
        want = e.search(Q, [[s]] * 5, k=10)
        w_df = e.scope_df([s])
    with EngineGroup([0, 0], capacity_rows=n // 2, placement="spread", id_bases=[0, n // 2], dim=dim) as g:
        gs = g.open_store("fileSearchStores/a")
        g.ingest_synthetic(gs, 0, seed=4, start_counter=0, n_rows=n)
        got = g.search(Q, [[gs]] * 5, k=10)
        for a, b in zip(want, got):
            assert (a == b).all()
        df, nn = g.scope_df([gs])
        assert nn == w_df[1] and (df == w_df[0]).all()
        ids, sc, cs, q = g.search_text(b"17 4242 9", [gs], 10)
        wq = co.query_vector(b"17 4242 9", dim)
        assert (q == wq).all()


def _device_batch(e, Q, scope, k):
    import torch
    qd = torch.from_numpy(np.ascontiguousarray(Q)).cuda()
    out = torch.zeros((Q.shape[0], k), dtype=torch.int64, device="cuda")
    e.search_keys_device(qd.data_ptr(), Q.shape[0], scope, k, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("dim", [512, 1024])
@pytest.mark.parametrize("n_rows,nq,k", [(100_001, 300, 10), (65_536, 128, 10), (200_000, 1024, 10), (150_000, 513, 3), (70_000, 17, 1)])
def test_wide_gemm_path_parity(co, zbs, n_rows, nq, k, dim):
    """D = 512 / 1024 batches on the tensor cores (streamed-K pair kernel: K in two / four slabs, accumulators alternating
    by tile) == oracle: ragged query groups, a ragged last chunk tile, k < 10."""
    zb = zbs[dim]
    with _engine(n_rows + 200, dim, id_base=7) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=21, start_counter=0, n_rows=n_rows)
        l0 = e.stats()["kernel_launches"]
        F = co.synth_rows(21, 0, n_rows, zb, dim=dim)
        seg = np.full(n_rows, s, np.uint32)
        Q = np.stack([co.synth_query(21, i, zb, dim=dim) for i in range(nq)])
        keys = _device_batch(e, Q, [s], k)
        assert e.stats()["kernel_launches"] - l0 == 4, "the batch should have taken the tensor-core path (floor pass, k-th largest, main pass, merge)"
        for i in range(0, nq, max(1, nq // 40)):
            assert keys[i].tolist() == co.score_topk_keys(F, seg, Q[i], [s], k=k, id_base=7).tolist(), i
        # the same batch through the host entry point
        ids, sc, cs, cnt = e.search(Q, [[s]] * nq, k=k)
        got = (sc.astype(np.int64).astype(np.uint64) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - ids)
        assert (got == keys).all()


@pytest.mark.parametrize("dim", [512, 1024])
def test_wide_gemm_masks_saturation_and_heavy_queries(co, zbs, dim):
    """Streamed-K kernel under the tenant mask (two stores interleaved in the scored range, tombstones), with saturated
    features (127) and queries that hit every K-slab."""
    zb = zbs[dim]
    rng = np.random.default_rng(11)
    with _engine(90_000, dim) as e:
        a = e.open_store("fileSearchStores/a"); b = e.open_store("fileSearchStores/b")
        parts, segs, start = [], [], 0
        for d in range(8):
            n = 9000 + 113 * d
            rows = co.synth_rows(33, start, n, zb, dim=dim)
            if d == 2:
                rows[::7, ::5] = 127                      # saturated features in every K-slab
            st = a if d % 3 else b
            e.ingest_features(st, d + 1, rows)
            parts.append(rows); segs.append(np.full(n, st, np.uint32)); start += n
        e.tombstone_doc(5)
        segs[4][:] = 0xFFFFFFFF
        F = np.concatenate(parts); seg = np.concatenate(segs)
        Q = np.stack([co.synth_query(33, i, zb, dim=dim) for i in range(40)])
        Q[3] = rng.integers(0, 4, dim).astype(np.int8)    # dense query: every feature position counts
        Q[4] = 127
        Q[5] = 0
        for scope in ([a], [a, b]):
            keys = _device_batch(e, Q, scope, 10)
            for i in range(0, 40, 3):
                assert keys[i].tolist() == co.score_topk_keys(F, seg, Q[i], scope, k=10).tolist(), (scope, i)
            for i in (3, 4, 5):
                assert keys[i].tolist() == co.score_topk_keys(F, seg, Q[i], scope, k=10).tolist(), (scope, i)


def test_one_million_wide_chunks_parity(co, zbs):
    """The D = 1024 scan at a streaming size (1 M chunks = 1.03 GB per query), 3 queries against the oracle."""
    dim, n = 1024, 1_000_000
    zb = zbs[dim]
    with _engine(n, dim) as e:
        s = e.open_store("fileSearchStores/a")
        e.ingest_synthetic(s, 0, seed=0, start_counter=0, n_rows=n)
        F, ff = co.synth_rows(0, 0, n, zb, with_ff=True, dim=dim)
        seg = np.full(n, s, np.uint32)
        for qi in range(3):
            q = co.synth_query(0, qi, zb, dim=dim)
            ids, sc, cs, cnt = e.search(q[None, :], [[s]], k=10)
            _check(co, (ids[0], sc[0], cs[0], cnt[0]), F, seg, q, [s], 10, 0, ff)


def test_relevance_at_1024_features_equals_the_numpy_model(tmp_path):
    """graded with the reference harness's citation_hit (scripts/benchmark/metrics.py:73-92,
    restated in tools/quality_eval.py), the D = 1024 engine behind the adapter must give exactly the hit rates of
    an independent numpy model of RF-1 / RF-1w at D = 1024, and IDF scoring must reach 20 %."""
    import importlib.util
    from rag_foundation_b200 import Engine
    from rag_foundation_b200 import adapter as ad
    spec = importlib.util.spec_from_file_location("quality_eval", os.path.join(os.path.dirname(__file__), "..", "tools", "quality_eval.py"))
    qe = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(qe)
    docs, questions = qe.make_labelled_set(600, 150, seed=2)
    reg = ad.Registry(Engine(capacity_rows=8192, dim=1024))
    try:
        res = qe.run_eval(lambda scoring: ad.B200Rag(registry=reg, scoring=scoring), docs, questions)
        model = qe.model_dim_sweep(docs, questions, dims=(256, 1024))
        for mode in ("tf", "idf"):
            assert res["scoring"][mode]["citation_hit_rate"] == pytest.approx(model["1024"][mode], abs=1e-12), (res, model)
        assert res["scoring"]["idf"]["citation_hit_rate"] >= 0.20
        assert model["1024"]["idf"] > model["256"]["idf"]          # the point of the wider rows
    finally:
        reg.engine.close()


def test_bad_width_is_refused():
    from rag_foundation_b200 import Engine
    for dim in (0, 128, 300, 2048):
        with pytest.raises(RuntimeError, match="dim"):
            Engine(capacity_rows=64, dim=dim)
