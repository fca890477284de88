"""The CPU oracle (oracle/rf1.py and the independent C restatement oracle/rf1_oracle.c) against
 - fixtures generated from the UNMODIFIED reference (tests/golden/make_reference_golden.py),
 - published FNV-1a vectors,
 - the frozen RF-1 known-answer vectors (tests/golden/make_rf1_golden.py),
and against each other on random inputs.  No GPU."""
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import c_oracle as co
from oracle import rf1


def _dense(sparse, n=256):
    row = np.zeros(n, np.int8)
    for i, v in sparse:
        row[i] = v
    return row


@pytest.fixture(scope="module")
def rf1_golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, "rf1_golden.json")))


@pytest.fixture(scope="module")
def zb():
    return rf1.zipf_bucket_table()


# ---- pinned against the reference -------------------------------------------------------------
def test_tokeniser_matches_reference_normalize(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "normalize_golden.json")))
    assert len(cases) >= 16
    for c in cases:
        text = c["text"]
        if not text.isascii():
            # _normalize lowercases with str.lower(); RF-1 is byte level.  They agree unless a
            # non-ASCII code point lowercases to ASCII (SPEC.md step 1) -- none in these cases.
            assert all(not ch.lower().isascii() or ch.isascii() for ch in text)
        assert rf1.normalize(text) == c["normalized"], text[:40]
        st_, en, _bk = co.tokenize(text.encode("utf-8"))
        toks = [text.encode("utf-8")[a:b].decode("ascii").lower() for a, b in zip(st_, en)]
        assert " ".join(toks) == c["normalized"], text[:40]


def test_fnv1a_published_vectors():
    # FNV-1a 32 test vectors from the FNV reference distribution (test_fnv.c)
    for data, want in [(b"", 0x811C9DC5), (b"a", 0xE40C292C), (b"b", 0xE70C2DE5), (b"foobar", 0xBF9CF968)]:
        assert rf1.fnv1a32(data) == want
        assert co.fnv1a32(data) == want


def test_adapter_shapes_match_reference_wire_golden(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "config1_wire.json")))
    for case in g["contents_to_text"]:
        assert rf1.contents_to_text(case["contents"]) == case["text"]
    ref_cit = g["citations"][0]
    resp = rf1.final_response([{"uri": ref_cit["uri"], "title": ref_cit["title"], "text": ref_cit["snippet"],
                                "file_search_store": ref_cit["store"]}])
    assert resp.text is None and len(resp.candidates) == g["chunk1"]["n_candidates"]
    cits = rf1.extract_citations(resp)
    assert cits == g["citations"]
    frames = [json.loads(f[len("data: "):]) for f in rf1.citation_frames(cits)]
    assert frames == g["source_document_frames"]
    assert sorted(cits[0].keys()) == g["citation_keys"]


# ---- frozen RF-1 vectors ----------------------------------------------------------------------
def test_fnv_and_mix64_golden(rf1_golden):
    for tok, h in rf1_golden["fnv1a32"].items():
        assert rf1.fnv1a32(tok.encode()) == h and co.fnv1a32(tok.encode()) == h
    for s, a, b, want in rf1_golden["mix64"]:
        assert int(rf1.mix64(s, a, b)) == want
        assert int(co.lib().rf1_mix64(s, a, b)) == want


def test_chunk_count_rule(rf1_golden):
    for t, n in rf1_golden["n_chunks_for"].items():
        assert rf1.n_chunks_for(int(t)) == n
        assert int(co.lib().rf1_n_chunks(int(t))) == n


@pytest.mark.parametrize("which", ["sample_report", "long_doc"])
def test_document_featurisation_golden(rf1_golden, which):
    g = rf1_golden[which]
    data = g["text"].encode("utf-8")
    want_F = np.stack([_dense(r) for r in g["rows_sparse"]])
    for impl in (rf1, co):
        F, ff, spans, ntok = impl.featurize_doc(data)
        assert ntok == g["n_tokens"] and F.shape[0] == g["n_chunks"]
        assert (F == want_F).all()
        assert ff.tolist() == g["ff"]
        assert spans.tolist() == g["spans"]


def test_demo_query_golden(rf1_golden):
    g = rf1_golden["sample_report"]
    q_want = _dense(g["q_sparse"])
    F = np.stack([_dense(r) for r in g["rows_sparse"]])
    for impl in (rf1, co):
        q = impl.query_vector(g["query"].encode())
        assert (q == q_want).all()
    assert rf1.scores(F, q_want).tolist() == g["scores"]
    ids, sc, cs = co.score_topk(F, np.zeros(len(F), np.uint32), q_want, [0], ff=np.array(g["ff"], np.int32))
    assert sc.tolist() == g["scores"]
    np.testing.assert_allclose(cs, np.array(g["cos"], np.float32), rtol=1e-6)
    np.testing.assert_allclose(rf1.cosine(g["scores"], g["qq"], g["ff"]), np.array(g["cos"], np.float32), rtol=1e-6)


def test_saturation(rf1_golden):
    g = rf1_golden["saturation"]
    for impl in (rf1, co):
        q = impl.query_vector(g["text"].encode())
        assert (q == _dense(g["q_sparse"])).all() and q.max() == 127


def test_tiny_rank_cases(rf1_golden):
    g = rf1_golden["tiny"]
    F = np.stack([_dense(r) for r in g["F_sparse"]])
    q = _dense(g["q_sparse"])
    seg = np.array(g["store_seg"], np.uint32)
    for case in g["cases"]:
        ids, sc = rf1.score_topk(F, seg, q, case["scope"], k=case["k"], id_base=case["id_base"])
        assert ids.tolist() == case["ids"] and sc.tolist() == case["scores"], case["name"]
        ids, sc, _ = co.score_topk(F, seg, q, case["scope"], k=case["k"], id_base=case["id_base"])
        assert ids.tolist() == case["ids"] and sc.tolist() == case["scores"], case["name"]


def test_synthetic_generator_golden(rf1_golden, zb):
    g = rf1_golden["synth"]
    assert zb[:64].tolist() == rf1_golden["zipf_bucket_sha_first64"]
    want = np.stack([_dense(r) for r in g["rows_sparse_0_3"]])
    for impl in (rf1, co):
        assert (impl.synth_rows(0, 0, 4, zb) == want).all()
        assert (impl.synth_rows(0, 99_999_999, 1, zb)[0] == _dense(g["rows_at_99999999"][0])).all()
        for i, qs in enumerate(g["queries_sparse_0_2"]):
            assert (impl.synth_query(0, i, zb) == _dense(qs)).all()


def test_synthetic_top10_golden(rf1_golden, zb):
    for case in rf1_golden["synth_top10"]:
        F = co.synth_rows(case["seed"], 0, case["n_rows"], zb)
        q = co.synth_query(case["seed"], case["qi"], zb)
        ids, sc, _ = co.score_topk(F, np.zeros(len(F), np.uint32), q, [0])
        assert ids.tolist() == case["ids"] and sc.tolist() == case["scores"]


# ---- Python oracle == C oracle on random inputs -----------------------------------------------
_alphabet = st.sampled_from(list("abcdeTHEthe an AN a A x9 \n\t.,-_/é中") + ["the ", " a ", "An "])


@settings(max_examples=150, deadline=None)
@given(st.lists(_alphabet, max_size=400).map("".join))
def test_featurize_python_equals_c(text):
    data = text.encode("utf-8")
    F1, ff1, sp1, n1 = rf1.featurize_doc(data)
    F2, ff2, sp2, n2 = co.featurize_doc(data)
    assert n1 == n2 and F1.shape == F2.shape
    assert (F1 == F2).all() and (ff1 == ff2).all() and (sp1 == sp2).all()
    assert (rf1.query_vector(data) == co.query_vector(data)).all()


def test_long_random_document_python_equals_c():
    rng = np.random.default_rng(7)
    words = ["w%d" % i for i in range(300)] + ["the", "a", "an", "The", "AN"]
    text = " ".join(rng.choice(words, 5000)).encode()
    F1, ff1, sp1, n1 = rf1.featurize_doc(text)
    F2, ff2, sp2, n2 = co.featurize_doc(text)
    assert n1 == n2 and n1 > 3000 and F1.shape[0] == rf1.n_chunks_for(n1)
    assert (F1 == F2).all() and (ff1 == ff2).all() and (sp1 == sp2).all()


@pytest.mark.parametrize("seed", [0, 3])
def test_score_topk_python_equals_c(zb, seed):
    rng = np.random.default_rng(seed)
    n = 5000
    F = co.synth_rows(seed, 10_000, n, zb)
    assert (F == rf1.synth_rows(seed, 10_000, n, zb)).all()
    seg = rng.integers(0, 4, n).astype(np.uint32)
    seg[rng.integers(0, n, 50)] = 0xFFFFFFFF
    for qi in range(3):
        q = co.synth_query(seed, qi, zb)
        for scope in ([0], [1, 3], [0, 1, 2, 3], [9]):
            a_ids, a_sc = rf1.score_topk(F, seg, q, scope, k=10, id_base=77)
            b_ids, b_sc, _ = co.score_topk(F, seg, q, scope, k=10, id_base=77)
            assert a_ids.tolist() == b_ids.tolist() and a_sc.tolist() == b_sc.tolist()
            for threads in (1, 3):
                c_ids, c_sc, _ = co.score_topk(F, seg, q, scope, k=10, id_base=77, threads=threads)
                assert c_ids.tolist() == a_ids.tolist()


def test_merge_topk_equals_global(zb):
    F = co.synth_rows(1, 0, 4000, zb)
    seg = np.zeros(4000, np.uint32)
    q = co.synth_query(1, 0, zb)
    whole = co.score_topk_keys(F, seg, q, [0])
    parts = [co.score_topk_keys(F, seg, q, [0], row_lo=lo, row_hi=hi) for lo, hi in [(0, 1000), (1000, 1001), (1001, 4000)]]
    assert co.merge_topk(np.concatenate(parts)).tolist() == whole.tolist()
    assert rf1.merge_topk(parts) == [int(x) for x in whole]


# ---- RF-1w (IDF-weighted variant) ----------------------------------------------------------------
@pytest.fixture(scope="module")
def rf1w_golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, "rf1w_golden.json")))


def test_rf1w_weight_cases(rf1w_golden):
    for n, df, want in rf1w_golden["weight_cases"]:
        assert int(rf1.idf_weights([df], n)[0]) == want
        assert int(co.idf_weights(np.full(256, df, np.uint64), n)[0]) == want
    assert rf1w_golden["saturation"] == [0, 31, 124, 127, 127]
    sat = co.weight_query(np.array([0, 1, 4, 5, 127] + [0] * 251, np.int8), np.full(256, 31, np.uint8))
    assert sat[:5].tolist() == [0, 31, 124, 127, 127]


def test_rf1w_synthetic_golden_both_oracles(rf1w_golden, zb):
    g = rf1w_golden["synthetic"]
    F = co.synth_rows(g["seed"], 0, g["rows"], zb)
    seg = np.zeros(g["rows"], np.uint32)
    seg[g["store1"][0]:g["store1"][1]] = 1
    seg[g["tombstones"]] = rf1.TOMBSTONE
    for case in g["cases"]:
        for mod in (rf1, co):
            df, n = mod.bucket_df(F, seg, case["scope"])
            assert n == case["n"] and df.tolist() == case["df"]
            w = mod.idf_weights(df, n)
            assert w.tolist() == case["w"]
        for qc in case["queries"]:
            qw = co.weight_query(co.synth_query(g["seed"], qc["qi"], zb), w)
            assert (qw == _dense(qc["qw_sparse"])).all()
            ids, sc, _ = co.score_topk(F, seg, qw, case["scope"])
            assert ids.tolist() == qc["ids"] and sc.tolist() == qc["scores"]


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 2 ** 40), st.integers(0, 2 ** 40))
def test_rf1w_weights_python_equals_c_and_are_monotone(n, df):
    df = min(df, n)
    a = int(rf1.idf_weights([df], n)[0])
    assert a == int(co.idf_weights(np.full(256, df, np.uint64), n)[0]) and 4 <= a <= 31
    assert int(rf1.idf_weights([df // 2], n)[0]) >= a      # rarer bucket, weight never smaller
