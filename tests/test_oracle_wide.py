"""The wider RF-1 rows (D = 512, 1024; oracle/SPEC.md "Wider rows"): the C oracle against the frozen vectors
the Python oracle produced (tests/golden/make_rf1_wide_golden.py), and Python == C on random inputs."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import c_oracle as co, rf1

DIMS = (512, 1024)


@pytest.fixture(scope="module")
def wide(golden_dir):
    return json.load(open(os.path.join(golden_dir, "rf1_wide_golden.json")))


@pytest.fixture(scope="module")
def base(golden_dir):
    return json.load(open(os.path.join(golden_dir, "rf1_golden.json")))


def _dense(sparse, n):
    row = np.zeros(n, np.int8)
    for i, v in sparse:
        row[i] = v
    return row


@pytest.mark.parametrize("dim", DIMS)
def test_documents_and_query(wide, base, dim):
    g = wide[str(dim)]
    for which in ("sample_report", "long_doc"):
        F, ff, spans, ntok = co.featurize_doc(base[which]["text"].encode("utf-8"), dim)
        want = np.stack([_dense(r, dim) for r in g[which]["rows_sparse"]])
        assert ntok == g[which]["n_tokens"] == base[which]["n_tokens"]          # tokens do not depend on the width
        assert (F == want).all() and ff.tolist() == g[which]["ff"] and spans.tolist() == g[which]["spans"]
        assert spans.tolist() == base[which]["spans"]
    q = co.query_vector(base["sample_report"]["query"].encode(), dim)
    assert (q == _dense(g["demo_query"]["q_sparse"], dim)).all()
    F, _, _, _ = co.featurize_doc(base["sample_report"]["text"].encode("utf-8"), dim)
    assert rf1.scores(F, q).tolist() == g["demo_query"]["scores"]


@pytest.mark.parametrize("dim", DIMS)
def test_folding_a_wide_row_gives_the_narrow_row(base, dim):
    """bucket = hash & (D - 1): summing the D / 256 sub-rows of a wide row gives the 256-feature row
    (before saturation) -- the widths are refinements of one another."""
    text = base["long_doc"]["text"].encode()
    Fw, _, _, _ = co.featurize_doc(text, dim)
    Fn, _, _, _ = co.featurize_doc(text, 256)
    folded = Fw.astype(np.int32).reshape(len(Fw), dim // 256, 256).sum(axis=1)
    assert (np.minimum(folded, 127) == Fn).all()


@pytest.mark.parametrize("dim", DIMS)
def test_synthetic_corpus_top10_and_weights(wide, dim):
    g = wide[str(dim)]
    zb = rf1.zipf_bucket_table(dim=dim)
    assert hashlib.sha256(zb.astype("<u2").tobytes()).hexdigest() == g["zipf_bucket_sha256"]
    t = g["synth_top10"]
    F, ff = co.synth_rows(t["seed"], t["start"], t["n_rows"], zb, with_ff=True, dim=dim)
    assert hashlib.sha256(F.tobytes()).hexdigest() == g["synth_rows_sha256"]
    assert (F[0] == _dense(g["synth_row_0_sparse"], dim)).all()
    assert (ff == (F.astype(np.int32) ** 2).sum(axis=1)).all()
    seg = np.zeros(t["n_rows"], np.uint32)
    seg[1::5] = 1
    seg[7::11] = rf1.TOMBSTONE
    for case in t["cases"]:
        q = co.synth_query(t["seed"], case["qi"], zb, dim=dim)
        ids, sc, _ = co.score_topk(F, seg, q, case["scope"], k=10, id_base=t["id_base"], ff=ff)
        assert ids.tolist() == case["ids"] and sc.tolist() == case["scores"]
    df, n = co.bucket_df(F, seg, [0, 1])
    assert n == g["rf1w"]["n"] and hashlib.sha256(df.astype("<u8").tobytes()).hexdigest() == g["rf1w"]["df_sha256"]
    assert co.idf_weights(df, n).tolist() == g["rf1w"]["weights"]


@pytest.mark.parametrize("dim", DIMS)
def test_python_equals_c_random(dim):
    rng = np.random.default_rng(dim)
    words = [bytes(rng.integers(97, 123, rng.integers(1, 9)).astype(np.uint8)) for _ in range(400)]
    text = b" ".join(words[i] for i in rng.integers(0, 400, 900))
    a = co.featurize_doc(text, dim)
    b = rf1.featurize_doc(text, dim)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[2] == b[2]).all() and a[3] == b[3]
    zb = rf1.zipf_bucket_table(dim=dim)
    F = co.synth_rows(9, 0, 500, zb, dim=dim)
    assert (F == rf1.synth_rows(9, 0, 500, zb, dim=dim)).all()
    seg = rng.integers(0, 3, 500).astype(np.uint32)
    for qi in range(5):
        q = co.synth_query(9, qi, zb, dim=dim)
        assert (q == rf1.synth_query(9, qi, zb, dim=dim)).all()
        ids, sc, _ = co.score_topk(F, seg, q, [0, 2], k=10, id_base=7, ff=np.ones(500, np.int32))
        pi, ps = rf1.score_topk(F, seg, q, [0, 2], k=10, id_base=7)
        assert ids.tolist() == pi.tolist() and sc.tolist() == ps.tolist()
