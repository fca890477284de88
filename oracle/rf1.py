"""RF-1 CPU oracle (readable Python / numpy restatement of oracle/SPEC.md).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product (rag_foundation_b200/) never does.

PARITY UNPINNED for ranking: the reference has no retrieval arithmetic (gemini_rag.py:704-718 is a
canned citation; real mode is a remote service, gemini_rag.py:517-551).  What follows reference
code, with the file:line it follows (paths relative to /root/reference):

  tokenize()              scripts/benchmark/metrics.py:6,13-19   (_normalize + ARTICLES)
  contents_to_text()      backend/app/services/gemini_rag.py:640-654
  mock_final_response()   backend/app/services/gemini_rag.py:704-718 (object shape only)
  extract_citations()     backend/app/services/gemini_rag.py:554-595
  citation_frames()       backend/app/routes/chat.py:576-586
  finish_frame()          backend/app/routes/chat.py:589-603

Everything else (chunking, hashing, int8 features, int32 scores, rank, cosine, synthetic corpora)
follows oracle/SPEC.md, this repository's own frozen definition.
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace
from typing import Any, Iterable, Iterator, List, Sequence, Tuple

import numpy as np

D = 256
CHUNK_TOKENS = 128
CHUNK_STRIDE = 112
TOPK = 10
TOMBSTONE = 0xFFFFFFFF
STOPWORDS = (b"a", b"an", b"the")  # metrics.py:6

_M64 = (1 << 64) - 1
_HERE = os.path.dirname(os.path.abspath(__file__))
ZIPF_PATH = os.path.join(_HERE, "..", "rag_foundation_b200", "data", "zipf_vocab_u16.bin")


# ----------------------------------------------------------------------------- step 1: tokenise
def _lower_byte(b: int) -> int:
    return b + 32 if 65 <= b <= 90 else b


def _is_token_byte(b: int) -> bool:
    return (97 <= b <= 122) or (48 <= b <= 57)


def tokenize(data: bytes) -> List[Tuple[int, int, bytes]]:
    """Kept tokens as (byte_start, byte_end, lowercase_bytes); follows metrics.py:13-19."""
    out: List[Tuple[int, int, bytes]] = []
    n = len(data)
    i = 0
    while i < n:
        if not _is_token_byte(_lower_byte(data[i])):
            i += 1
            continue
        j = i
        while j < n and _is_token_byte(_lower_byte(data[j])):
            j += 1
        tok = bytes(_lower_byte(b) for b in data[i:j])
        if tok not in STOPWORDS:
            out.append((i, j, tok))
        i = j
    return out


def normalize(text: str) -> str:
    """What metrics.py:_normalize returns, rebuilt from tokenize() (ASCII-equivalent)."""
    return " ".join(t.decode("ascii") for _, _, t in tokenize(text.encode("utf-8")))


# ----------------------------------------------------------------------------- step 3: hash
def fnv1a32(data: bytes) -> int:
    h = 0x811C9DC5
    for b in data:
        h ^= b
        h = (h * 0x01000193) & 0xFFFFFFFF
    return h


def bucket(tok: bytes, dim: int = D) -> int:
    """SPEC.md step 3; `dim` is the row width (a power of two: 256 by default, 512 / 1024 for the wider variants)."""
    return fnv1a32(tok) & (dim - 1)


# ----------------------------------------------------------------------------- step 2: chunk
def n_chunks_for(n_tokens: int) -> int:
    if n_tokens == 0:
        return 0
    return 1 + (max(n_tokens - CHUNK_TOKENS, 0) + CHUNK_STRIDE - 1) // CHUNK_STRIDE


def chunk_windows(n_tokens: int) -> List[Tuple[int, int]]:
    out = []
    for w in range(n_chunks_for(n_tokens)):
        lo = CHUNK_STRIDE * w
        out.append((lo, min(lo + CHUNK_TOKENS, n_tokens)))
    return out


# ----------------------------------------------------------------------------- steps 4-5: features
def _row_from_buckets(buckets: Iterable[int], dim: int = D) -> np.ndarray:
    tf = np.zeros(dim, dtype=np.int64)
    for b in buckets:
        tf[b] += 1
    return np.minimum(tf, 127).astype(np.int8)


def featurize_doc(data: bytes, dim: int = D):
    """-> (F int8 [n,dim], ff int32 [n], spans int64 [n,2] byte spans, n_tokens)."""
    toks = tokenize(data)
    wins = chunk_windows(len(toks))
    F = np.zeros((len(wins), dim), dtype=np.int8)
    spans = np.zeros((len(wins), 2), dtype=np.int64)
    for w, (lo, hi) in enumerate(wins):
        F[w] = _row_from_buckets((bucket(t, dim) for _, _, t in toks[lo:hi]), dim)
        spans[w, 0] = toks[lo][0]
        spans[w, 1] = toks[hi - 1][1]
    ff = (F.astype(np.int32) ** 2).sum(axis=1).astype(np.int32)
    return F, ff, spans, len(toks)


def query_vector(data: bytes, dim: int = D) -> np.ndarray:
    return _row_from_buckets((bucket(t, dim) for _, _, t in tokenize(data)), dim)


# ----------------------------------------------------------------------------- steps 6-8
def scores(F: np.ndarray, q: np.ndarray) -> np.ndarray:
    return (F.astype(np.int32) @ q.astype(np.int32)).astype(np.int32)


def pack_key(score: int, gid: int) -> int:
    return (int(score) << 32) | (0xFFFFFFFF - int(gid))


def unpack_key(key: int) -> Tuple[int, int]:
    return int(key) >> 32, 0xFFFFFFFF - (int(key) & 0xFFFFFFFF)


def score_topk(F, store_seg, q, scope: Sequence[int], k: int = TOPK, id_base: int = 0,
               row_ranges: Sequence[Tuple[int, int]] | None = None):
    """Steps 6-7. -> (ids uint64 [m], scores int32 [m]), m <= k, ordered (score desc, id asc)."""
    store_seg = np.asarray(store_seg, dtype=np.uint32)
    s = scores(F, q).astype(np.int64)
    ok = np.isin(store_seg, np.asarray(list(scope), dtype=np.uint32)) & (store_seg != TOMBSTONE)
    if row_ranges is not None:
        in_rng = np.zeros(len(s), dtype=bool)
        for lo, hi in row_ranges:
            in_rng[lo:hi] = True
        ok &= in_rng
    rows = np.nonzero(ok)[0]
    gids = rows.astype(np.int64) + id_base
    order = np.lexsort((gids, -s[rows]))[:k]
    return gids[order].astype(np.uint64), s[rows][order].astype(np.int32)


def merge_topk(key_lists: Iterable[Sequence[int]], k: int = TOPK) -> List[int]:
    """k-way merge of packed keys (step 7): largest k distinct non-zero keys, descending."""
    allk = sorted({int(x) for ks in key_lists for x in ks if int(x) != 0}, reverse=True)
    return allk[:k]


def cosine(score, qq, ff) -> np.ndarray:
    """Step 8, float32 operations in spec order."""
    s = np.asarray(score, dtype=np.int64).astype(np.float32)
    nq = np.sqrt(np.asarray(qq, dtype=np.int64).astype(np.float32))
    nf = np.sqrt(np.asarray(ff, dtype=np.int64).astype(np.float32))
    den = (nq * nf).astype(np.float32)
    out = np.zeros_like(s, dtype=np.float32)
    nz = den != 0
    out[nz] = (s[nz] / den[nz]).astype(np.float32)
    return out


# ----------------------------------------------------------------------------- RF-1w (SPEC.md "IDF-weighted variant")
def bucket_df(F, store_seg, scope: Sequence[int]):
    """-> (df uint64 [D], n): per-bucket document frequency over the live rows of the scope."""
    store_seg = np.asarray(store_seg, dtype=np.uint32)
    ok = np.isin(store_seg, np.asarray(list(scope), dtype=np.uint32)) & (store_seg != TOMBSTONE)
    return (np.asarray(F)[ok] > 0).sum(axis=0).astype(np.uint64), int(ok.sum())


def idf_weights(df, n: int) -> np.ndarray:
    w = np.zeros(len(df), dtype=np.uint8)
    for d, x in enumerate(df):
        r = ((int(n) + 1) * 256) // (int(x) + 1)
        lg = r.bit_length() - 1
        w[d] = min(4 + 4 * (lg - 8) + ((r >> (lg - 2)) & 3), 31)
    return w


def weight_query(q, w) -> np.ndarray:
    return np.minimum(np.asarray(q, dtype=np.int32) * np.asarray(w, dtype=np.int32), 127).astype(np.int8)


# ----------------------------------------------------------------------------- synthetic corpora
def load_zipf_vocab() -> np.ndarray:
    t = np.fromfile(ZIPF_PATH, dtype="<u2")
    assert t.shape == (65536,)
    return t


def zipf_bucket_table(zipf_vocab: np.ndarray | None = None, dim: int = D) -> np.ndarray:
    """uint16[65536]: bucket (< dim) of the decimal-ASCII token of zipf_vocab[r]."""
    zv = load_zipf_vocab() if zipf_vocab is None else zipf_vocab
    per_vocab = np.array([bucket(str(v).encode(), dim) for v in range(int(zv.max()) + 1)], dtype=np.uint16)
    return per_vocab[zv]


def mix64(seed, a, b):
    """splitmix64-finalised counter hash; accepts python ints or numpy uint64 arrays."""
    with np.errstate(over="ignore"):
        seed = np.uint64(seed)
        a = np.asarray(a, dtype=np.uint64)
        b = np.asarray(b, dtype=np.uint64)
        x = (seed * np.uint64(0x9E3779B97F4A7C15) + a * np.uint64(0xBF58476D1CE4E5B9)
             + b * np.uint64(0x94D049BB133111EB) + np.uint64(0x2545F4914F6CDD1D))
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def synth_rows(seed: int, start: int, n: int, zb: np.ndarray | None = None, dim: int = D) -> np.ndarray:
    """int8 [n,dim] rows `start .. start+n-1` of the synthetic corpus `seed` (zb: the bucket table for `dim`)."""
    zb = zipf_bucket_table(dim=dim) if zb is None else zb
    c = np.arange(start, start + n, dtype=np.uint64)
    lens = 64 + (mix64(seed ^ 0xA5, c, 0) & np.uint64(63)).astype(np.int64)
    tf = np.zeros((n, dim), dtype=np.int32)
    rows = np.arange(n)
    for j in range(127):
        live = lens > j
        if not live.any():
            break
        r = (mix64(seed, c, j) >> np.uint64(48)).astype(np.int64)
        np.add.at(tf, (rows[live], zb[r[live]].astype(np.int64)), 1)
    return np.minimum(tf, 127).astype(np.int8)


def synth_query(seed: int, qi: int, zb: np.ndarray | None = None, n_tokens: int = 8, dim: int = D) -> np.ndarray:
    zb = zipf_bucket_table(dim=dim) if zb is None else zb
    r = (mix64(seed ^ 0x51, qi, np.arange(n_tokens, dtype=np.uint64)) >> np.uint64(48)).astype(np.int64)
    return _row_from_buckets((int(zb[x]) for x in r), dim)


def synth_text(seed: int, n_tokens: int, zipf_vocab: np.ndarray | None = None) -> bytes:
    """ASCII text for the ingest benchmark: decimal vocab ids, separators cycle ' ', ', ', '\n'."""
    zv = load_zipf_vocab() if zipf_vocab is None else zipf_vocab
    j = np.arange(n_tokens, dtype=np.uint64)
    r = (mix64(seed ^ 0x7E, 0, j) >> np.uint64(48)).astype(np.int64)
    seps = (" ", ", ", "\n")
    return "".join(str(int(zv[x])) + seps[i % 3] for i, x in enumerate(r)).encode()


# ----------------------------------------------------------------------------- adapter shapes
def contents_to_text(contents: Any) -> str:
    """gemini_rag.py:640-654: last non-empty user turn."""
    if isinstance(contents, str):
        return contents
    if isinstance(contents, list):
        for item in reversed(contents):
            if isinstance(item, str) and item.strip():
                return item.strip()
            if isinstance(item, dict):
                parts = item.get("parts")
                if isinstance(parts, list) and parts and isinstance(parts[0], dict):
                    text = parts[0].get("text")
                    if isinstance(text, str) and text.strip():
                        return text.strip()
    return str(contents)


def final_response(grounding: Sequence[dict]) -> Any:
    """Object shape of gemini_rag.py:704-718 carrying `grounding` retrieved contexts."""
    usage = SimpleNamespace(prompt_token_count=0, candidates_token_count=0)
    chunks = [SimpleNamespace(retrieved_context=SimpleNamespace(**g), web=None) for g in grounding]
    cand = SimpleNamespace(grounding_metadata=SimpleNamespace(grounding_chunks=chunks), usage_metadata=usage)
    return SimpleNamespace(text=None, candidates=[cand], usage_metadata=usage)


def extract_citations(response: Any) -> List[dict]:
    """gemini_rag.py:554-595 (retrieved_context branch)."""
    out: List[dict] = []
    try:
        gm = getattr(response.candidates[0], "grounding_metadata", None)
        if not gm:
            return out
        for i, ch in enumerate(list(getattr(gm, "grounding_chunks", []) or [])):
            rc = getattr(ch, "retrieved_context", None)
            if rc:
                out.append({"index": i, "source_type": "retrieved_context", "uri": getattr(rc, "uri", None),
                            "title": getattr(rc, "title", None), "snippet": getattr(rc, "text", None),
                            "store": getattr(rc, "file_search_store", None)})
        return out
    except (AttributeError, KeyError, IndexError, TypeError):
        return out


def citation_frames(citations: Sequence[dict]) -> Iterator[str]:
    """routes/chat.py:576-586."""
    for c in citations:
        payload = {"type": "source-document", "sourceId": f"cit-{c['index']}", "mediaType": "file",
                   "title": c.get("title") or c.get("uri") or "Source", "snippet": c.get("snippet")}
        yield f"data: {json.dumps(payload)}\n\n"


def finish_frame(prompt_tokens: int, completion_tokens: int, model: str) -> str:
    """routes/chat.py:589-603."""
    payload = {"type": "finish", "finishReason": "stop", "promptTokens": prompt_tokens,
               "completionTokens": completion_tokens,
               "usage": {"prompt_tokens": prompt_tokens, "completion_tokens": completion_tokens, "model": model}}
    return f"data: {json.dumps(payload)}\n\n"
