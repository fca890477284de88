/* RF-1 CPU oracle, C restatement of oracle/SPEC.md (independent of oracle/rf1.py).
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (rag_foundation_b200/) never does.
 *
 * PARITY UNPINNED for ranking: the reference (/root/reference) has no retrieval arithmetic --
 * backend/app/services/gemini_rag.py:704-718 returns one canned citation and :517-551 calls a
 * remote service.  The only reference code restated here is the tokeniser rule,
 * scripts/benchmark/metrics.py:6,13-19 (_normalize + ARTICLES); see rf1_tokenize().
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp; the dot product is multi-versioned at run time,
 * AVX-512 VNNI when the host has it, so one binary is valid on this container and the GPU box).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#define RF1_D 256          /* default row width; every function takes `dim` (a power of two, 128 .. RF1_D_MAX; SPEC.md step 3) */
#define RF1_D_MAX 4096
#define RF1_L 128
#define RF1_S 112
#define RF1_TOMBSTONE 0xFFFFFFFFu

/* ------------------------------------------------------------------ step 3: FNV-1a 32 */
uint32_t rf1_fnv1a32(const uint8_t *p, size_t n) {
    uint32_t h = 0x811C9DC5u;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x01000193u; }
    return h;
}

/* ------------------------------------------------------------------ step 1: tokenise
 * metrics.py:13-19: lower(); [^a-z0-9\s] -> ' '; split(); drop ARTICLES (metrics.py:6). */
static inline uint8_t lower_byte(uint8_t b) { return (b >= 'A' && b <= 'Z') ? (uint8_t)(b + 32) : b; }
static inline int token_byte(uint8_t b) { return (b >= 'a' && b <= 'z') || (b >= '0' && b <= '9'); }

static int is_stopword(const uint8_t *p, size_t len) {
    if (len == 1) return lower_byte(p[0]) == 'a';
    if (len == 2) return lower_byte(p[0]) == 'a' && lower_byte(p[1]) == 'n';
    if (len == 3) return lower_byte(p[0]) == 't' && lower_byte(p[1]) == 'h' && lower_byte(p[2]) == 'e';
    return 0;
}

/* Kept tokens: writes up to max_tokens (start, end, bucket) triples; returns the total count. */
int64_t rf1_tokenize(const uint8_t *data, size_t n, int dim, int64_t *starts, int64_t *ends, uint16_t *buckets,
                     int64_t max_tokens) {
    int64_t cnt = 0;
    size_t i = 0;
    while (i < n) {
        if (!token_byte(lower_byte(data[i]))) { ++i; continue; }
        size_t j = i;
        uint32_t h = 0x811C9DC5u;
        while (j < n && token_byte(lower_byte(data[j]))) { h ^= lower_byte(data[j]); h *= 0x01000193u; ++j; }
        if (!is_stopword(data + i, j - i)) {
            if (cnt < max_tokens) {
                if (starts) starts[cnt] = (int64_t)i;
                if (ends) ends[cnt] = (int64_t)j;
                if (buckets) buckets[cnt] = (uint16_t)(h & (uint32_t)(dim - 1));
            }
            ++cnt;
        }
        i = j;
    }
    return cnt;
}

int64_t rf1_n_chunks(int64_t n_tokens) {
    if (n_tokens == 0) return 0;
    int64_t extra = n_tokens > RF1_L ? n_tokens - RF1_L : 0;
    return 1 + (extra + RF1_S - 1) / RF1_S;
}

static void row_from_buckets(const uint16_t *b, int64_t n, int dim, int8_t *row, int32_t *ff) {
    int32_t tf[RF1_D_MAX];
    memset(tf, 0, sizeof(int32_t) * (size_t)dim);
    for (int64_t i = 0; i < n; ++i) tf[b[i]]++;
    int32_t acc = 0;
    for (int d = 0; d < dim; ++d) {
        int32_t v = tf[d] > 127 ? 127 : tf[d];
        row[d] = (int8_t)v;
        acc += v * v;
    }
    if (ff) *ff = acc;
}

/* steps 1-4 for one document.  Returns n_chunks (may exceed max_chunks: then only the first
 * max_chunks rows are written), or -1 on allocation failure. */
int64_t rf1_featurize_doc(const uint8_t *data, size_t n, int dim, int8_t *F, int32_t *ff, int64_t *spans,
                          int64_t max_chunks, int64_t *n_tokens_out) {
    int64_t T = rf1_tokenize(data, n, dim, NULL, NULL, NULL, 0);
    if (n_tokens_out) *n_tokens_out = T;
    int64_t nc = rf1_n_chunks(T);
    if (T == 0) return 0;
    int64_t *st = (int64_t *)malloc(sizeof(int64_t) * (size_t)T);
    int64_t *en = (int64_t *)malloc(sizeof(int64_t) * (size_t)T);
    uint16_t *bk = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)T);
    if (!st || !en || !bk) { free(st); free(en); free(bk); return -1; }
    rf1_tokenize(data, n, dim, st, en, bk, T);
    for (int64_t w = 0; w < nc && w < max_chunks; ++w) {
        int64_t lo = (int64_t)RF1_S * w;
        int64_t hi = lo + RF1_L < T ? lo + RF1_L : T;
        row_from_buckets(bk + lo, hi - lo, dim, F + w * (int64_t)dim, ff ? ff + w : NULL);
        if (spans) { spans[2 * w] = st[lo]; spans[2 * w + 1] = en[hi - 1]; }
    }
    free(st); free(en); free(bk);
    return nc;
}

/* step 5 */
void rf1_query_vector(const uint8_t *data, size_t n, int dim, int8_t *q) {
    int64_t T = rf1_tokenize(data, n, dim, NULL, NULL, NULL, 0);
    uint16_t *bk = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)(T > 0 ? T : 1));
    rf1_tokenize(data, n, dim, NULL, NULL, bk, T);
    row_from_buckets(bk, T, dim, q, NULL);
    free(bk);
}

/* ------------------------------------------------------------------ step 6: int8 dot */
static int32_t dot256_generic(const int8_t *a, const int8_t *b, int dim) {
    int32_t s = 0;
    for (int d = 0; d < dim; ++d) s += (int32_t)a[d] * (int32_t)b[d];
    return s;
}

#if defined(__x86_64__)
__attribute__((target("avx2")))
static int32_t dot256_avx2(const int8_t *a, const int8_t *b, int dim) {
    __m256i acc = _mm256_setzero_si256();
    for (int d = 0; d < dim; d += 16) {
        __m256i x = _mm256_cvtepi8_epi16(_mm_loadu_si128((const __m128i *)(a + d)));
        __m256i y = _mm256_cvtepi8_epi16(_mm_loadu_si128((const __m128i *)(b + d)));
        acc = _mm256_add_epi32(acc, _mm256_madd_epi16(x, y));
    }
    __m128i lo = _mm_add_epi32(_mm256_castsi256_si128(acc), _mm256_extracti128_si256(acc, 1));
    lo = _mm_add_epi32(lo, _mm_shuffle_epi32(lo, 0x4E));
    lo = _mm_add_epi32(lo, _mm_shuffle_epi32(lo, 0xB1));
    return _mm_cvtsi128_si32(lo);
}

/* Features are counts in [0,127], so the unsigned x signed VNNI form is exact. */
__attribute__((target("avx512f,avx512bw,avx512vl,avx512vnni")))
static int32_t dot256_vnni(const int8_t *a, const int8_t *b, int dim) {
    __m512i acc = _mm512_setzero_si512();
    for (int d = 0; d < dim; d += 64)
        acc = _mm512_dpbusd_epi32(acc, _mm512_loadu_si512((const void *)(a + d)),
                                  _mm512_loadu_si512((const void *)(b + d)));
    return _mm512_reduce_add_epi32(acc);
}
#endif

typedef int32_t (*dot_fn)(const int8_t *, const int8_t *, int);   /* dim: a multiple of 64 */
static dot_fn pick_dot(void) {
#if defined(__x86_64__)
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512vnni") && __builtin_cpu_supports("avx512bw")) return dot256_vnni;
    if (__builtin_cpu_supports("avx2")) return dot256_avx2;
#endif
    return dot256_generic;
}

const char *rf1_dot_isa(void) {
#if defined(__x86_64__)
    dot_fn f = pick_dot();
    if (f == dot256_vnni) return "avx512vnni";
    if (f == dot256_avx2) return "avx2";
#endif
    return "generic";
}

int rf1_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ step 7: rank */
static inline uint64_t pack_key(int32_t s, uint64_t gid) {
    return ((uint64_t)(uint32_t)s << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)gid);
}

/* keep the k largest keys in `top` (descending, zero-filled) */
static inline void topk_insert(uint64_t *top, int k, uint64_t key) {
    if (key <= top[k - 1]) return;
    int i = k - 1;
    while (i > 0 && top[i - 1] < key) { top[i] = top[i - 1]; --i; }
    top[i] = key;
}

static int in_scope(uint32_t seg, const uint32_t *scope, int n_scope) {
    if (seg == RF1_TOMBSTONE) return 0;
    for (int i = 0; i < n_scope; ++i) if (scope[i] == seg) return 1;
    return 0;
}

/* Largest k keys over rows [row_lo,row_hi) of F that are in scope; out_keys zero-padded.
 * Returns the number of valid results. threads <= 0: all OpenMP threads. */
int rf1_score_topk_keys(const int8_t *F, const uint32_t *store_seg, int dim, int64_t row_lo, int64_t row_hi,
                        const int8_t *q, const uint32_t *scope, int n_scope, int k, uint64_t id_base,
                        uint64_t *out_keys, int threads) {
    if (k <= 0 || k > 64) return -1;
    dot_fn dot = pick_dot();
    int nt = threads > 0 ? threads : rf1_max_threads();
    uint64_t *all = (uint64_t *)calloc((size_t)nt * (size_t)k, sizeof(uint64_t));
    if (!all) return -1;
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
        int T = omp_get_num_threads();
#else
        int t = 0, T = 1;
#endif
        uint64_t *top = all + (size_t)t * (size_t)k;
        int64_t span = row_hi - row_lo;
        int64_t lo = row_lo + span * t / T, hi = row_lo + span * (t + 1) / T;
        for (int64_t r = lo; r < hi; ++r) {
            if (!in_scope(store_seg[r], scope, n_scope)) continue;
            int32_t s = dot(F + r * (int64_t)dim, q, dim);
            topk_insert(top, k, pack_key(s, id_base + (uint64_t)r));
        }
    }
    for (int i = 0; i < k; ++i) out_keys[i] = 0;
    for (int i = 0; i < nt * k; ++i) if (all[i]) topk_insert(out_keys, k, all[i]);
    free(all);
    int n = 0;
    while (n < k && out_keys[n]) ++n;
    return n;
}

/* step 8 */
float rf1_cosine(int32_t s, int32_t qq, int32_t ff) {
    float nq = sqrtf((float)qq), nf = sqrtf((float)ff);
    float den = nq * nf;
    return den == 0.0f ? 0.0f : (float)s / den;
}

/* steps 6-8 with unpacked outputs; ff may be NULL (then out_cos untouched). */
int rf1_score_topk(const int8_t *F, const uint32_t *store_seg, const int32_t *ff, int dim, int64_t n_rows,
                   const int8_t *q, const uint32_t *scope, int n_scope, int k, uint64_t id_base,
                   uint64_t *out_ids, int32_t *out_scores, float *out_cos, int threads) {
    uint64_t keys[64];
    int n = rf1_score_topk_keys(F, store_seg, dim, 0, n_rows, q, scope, n_scope, k, id_base, keys, threads);
    if (n < 0) return n;
    int32_t qq = 0;
    for (int d = 0; d < dim; ++d) qq += (int32_t)q[d] * (int32_t)q[d];
    for (int i = 0; i < n; ++i) {
        out_ids[i] = (uint64_t)(0xFFFFFFFFu - (uint32_t)(keys[i] & 0xFFFFFFFFu));
        out_scores[i] = (int32_t)(keys[i] >> 32);
        if (ff && out_cos) out_cos[i] = rf1_cosine(out_scores[i], qq, ff[out_ids[i] - id_base]);
    }
    return n;
}

/* nq queries one after another (each uses all threads): the CPU baseline for batched configs. */
int rf1_score_topk_batch(const int8_t *F, const uint32_t *store_seg, int dim, int64_t n_rows, const int8_t *Q,
                         int nq, const uint32_t *scope_flat, const int32_t *scope_off, int k,
                         uint64_t id_base, uint64_t *out_keys, int threads) {
    for (int i = 0; i < nq; ++i) {
        int n = rf1_score_topk_keys(F, store_seg, dim, 0, n_rows, Q + (size_t)i * (size_t)dim,
                                    scope_flat + scope_off[i], scope_off[i + 1] - scope_off[i], k,
                                    id_base, out_keys + (size_t)i * (size_t)k, threads);
        if (n < 0) return n;
    }
    return 0;
}

int rf1_merge_topk(const uint64_t *keys, int64_t n, int k, uint64_t *out) {
    if (k <= 0 || k > 64) return -1;
    for (int i = 0; i < k; ++i) out[i] = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!keys[i]) continue;
        int dup = 0;
        for (int j = 0; j < k; ++j) if (out[j] == keys[i]) dup = 1;
        if (!dup) topk_insert(out, k, keys[i]);
    }
    int m = 0;
    while (m < k && out[m]) ++m;
    return m;
}

/* ---- RF-1w (SPEC.md "IDF-weighted variant"): document frequencies, weights, weighted query ---- */
int64_t rf1_bucket_df(const int8_t *F, const uint32_t *store_seg, int dim, int64_t n_rows, const uint32_t *scope, int n_scope,
                      uint64_t *df /* [dim] */) {
    int64_t n = 0;
    memset(df, 0, sizeof(uint64_t) * (size_t)dim);
    for (int64_t r = 0; r < n_rows; ++r) {
        if (!in_scope(store_seg[r], scope, n_scope)) continue;
        ++n;
        const int8_t *row = F + r * (int64_t)dim;
        for (int d = 0; d < dim; ++d) df[d] += row[d] > 0;
    }
    return n;
}

void rf1_idf_weights(const uint64_t *df, uint64_t n, int dim, uint8_t *w /* [dim] */) {
    for (int d = 0; d < dim; ++d) {
        const uint64_t r = ((n + 1) * 256) / (df[d] + 1);
        int lg = 63;
        while (!(r >> lg)) --lg;
        const uint64_t v = 4 + 4 * (uint64_t)(lg - 8) + ((r >> (lg - 2)) & 3);
        w[d] = (uint8_t)(v < 31 ? v : 31);
    }
}

void rf1_weight_query(const int8_t *q, const uint8_t *w, int dim, int8_t *qw) {
    for (int d = 0; d < dim; ++d) {
        const int v = (int)q[d] * (int)w[d];
        qw[d] = (int8_t)(v < 127 ? v : 127);
    }
}

/* ------------------------------------------------------------------ synthetic corpora */
static inline uint64_t mix64(uint64_t seed, uint64_t a, uint64_t b) {
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + a * 0xBF58476D1CE4E5B9ull + b * 0x94D049BB133111EBull
                 + 0x2545F4914F6CDD1Dull;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

uint64_t rf1_mix64(uint64_t seed, uint64_t a, uint64_t b) { return mix64(seed, a, b); }

/* zb: uint16[65536] bucket (< dim) of the decimal-ASCII token of zipf_vocab[r] */
void rf1_synth_rows(uint64_t seed, uint64_t start, int64_t n, const uint16_t *zb, int dim, int8_t *F, int32_t *ff,
                    int threads) {
    int nt = threads > 0 ? threads : rf1_max_threads();
    (void)nt;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (int64_t i = 0; i < n; ++i) {
        uint64_t c = start + (uint64_t)i;
        int len = 64 + (int)(mix64(seed ^ 0xA5ull, c, 0) & 63);
        uint16_t bk[128];
        for (int j = 0; j < len; ++j) bk[j] = zb[mix64(seed, c, (uint64_t)j) >> 48];
        row_from_buckets(bk, len, dim, F + i * (int64_t)dim, ff ? ff + i : NULL);
    }
}

void rf1_synth_query(uint64_t seed, uint64_t qi, int n_tokens, const uint16_t *zb, int dim, int8_t *q) {
    uint16_t bk[256];
    if (n_tokens > 256) n_tokens = 256;
    for (int j = 0; j < n_tokens; ++j) bk[j] = zb[mix64(seed ^ 0x51ull, qi, (uint64_t)j) >> 48];
    row_from_buckets(bk, n_tokens, dim, q, NULL);
}
