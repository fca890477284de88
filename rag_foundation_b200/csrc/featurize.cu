// Ingest-side chunk featurisation on the GPU (sm_100a): RF-1 steps 1-5 (oracle/SPEC.md).
//
// Replaces the work GeminiRag.upload_file hands to the remote service (reference
// backend/app/services/gemini_rag.py:307-352; the mock, :614-629, reads nothing), reached from the
// ARQ worker at backend/app/services/ingestion.py:45-52.  The tokeniser rule restates
// scripts/benchmark/metrics.py:6,13-19 at byte level.
//
// Byte/integer work, bound by the H2D copy of the text (PCIe), which the engine overlaps with it:
//   tokenise  ONE pass per copied chunk of text: a CTA stages a contiguous span (<= 8 KB) in shared memory,
//             derives token-byte / kept-start masks with packed arithmetic, hashes each token (FNV-1a 32) into
//             shared-memory records, and places (bucket, byte end [, byte start of a window's first token]) at
//             the global token ordinals once it has summed the token counts of the CTAs before it
//   rows      one warp per chunk window [112 w, 112 w + 128): shared-memory byte histogram of the window's
//             buckets = the dim-byte int8 row itself; its sum of squares, store word, byte span
// A token is "kept" when it is not one of a / an / the.
#include <algorithm>
#include <cstdlib>

#include "rf_device.cuh"
#include "rf_internal.h"

namespace rf {

namespace {

constexpr int kFeatThreads = 256;
constexpr int kChunkTokens = 128;
constexpr int kChunkStride = 112;

__device__ __forceinline__ uint8_t lower_byte(uint8_t b) { return (b >= 'A' && b <= 'Z') ? b + 32 : b; }
__device__ __forceinline__ bool token_byte(uint8_t lowered) {
    return (lowered >= 'a' && lowered <= 'z') || (lowered >= '0' && lowered <= '9');
}

// s points at lowered bytes with s[-1] .. s[+3] readable.  True when a kept token starts at s[0].
__device__ __forceinline__ bool kept_start(const uint8_t *s) {
    const uint8_t c0 = s[0];
    if (!token_byte(c0) || token_byte(s[-1])) return false;
    const uint8_t c1 = s[1], c2 = s[2], c3 = s[3];
    const bool t1 = token_byte(c1), t2 = token_byte(c2), t3 = token_byte(c3);
    const bool stop = (c0 == 'a' && !t1) || (c0 == 'a' && c1 == 'n' && !t2) ||
                      (c0 == 't' && c1 == 'h' && c2 == 'e' && !t3);
    return !stop;
}

// four packed ASCII bytes -> lower case (bytes >= 0x80 are left alone)
__device__ __forceinline__ uint32_t lower4(uint32_t x) {
    const uint32_t lo7 = x & 0x7F7F7F7Fu;
    const uint32_t ge_A = lo7 + 0x3F3F3F3Fu;           // bit 7 of a byte set <=> byte >= 'A' (0x41)
    const uint32_t ge_bracket = lo7 + 0x25252525u;     // bit 7 set <=> byte >= '[' (0x5B)
    const uint32_t upper = ge_A & ~ge_bracket & ~x & 0x80808080u;
    return x | (upper >> 2);                            // + 0x20
}

// A count word carries its own tag (launch number) and nothing else is read on its strength, so polling it needs
// no acquire: a relaxed L2 load (ld.acquire.gpu costs an L1 invalidation, CCTL.IVALL, per poll).
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- single-pass tokeniser ----------------------------------------------------------------------------------
// CTA c of a launch (numbered by a ticket, so the CTAs before it have always started) keeps a CONTIGUOUS span
// of the text (8 KB by default, at most 12 KB) in shared memory and produces its tokens there before it needs
// to know how many tokens precede it.  The kernel is shaped by two measurements of its predecessor (one 4 KB block
// per CTA, decoupled look-back, one thread hashing each token byte by byte: 136 us per 22.8 MB): ~4.9 warp
// instructions per text byte, and two scattered 2 / 4-byte global stores per token.
//   stage   16 bytes per thread and step: lower-cased with packed arithmetic, and the "token byte" /
//           "letter" predicates of the 16 bytes computed 4 bytes per operation and kept as two 16-bit masks
//           per 16-byte unit
//   flag    kept-token starts from the masks alone (start = token bit whose predecessor is clear; lengths
//           1-3 from shifted masks); only short tokens that begin with a letter look at their bytes for
//           a / an / the.  The CTA's token count is published at once, tagged with the launch number
//   hash    a thread walks the 20 byte positions of its unit (16 + 4 look-ahead) unrolled out of registers.
//           Adding the start bits S to the token-byte mask T carries through each kept run: T & ~(T + S) are the
//           bytes of the unit's kept tokens (they advance the FNV-1a state), (T + S) & ~T marks where a kept
//           token ends -- there (bucket, byte end) go to the token's CTA-local index in shared memory with
//           predicated stores and the state starts over; a token that runs further is finished from the masks
//   place   only now the counts of the CTAs before this one are summed (every thread takes a share of those
//           words; by now they have long been published), and the token records leave in coalesced stores at
//           their global ordinals; the byte start of every 112th token (a chunk window opens there) is found
//           again from the kept-start masks
// Tokens before this launch's bytes: ctl[kCtlTokens], read before the count is published and advanced by the
// launch's last CTA after it has seen every other count.  The text may still be arriving: bytes are valid up to
// `avail_end`; a token that runs past it (only possible for a token longer than a whole copy chunk) is parked
// and finished by hash_deferred_kernel.
constexpr uint32_t kSpanDefaultBytes = 8u << 10;
constexpr int kSpanMaxIt = (static_cast<int>(kSpanMaxBytes) / 16 + kFeatThreads - 1) / kFeatThreads;   // steps of 4 KB
constexpr int kSpanRows = kSpanMaxIt * (kFeatThreads / 32);                                               // warp rows per CTA
static_assert(kSpanRows <= 32, "a warp scans the row totals one per lane");
static_assert(kSpanMaxBytes + 64 < 0xFFFFu, "token ends inside the staged span fit 16 bits");

// Shared-memory layout of a CTA with span bytes of text (span a multiple of 16; U = span / 16 units):
//   [0, 16)            the byte before the span in [15]
//   [16, 16 + span+16) lowered text, one look-ahead unit included
//   s_tok  u16[U + 4]  token-byte mask of unit u at [1 + u] (bit j = byte j); [0] bit 15 = the byte before the span
//   s_aux  u16[U + 4]  letter mask of the unit, replaced by its kept-start mask
//   s_bkt  u16[span/2] bucket of the CTA's i-th kept token (a kept token takes >= 2 bytes of text)
//   s_end  u16[span/2] its byte end relative to the span (0xFFFF: it runs far past the span, see s_long_end)
__host__ __device__ constexpr uint32_t span_tok_offset(uint32_t span) { return 16u + span + 16u; }
__host__ __device__ constexpr uint32_t span_aux_offset(uint32_t span) { return span_tok_offset(span) + (span / 16u + 4u) * 2u; }
__host__ __device__ constexpr uint32_t span_bkt_offset(uint32_t span) { return span_aux_offset(span) + (span / 16u + 4u) * 2u; }
__host__ __device__ constexpr uint32_t span_end_offset(uint32_t span) { return span_bkt_offset(span) + span + 16u; }
__host__ __device__ constexpr uint32_t span_smem_bytes(uint32_t span) { return span_end_offset(span) + span + 16u; }

// token-byte / letter predicates of four lowered bytes -> 4 bits each (bit i = byte i)
__device__ __forceinline__ uint32_t tok_bits4(uint32_t x, uint32_t &letter_bits) {
    const uint32_t lo7 = x & 0x7F7F7F7Fu;
    const uint32_t ascii = ~x & 0x80808080u;
    const uint32_t letter = (lo7 + 0x1F1F1F1Fu) & ~(lo7 + 0x05050505u) & ascii;   // 'a' .. 'z'
    const uint32_t digit = (lo7 + 0x50505050u) & ~(lo7 + 0x46464646u) & ascii;    // '0' .. '9'
    // bits 7 / 15 / 23 / 31 -> 0 / 1 / 2 / 3: the four partial products that land on bits 21-24 are the only ones there
    letter_bits = (((letter >> 7) * 0x00204081u) >> 21) & 0xFu;
    return ((((letter | digit) >> 7) * 0x00204081u) >> 21) & 0xFu;
}

__global__ void __launch_bounds__(kFeatThreads) tokenize_span_kernel(const TokenizeArgs a) {
    extern __shared__ __align__(16) uint8_t span_smem[];
    __shared__ uint32_t s_row[32];                    // tokens per warp row (step, warp)
    __shared__ uint32_t s_red[kFeatThreads / 32];
    __shared__ uint32_t s_cta, s_base, s_long_end, s_def_li, s_def_start;
    if (threadIdx.x == 0) {
        s_base = __ldcg(a.ctl + kCtlTokens);          // read before this CTA publishes anything: see above
        s_cta = atomicAdd(a.ctl + kCtlTicket, 1u) - a.ticket_base;
        s_def_li = 0xFFFFFFFFu;
    }
    __syncthreads();
    const uint32_t cta = s_cta;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t span_lo = a.byte_lo + static_cast<size_t>(cta) * a.span;
    const size_t span_hi = span_lo < a.byte_hi ? (span_lo + a.span < a.byte_hi ? span_lo + a.span : a.byte_hi) : span_lo;
    const uint32_t n_units = static_cast<uint32_t>((span_hi - span_lo + 15) / 16);
    uint8_t *s_txt = span_smem + 16;                  // s_txt[i] = lowered text[span_lo + i]
    uint16_t *s_tok = reinterpret_cast<uint16_t *>(span_smem + span_tok_offset(a.span));
    uint16_t *s_aux = reinterpret_cast<uint16_t *>(span_smem + span_aux_offset(a.span));
    uint16_t *s_bkt = reinterpret_cast<uint16_t *>(span_smem + span_bkt_offset(a.span));
    uint16_t *s_end = reinterpret_cast<uint16_t *>(span_smem + span_end_offset(a.span));

    // ---- stage: n_units + 1 units (the last one is look-ahead), bytes past the document read as separators;
    // a thread's next unit is loaded before the current one is worked on
    auto load_unit = [&](uint32_t u) {
        const size_t at = span_lo + static_cast<size_t>(u) * 16;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (u <= n_units && at < a.n) {
            v = *reinterpret_cast<const uint4 *>(a.text + at);        // the text buffer is padded by 64 bytes
            if (at + 16 > a.n) {
                const uint32_t keep = static_cast<uint32_t>(a.n - at);   // 1 .. 15
                uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int valid = static_cast<int>(keep) - 4 * i;
                    w[i] = valid >= 4 ? w[i] : valid <= 0 ? 0u : (w[i] & ((1u << (8 * valid)) - 1u));
                }
                v = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        return v;
    };
    uint4 nxt = load_unit(threadIdx.x);
    for (uint32_t u = threadIdx.x; u <= n_units; u += kFeatThreads) {
        uint4 v = nxt;
        nxt = load_unit(u + kFeatThreads);
        v = make_uint4(lower4(v.x), lower4(v.y), lower4(v.z), lower4(v.w));      // (zero bytes stay zero)
        *reinterpret_cast<uint4 *>(s_txt + u * 16) = v;
        uint32_t l0, l1, l2, l3;
        const uint32_t t0 = tok_bits4(v.x, l0), t1 = tok_bits4(v.y, l1), t2 = tok_bits4(v.z, l2), t3 = tok_bits4(v.w, l3);
        s_tok[1 + u] = static_cast<uint16_t>(t0 | (t1 << 4) | (t2 << 8) | (t3 << 12));
        s_aux[u] = static_cast<uint16_t>(l0 | (l1 << 4) | (l2 << 8) | (l3 << 12));
    }
    if (threadIdx.x == 0) {
        const uint8_t before = span_lo > 0 ? lower_byte(a.text[span_lo - 1]) : 0;
        s_txt[-1] = before;
        s_tok[0] = token_byte(before) ? 0x8000u : 0u;
    }
    if (threadIdx.x < 32) s_row[threadIdx.x] = 0;
    __syncthreads();

    // ---- flag the kept-token starts + count
    const uint32_t n_it = (n_units + kFeatThreads - 1) / kFeatThreads;
    for (uint32_t it = 0; it < n_it; ++it) {
        const uint32_t u = it * kFeatThreads + threadIdx.x;
        uint32_t S = 0;
        if (u < n_units) {
            const uint32_t tp = s_tok[u], tc = s_tok[u + 1], tn = s_tok[u + 2];
            const uint32_t T = tc | ((tn & 0xFu) << 16);             // bit j = byte j of the unit is a token byte, j < 20
            S = T & ~((T << 1) | (tp >> 15)) & 0xFFFFu;              // ... and the byte before it is not: a token starts
            const uint32_t n1 = ~(T >> 1), n2 = ~(T >> 2), n3 = ~(T >> 3);
            const uint32_t L1 = S & n1, L2 = S & ~n1 & n2, L3 = S & ~n1 & ~n2 & n3;   // tokens of 1 / 2 / 3 bytes
            uint32_t cand = (L1 | L2 | L3) & s_aux[u];               // short and starting with a letter: a / an / the?
            while (cand) {
                const int j = __ffs(cand) - 1;
                cand &= cand - 1;
                const uint8_t *c = s_txt + u * 16 + j;
                const bool stop = ((L1 >> j) & 1u)   ? c[0] == 'a'
                                  : ((L2 >> j) & 1u) ? (c[0] == 'a' && c[1] == 'n')
                                                     : (c[0] == 't' && c[1] == 'h' && c[2] == 'e');
                if (stop) S &= ~(1u << j);
            }
            s_aux[u] = static_cast<uint16_t>(S);
        }
        const uint32_t row_total = __reduce_add_sync(kFull, __popc(S));
        if (lane == 0) s_row[it * (kFeatThreads / 32) + warp] = row_total;
    }
    __syncthreads();
    // exclusive prefix of the (<= 24) row totals: every warp scans them itself -- no second barrier, no broadcast
    uint32_t row_ex, total;
    {
        const uint32_t tr = s_row[lane];
        uint32_t inc_r = tr;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(kFull, inc_r, o);
            if (lane >= o) inc_r += x;
        }
        row_ex = inc_r - tr;
        total = __shfl_sync(kFull, inc_r, 31);
    }
    if (threadIdx.x == 0) st_release_u64(a.state + cta, (static_cast<uint64_t>(a.seq) << 32) | total);

    // ---- hash: token records at their CTA-local index
    const uint32_t staged = (n_units + 1) * 16;       // bytes of s_txt (and bits of s_tok) that hold text
    const size_t avail_left = a.avail_end - span_lo;  // launches only cover blocks that start below avail_end
    const uint32_t avail_rel = avail_left > 0xFFFFFFF0ull ? 0xFFFFFFF0u : static_cast<uint32_t>(avail_left);
    const uint32_t dim_mask = a.dim_mask;
    const uint32_t bkt_s = static_cast<uint32_t>(__cvta_generic_to_shared(s_bkt));
    const uint32_t end_s = static_cast<uint32_t>(__cvta_generic_to_shared(s_end));
    for (uint32_t it = 0; it < n_it; ++it) {
        const uint32_t u = it * kFeatThreads + threadIdx.x;
        const uint32_t S = u < n_units ? s_aux[u] : 0u;
        const uint32_t c = __popc(S);
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += t;
        }
        const uint32_t row_off = __shfl_sync(kFull, row_ex, static_cast<int>(it * (kFeatThreads / 32) + warp));
        if (S == 0) continue;                          // (no shuffles below)
        uint32_t li = row_off + inc - c;               // CTA-local index of the unit's first kept token
        const uint32_t pos0 = u * 16;                  // relative to span_lo
        const uint32_t T = s_tok[u + 1] | ((s_tok[u + 2] & 0xFu) << 16);
        const uint32_t E = (T + S) & ~T;               // bit j: a kept token ends in front of byte j (bit 20: it runs on)
        const uint32_t Kb = T & ~(T + S);              // bit j: byte j belongs to a kept token that starts in this unit
        const uint4 v = *reinterpret_cast<const uint4 *>(s_txt + u * 16);
        const uint32_t words[5] = {v.x, v.y, v.z, v.w, *reinterpret_cast<const uint32_t *>(s_txt + u * 16 + 16)};
        uint32_t h = 0x811C9DC5u;
#pragma unroll
        for (int j = 0; j < 20; ++j) {
            const uint32_t b = (words[j >> 2] >> (8 * (j & 3))) & 0xFFu;
            if (j > 0) {
                // predicated stores: nearly every step has SOME lane at a token end, a branch would only add reconvergence
                const uint32_t e = (E >> j) & 1u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.shared.u16 [%1], %2;\n\t@p st.shared.u16 [%3], %4;\n\t}"
                    ::"r"(e), "r"(bkt_s + 2u * li), "h"(static_cast<uint16_t>(h & dim_mask)), "r"(end_s + 2u * li),
                      "h"(static_cast<uint16_t>(pos0 + j))
                    : "memory");
                li += e;
                h = e ? 0x811C9DC5u : h;               // the next kept token starts from the offset basis
            }
            h = ((Kb >> j) & 1u) ? (h ^ b) * 0x01000193u : h;
        }
        if ((E >> 20) & 1u) {
            // the unit's last kept token runs past the look-ahead: finish it from the staged masks / bytes,
            // and past those (a token longer than the rest of the span) from L2
            uint32_t p = pos0 + 20;
            bool ended = false;
            while (p < staged) {
                if (!((s_tok[1 + (p >> 4)] >> (p & 15u)) & 1u)) { ended = true; break; }
                h = (h ^ s_txt[p]) * 0x01000193u;
                ++p;
            }
            if (!ended) {
                while (p < avail_rel) {
                    const uint8_t cb = lower_byte(a.text[span_lo + p]);
                    if (!token_byte(cb)) break;
                    h = (h ^ cb) * 0x01000193u;
                    ++p;
                }
            }
            s_bkt[li] = static_cast<uint16_t>(h & dim_mask);
            s_end[li] = static_cast<uint16_t>(p < 0xFFFFu ? p : 0xFFFFu);
            if (p >= 0xFFFFu) s_long_end = p;         // at most one token of a span can run this far past it
            if (!ended && p >= avail_rel && a.avail_end < a.n) {
                // ran out of copied bytes mid-token (at most one token of a launch does): hash_deferred_kernel
                // finishes it -- parked below, once its ordinal is known
                s_def_li = li;
                s_def_start = pos0 + (31 - __clz(S));  // its start: the unit's last kept start
            }
        }
    }

    // ---- tokens of the CTAs before this one (they have all started: ticket order)
    uint32_t before = 0;
    for (uint32_t j = threadIdx.x; j < cta; j += kFeatThreads) {
        uint64_t v;
        const long long t0 = clock64();
        while (static_cast<uint32_t>((v = ld_relaxed_u64(a.state + j)) >> 32) != a.seq) {
            __nanosleep(200);
            if (clock64() - t0 > (4ll << 30)) {       // ~2 s: fail loudly (the host reports it), never hang
                atomicExch(a.ctl + kCtlStalled, 1u);
                v = 0;
                break;
            }
        }
        before += static_cast<uint32_t>(v);
    }
    before = __reduce_add_sync(kFull, before);
    if (lane == 0) s_red[warp] = before;
    __syncthreads();                                   // (also: every token record is in shared memory)
    uint32_t ex = 0;
#pragma unroll
    for (int w = 0; w < kFeatThreads / 32; ++w) ex += s_red[w];
    if (threadIdx.x == 0 && cta + 1 == gridDim.x) a.ctl[kCtlTokens] = s_base + ex + total;   // every other CTA has read the old value
    const uint32_t ord_base = s_base + ex;
    const uint32_t lo32 = static_cast<uint32_t>(span_lo);             // documents are < 4 GiB

    // ---- place: coalesced records, chunk-window starts, the parked token
    for (uint32_t i = threadIdx.x; i < total; i += kFeatThreads) {
        const uint32_t e = s_end[i];
        a.tok_bucket[ord_base + i] = s_bkt[i];
        a.tok_end[ord_base + i] = lo32 + (e == 0xFFFFu ? s_long_end : e);
    }
    for (uint32_t k = (ord_base + kChunkStride - 1) / kChunkStride + threadIdx.x; k * kChunkStride < ord_base + total; k += kFeatThreads) {
        // token k * 112 opens a chunk window: its start is the first kept start at or behind the end of the token before it
        const uint32_t i = k * kChunkStride - ord_base;
        const uint32_t from = i ? s_end[i - 1] : 0u;
        uint32_t u = from >> 4;
        uint32_t m = s_aux[u] & ~((1u << (from & 15u)) - 1u);
        while (m == 0) m = s_aux[++u];
        a.chunk_start[k] = lo32 + u * 16 + (__ffs(m) - 1);
    }
    if (threadIdx.x == 0 && s_def_li != 0xFFFFFFFFu) {
        const uint32_t slot = atomicAdd(a.ctl + kCtlDeferred, 1u);
        if (slot < kMaxDeferred) {
            a.deferred[2 * slot] = ord_base + s_def_li;
            a.deferred[2 * slot + 1] = lo32 + s_def_start;
        }
    }
}

__global__ void __launch_bounds__(64) hash_deferred_kernel(const TokenizeArgs a, uint32_t count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t ord = a.deferred[2 * i];
    size_t p = a.deferred[2 * i + 1];
    uint32_t h = 0x811C9DC5u;
    while (p < a.n) {
        const uint8_t cb = lower_byte(a.text[p]);
        if (!token_byte(cb)) break;
        h ^= cb;
        h *= 0x01000193u;
        ++p;
    }
    a.tok_bucket[ord] = static_cast<uint16_t>(h & a.dim_mask);
    a.tok_end[ord] = static_cast<uint32_t>(p);
}

constexpr int kRowWarps = 8;

// kM: 256-byte sub-rows per row (dim = 256 * kM).  A window holds at most 128 tokens, so a bucket's count fits
// a byte: the warp's histogram IS the row -- dim byte counters bumped four to a 32-bit word (shared atomicAdd of
// 1 << 8 * (bucket & 3); 128 in one byte cannot carry), read back 8 bytes per lane and sub-row without bank
// conflicts, clamped to 127 (only a count of exactly 128 is above it) and stored as they are.
template <int kM>
__global__ void __launch_bounds__(kRowWarps * 32) rows_from_tokens_kernel(
    const uint16_t *__restrict__ tok_bucket, const uint32_t *__restrict__ chunk_start,
    const uint32_t *__restrict__ tok_end, uint32_t n_tokens, uint32_t n_chunks, int8_t *__restrict__ F,
    int32_t *__restrict__ ff, uint32_t *__restrict__ seg, uint32_t store_seg, int64_t *__restrict__ spans) {
    constexpr int kD = kSubDim * kM;
    static_assert(kChunkTokens <= 128, "a byte counter holds a window's largest count");
    __shared__ __align__(8) uint32_t hist[kRowWarps][kD / 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *h = hist[warp];
    for (uint32_t w = blockIdx.x * kRowWarps + warp; w < n_chunks; w += gridDim.x * kRowWarps) {
#pragma unroll
        for (int sub = 0; sub < kM; ++sub) *reinterpret_cast<uint2 *>(h + sub * (kSubDim / 4) + lane * 2) = make_uint2(0u, 0u);
        __syncwarp();
        const uint32_t lo = w * kChunkStride;
        const uint32_t hi = min(lo + kChunkTokens, n_tokens);
#pragma unroll
        for (int j = 0; j < kChunkTokens / 32; ++j) {
            const uint32_t t = lo + j * 32 + lane;
            if (t < hi) {
                const uint32_t b = tok_bucket[t];
                atomicAdd(&h[b >> 2], 1u << (8 * (b & 3u)));
            }
        }
        __syncwarp();
        int sq = 0;
#pragma unroll
        for (int sub = 0; sub < kM; ++sub) {
            uint2 v = *reinterpret_cast<const uint2 *>(h + sub * (kSubDim / 4) + lane * 2);
            v.x -= (v.x >> 7) & 0x01010101u;          // 128 -> 127
            v.y -= (v.y >> 7) & 0x01010101u;
            sq = __dp4a(static_cast<int>(v.x), static_cast<int>(v.x), sq);
            sq = __dp4a(static_cast<int>(v.y), static_cast<int>(v.y), sq);
            *reinterpret_cast<uint2 *>(F + (static_cast<size_t>(w) * kM + sub) * kSubBytes + lane * 8) = v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(kFull, sq, o);
        if (lane == 0) {
            ff[w] = sq;
            if (seg) seg[w] = store_seg;
            if (spans) {
                spans[2 * static_cast<size_t>(w)] = chunk_start[w];
                spans[2 * static_cast<size_t>(w) + 1] = tok_end[hi - 1];
            }
        }
        __syncwarp();
    }
}

// Query featurisation (RF-1 step 5): no chunking, so no token ordinals are needed -- one block
// histograms every kept token of the (<= 32 KB, routes/chat.py:48) query text.
__global__ void __launch_bounds__(256) featurize_query_kernel(const uint8_t *__restrict__ text, uint32_t n,
                                                              const uint8_t *__restrict__ weights,
                                                              int8_t *__restrict__ q_out, uint32_t dim) {
    __shared__ uint32_t hist[kSubDim * kMaxSub];
    for (uint32_t d = threadIdx.x; d < dim; d += blockDim.x) hist[d] = 0;
    __syncthreads();
    auto at = [&](long long i) -> uint8_t { return (i >= 0 && i < static_cast<long long>(n)) ? lower_byte(text[i]) : 0; };
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        uint8_t w[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) w[j] = at(static_cast<long long>(i) + j - 1);
        if (!kept_start(w + 1)) continue;
        uint32_t h = 0x811C9DC5u;
        for (uint32_t p = i; p < n; ++p) {
            const uint8_t cb = lower_byte(text[p]);
            if (!token_byte(cb)) break;
            h ^= cb;
            h *= 0x01000193u;
        }
        atomicAdd(&hist[h & (dim - 1)], 1u);
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < dim; d += blockDim.x) {
        const uint32_t tf = min(hist[d], 127u);
        q_out[d] = static_cast<int8_t>(weights ? min(tf * weights[d], 127u) : tf);
    }
}

// RF-1w document frequencies (oracle/SPEC.md "IDF-weighted variant").  A warp takes four 256-byte
// sub-rows per step (4 / kM rows); a lane owns eight adjacent buckets of each sub-row position (one
// 8-byte load per sub-row).  Features are counts in [0, 127], so (x + 0x7F7F7F7F) has bit 7 of a byte set
// exactly when that byte is non-zero: the per-byte flags accumulate packed, four buckets per register,
// and are spilled into 32-bit counters every 252 rows.  HBM-bound: dim + 4 B per row, read once.
template <int kM>
__global__ void __launch_bounds__(256) bucket_df_kernel(const __grid_constant__ DfArgs a) {
    constexpr int kD = kSubDim * kM;
    constexpr int kRowsPerStep = 4 / kM;
    __shared__ unsigned int acc[kD + 1];
    for (int i = threadIdx.x; i <= kD; i += blockDim.x) acc[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t total = a.prefix[a.n_ext];
    uint32_t cnt[kM][8];
    uint32_t p0[kM], p1[kM];
#pragma unroll
    for (int m = 0; m < kM; ++m) {
        p0[m] = p1[m] = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) cnt[m][j] = 0;
    }
    uint32_t pending = 0, live = 0;
    auto spill = [&]() {
#pragma unroll
        for (int m = 0; m < kM; ++m) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                cnt[m][j] += (p0[m] >> (8 * j)) & 0xFF;
                cnt[m][4 + j] += (p1[m] >> (8 * j)) & 0xFF;
            }
            p0[m] = p1[m] = 0;
        }
        pending = 0;
    };
    uint32_t e = 0;
    for (uint64_t v0 = static_cast<uint64_t>(warp) * kRowsPerStep; v0 < total; v0 += static_cast<uint64_t>(warps_total) * kRowsPerStep) {
        int2 x[4];
#pragma unroll
        for (int j = 0; j < kRowsPerStep; ++j) {
#pragma unroll
            for (int m = 0; m < kM; ++m) x[j * kM + m] = make_int2(0, 0);
            const uint64_t v = v0 + j;
            if (v >= total) continue;
            while (v >= a.prefix[e + 1]) ++e;
            const uint32_t row = a.lo[e] + static_cast<uint32_t>(v - a.prefix[e]);
            const uint32_t sg = __ldg(a.seg + row);
            bool ok = false;
            for (uint32_t t = 0; t < a.n_scope; ++t) ok |= (a.scope[t] == sg);
            if (!ok || sg == RF_TOMBSTONE) continue;
            ++live;
#pragma unroll
            for (int m = 0; m < kM; ++m)
                x[j * kM + m] = __ldg(reinterpret_cast<const int2 *>(a.F + (static_cast<size_t>(row) * kM + m) * kSubBytes) + lane);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            p0[u % kM] += ((static_cast<uint32_t>(x[u].x) + 0x7F7F7F7Fu) >> 7) & 0x01010101u;
            p1[u % kM] += ((static_cast<uint32_t>(x[u].y) + 0x7F7F7F7Fu) >> 7) & 0x01010101u;
        }
        pending += kRowsPerStep;
        if (pending >= 252) spill();
    }
    spill();
#pragma unroll
    for (int m = 0; m < kM; ++m)
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (cnt[m][j]) atomicAdd(&acc[m * kSubDim + lane * 8 + j], cnt[m][j]);
    if (lane == 0 && live) atomicAdd(&acc[kD], live);
    __syncthreads();
    for (int i = threadIdx.x; i <= kD; i += blockDim.x)
        if (acc[i]) atomicAdd(a.out + i, static_cast<unsigned long long>(acc[i]));
}

__global__ void __launch_bounds__(256) row_meta_kernel(const int8_t *__restrict__ F, uint64_t n_rows, uint32_t m_sub,
                                                       int32_t *__restrict__ ff, uint32_t *__restrict__ seg,
                                                       uint32_t store_seg) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps_total = static_cast<uint64_t>(gridDim.x) * (blockDim.x >> 5);
    for (uint64_t r = static_cast<uint64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows;
         r += warps_total) {
        int sq = 0;
        for (uint32_t m = 0; m < m_sub; ++m) {
            const int2 v = *reinterpret_cast<const int2 *>(F + (r * m_sub + m) * kSubBytes + lane * 8);
            sq = __dp4a(v.x, v.x, sq);
            sq = __dp4a(v.y, v.y, sq);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(kFull, sq, o);
        if (lane == 0) {
            ff[r] = sq;
            if (seg) seg[r] = store_seg;
        }
    }
}

__global__ void __launch_bounds__(256) fill_u32_kernel(uint32_t *__restrict__ p, uint64_t n, uint32_t value) {
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
        p[i] = value;
}

}  // namespace

cudaError_t launch_fill_u32(uint32_t *p, uint64_t n, uint32_t value, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const unsigned blocks = static_cast<unsigned>(std::min<uint64_t>((n + 255) / 256, 148ull * 8ull));
    fill_u32_kernel<<<blocks, 256, 0, s>>>(p, n, value);
    return cudaGetLastError();
}

cudaError_t launch_row_meta(const int8_t *F, uint64_t n_rows, uint32_t dim, int32_t *ff, uint32_t *seg, uint32_t store_seg,
                            cudaStream_t s) {
    if (n_rows == 0) return cudaSuccess;
    uint64_t blocks = (n_rows + 7) / 8;
    if (blocks > 148ull * 8ull) blocks = 148ull * 8ull;
    row_meta_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(F, n_rows, dim / kSubDim, ff, seg, store_seg);
    return cudaGetLastError();
}

cudaError_t launch_tokenize(TokenizeArgs &t, uint32_t blk_first, uint32_t n_blocks_here, cudaStream_t s) {
    if (n_blocks_here == 0) return cudaSuccess;
    // bytes per CTA: small spans keep several generations of CTAs in flight per SM, so one CTA's latencies (ticket,
    // first loads, the counts before it) are covered by its neighbours' hashing (RF_TOKENIZE_SPAN_KB: measurement knob)
    static const uint32_t span_target = [] {
        const char *e = std::getenv("RF_TOKENIZE_SPAN_KB");
        const long kb = e ? std::atol(e) : 0;
        return kb >= 4 && kb * 1024 <= static_cast<long>(kSpanMaxBytes) ? static_cast<uint32_t>(kb) * 1024u : kSpanDefaultBytes;
    }();
    static_assert(span_smem_bytes(kSpanMaxBytes) <= 48u * 1024u, "the span kernel launches without a shared-memory opt-in");
    size_t lo = static_cast<size_t>(blk_first) * kFeatBlockBytes;
    const size_t hi = std::min(t.n, (static_cast<size_t>(blk_first) + n_blocks_here) * kFeatBlockBytes);
    while (lo < hi) {
        // a CTA reads one count word per CTA before it: longer ranges take several launches (the running token
        // count travels in ctl[kCtlTokens])
        const size_t len = std::min(hi - lo, static_cast<size_t>(kSpanMaxCtas) * span_target);
        uint32_t ctas = static_cast<uint32_t>((len + span_target - 1) / span_target);
        const uint32_t span = static_cast<uint32_t>(((len + ctas - 1) / ctas + 15) / 16 * 16);
        ctas = static_cast<uint32_t>((len + span - 1) / span);
        t.byte_lo = lo;
        t.byte_hi = lo + len;
        t.span = span;
        t.seq += 1;
        tokenize_span_kernel<<<ctas, kFeatThreads, span_smem_bytes(span), s>>>(t);
        if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) return e;
        t.ticket_base += ctas;
        lo += len;
    }
    return cudaSuccess;
}

cudaError_t launch_hash_deferred(const TokenizeArgs &a, uint32_t count, cudaStream_t s) {
    if (count == 0) return cudaSuccess;
    hash_deferred_kernel<<<(count + 63) / 64, 64, 0, s>>>(a, count);
    return cudaGetLastError();
}

cudaError_t launch_rows_from_tokens(const TokenizeArgs &w, uint32_t n_tokens, uint32_t n_chunks, int8_t *F, int32_t *ff,
                                    uint32_t *seg, uint32_t store_seg, int64_t *spans_dev, cudaStream_t s) {
    if (n_chunks == 0) return cudaSuccess;
    uint32_t blocks = (n_chunks + kRowWarps - 1) / kRowWarps;
    if (blocks > 148u * 8u) blocks = 148u * 8u;
    auto go = [&](auto kern) {
        kern<<<blocks, kRowWarps * 32, 0, s>>>(w.tok_bucket, w.chunk_start, w.tok_end, n_tokens, n_chunks, F, ff, seg, store_seg, spans_dev);
    };
    switch (w.dim_mask + 1u) {
        case 256: go(rows_from_tokens_kernel<1>); break;
        case 512: go(rows_from_tokens_kernel<2>); break;
        case 1024: go(rows_from_tokens_kernel<4>); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_featurize_query(const uint8_t *text_dev, uint32_t n_bytes, const uint8_t *weights, int8_t *q_out, uint32_t dim,
                                   cudaStream_t s) {
    if (dim > kSubDim * kMaxSub || (dim & (dim - 1))) return cudaErrorInvalidValue;
    featurize_query_kernel<<<1, 256, 0, s>>>(text_dev, n_bytes, weights, q_out, dim);
    return cudaGetLastError();
}

cudaError_t launch_bucket_df(const DfArgs &a, int sm_count, cudaStream_t s) {
    const uint32_t total = a.prefix[a.n_ext];
    if (total == 0) return cudaSuccess;
    const uint32_t m = a.dim / kSubDim;
    const uint32_t want = (total * m + 8 * 4 * 16 - 1) / (8 * 4 * 16);   // >= 16 steps per warp before another block pays off
    const uint32_t blocks = std::max(1u, std::min(want, static_cast<uint32_t>(sm_count) * 8u));
    switch (m) {
        case 1: bucket_df_kernel<1><<<blocks, 256, 0, s>>>(a); break;
        case 2: bucket_df_kernel<2><<<blocks, 256, 0, s>>>(a); break;
        case 4: bucket_df_kernel<4><<<blocks, 256, 0, s>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace rf
