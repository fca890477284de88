// Host-visible declarations shared between the engine (engine.cu) and the kernel translation units.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/rf_b200.h"

namespace rf {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): the attribute is per
// device, and one process may own engines on several GPUs.  Keyed by the function's address (kernels
// of one signature share a C++ type).  Benign if two threads race.
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kern, int bytes) {
    struct Seen { const void *fn; int dev; };
    static Seen seen[256];
    static int n_seen = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const void *fn = reinterpret_cast<const void *>(kern);
    const int n = n_seen;
    for (int i = 0; i < n; ++i)
        if (seen[i].fn == fn && seen[i].dev == dev) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && n < 256) {
        seen[n] = {fn, dev};
        n_seen = n + 1;
    }
    return e;
}

constexpr uint32_t kScanTileSubRows = 32;  // 256-byte sub-rows per warp step of the scan kernel (== rf::kTileSubRows)
// rows per scan tile for rows of `dim` features (dim = 256 * m): 32 / m
inline constexpr uint32_t scan_tile_rows(uint32_t dim) { return kScanTileSubRows * 256u / dim; }

// One query's scan plan (device memory).  The rows to scan are the union of `n_ext` extents
// [ext_lo, ext_hi); extents are a SUPERSET hint of the scoped stores' rows -- the per-row store
// segment word is always checked against `scope`, so correctness never depends on the extents.
struct ScanPlan {
    uint32_t ext_off;      // first extent of this query in the flat extent arrays
    uint32_t n_ext;
    uint32_t total_tiles;  // sum over extents of ceil(rows / scan_tile_rows(dim))
    uint32_t n_scope;
    uint32_t scope[RF_SCOPE_MAX];
};

// Device-resident store table (one entry per store segment, rebuilt by the engine when extents change):
// a batch of store-scoped queries then needs only its scope lists on the device -- the kernel gathers each
// query's extents from the table itself, and the host builds no per-query plan.
struct StoreEntry {
    uint32_t ext_off;       // first extent of the store in the table's flat lo / hi arrays
    uint32_t n_ext;         // 0 for an empty or dropped store (at most kInlineExt: more are coalesced)
    uint32_t total_tiles;   // sum over extents of ceil(rows / 32)
    uint32_t reserved;
};

struct ScanArgs {
    const int8_t *F;            // [rows, dim] int8, row-major (dim = 256 * m: the kernel's template parameter)
    const uint32_t *seg;        // [rows] store segment word (0xFFFFFFFF = tombstone)
    const int32_t *ff;          // [rows] sum of squares (read for the k winners only); may be null
    const int8_t *q;            // [nq, dim] query vectors (device)
    const uint32_t *q_index;    // null, or [nq]: launch query qi reads row q_index[qi] of q (and publishes under that number)
    const ScanPlan *plans;      // [nq]
    const uint32_t *ext_lo;     // flat extents
    const uint32_t *ext_hi;
    const uint32_t *ext_tile0;  // exclusive prefix of tile counts within the query, n_ext + 1 entries per query (offset ext_off + qi)
    uint64_t *partial;          // [nq, gridDim.x, k] per-block top-k keys
    uint32_t *tickets;          // [nq] zero-initialised; reset to zero by the finishing block
    uint64_t *floors;           // [nq] zero-initialised; per-query shared lower bound (atomicMax), reset likewise
    uint32_t *tile_ctr;         // [nq] zero-initialised; per-query work-stealing tile counter, reset likewise
    uint64_t *out_keys;         // [nq, k]
    uint64_t *out_ids;          // [nq, k] or null
    int32_t *out_scores;        // [nq, k] or null
    float *out_cos;             // [nq, k] or null
    uint32_t *out_counts;       // [nq] or null
    // plans from the store table (st_tbl != null; plans / ext_* are then unused): query qi's scope is
    // q_segs[q_seg_off[qi] .. q_seg_off[qi + 1]) (at most RF_SCOPE_MAX stores, at most kInlineExt extents in all)
    const StoreEntry *st_tbl;
    const uint32_t *st_lo, *st_hi;
    const uint32_t *q_seg_off, *q_segs;
    uint32_t st_n_stores;
    uint32_t id_base;           // global id of row 0
    uint32_t k;
    uint32_t shared_plan;       // 1: every query uses plans[0] (one scope for the whole batch)
    // Inline copy of plans[0] and its extents (the engine caps a plan at kInlineExt of them): used when
    // inline_plan != 0, which requires shared_plan or nq == 1.  Kernel parameters live in the constant
    // bank, so a single-scope search needs no device-resident plan at all (nothing to cache or evict).
    uint32_t inline_plan;
    ScanPlan plan0;
    uint32_t inl_lo[64];
    uint32_t inl_hi[64];
    uint32_t inl_tile0[65];
    // Fused top-k exchange over NVLink peer memory (sharded search): when px_world > 1 the block
    // that finishes a query stores its k keys into every rank's gather buffer and then releases
    // a per-(rank, query) flag there; merge_wait_kernel on each rank acquires the flags and merges.
    uint64_t *px_keys[8];          // rank r's gather buffer [world][px_nq_cap][k] (peer-mapped pointers)
    uint32_t *px_flags[8];         // rank r's flags [world][px_nq_cap]
    uint32_t px_rank, px_world, px_seq, px_nq_cap;
    uint32_t px_publish_only;      // 1: store + release only; launch_merge_wait acquires and merges (large batches: a block
                                   //    that waits for peers must not hold an SM slot the scan still needs)
    uint64_t *px_out;              // [nq, k] merged result of all ranks (this rank's copy)
    uint32_t *px_timeout;          // set to 1 if a peer's keys did not arrive within the wait bound
    uint32_t *done_flag;           // mapped host word: set to done_seq after the results are visible to the host (or null)
    uint32_t *done_count;          // device word (zero between launches): queries of this launch finished so far
    uint32_t done_seq;
    alignas(16) int8_t q_inline[1024];   // the query vector itself (dim bytes) when q == nullptr (nq == 1, host-side searches)
    uint32_t dbg_flags;            // diagnostics (RF_SCAN_DBG): 1 = no seg bulk copy, 2 = static round-robin tiles instead of stealing
    unsigned long long *debug_ts;  // diagnostics (RF_SCAN_DEBUG=1): [grid.x][8] globaltimer stamps, else null
};
constexpr uint32_t kInlineExt = 64;

// launchers (each returns the cudaError_t of the launch)
enum : int {                    // scan kernel variants (RF_SCAN_VARIANT env, default tma 8x24)
    kScanVariantLdg = 0,        // direct 128-bit global loads, 8 warps, 2 blocks/SM
    kScanVariantTma8x24 = 1,    // producer warp + 8 consumer warps, 24-stage (192 KB) bulk-copy ring
    kScanVariantTma12x24 = 2,
    kScanVariantTma8x16 = 3,
    kScanVariantTma6x12 = 4,    // 96 KB ring: two blocks per SM
    kScanVariantTma12x12 = 5,
    kScanVariantTma4x12 = 6,
    kScanVariantTma4x8 = 7,     // 64 KB ring: three blocks per SM
    kScanVariantCount = 8
};
// `overlap`: launch with programmatic stream serialisation (the kernel's scan phase may start while the
// previous kernel in the stream is still in its merge tail).  Everything the kernel reads before its
// griddepcontrol.wait -- the query vectors and the plan -- must then be complete before the PREVIOUS
// kernel in the stream started: true for kernel parameters and for buffers written by a copy, not for a
// query vector another kernel has just produced.
// dim: 256, 512 or 1024 (wider rows run the default ring variant only)
cudaError_t launch_score_topk_scan(const ScanArgs &a, uint32_t nq, uint32_t blocks_per_query, int variant, uint32_t dim, cudaStream_t s,
                                   bool overlap);
// packed keys [nq, k] -> ids / scores / cosines / counts (host-visible result layout); q: [nq, dim] device
cudaError_t launch_unpack_keys(const uint64_t *keys, const int8_t *q, uint32_t dim, const int32_t *ff, uint32_t id_base, uint32_t nq, uint32_t k,
                               uint64_t *out_ids, int32_t *out_scores, float *out_cos, uint32_t *out_counts, cudaStream_t s);
cudaError_t launch_merge_topk(const uint64_t *keys, uint32_t n_lists, uint32_t nq, uint32_t k,
                              uint64_t *out_keys, cudaStream_t s);
uint32_t scan_default_blocks_per_query(int sm_count, int variant);
// Second half of the publish-only exchange: one warp per query acquires the `world` flags of its query in this
// rank's buffer (bounded wait -> zeroed result + *timeout = 1) and merges the `world` lists into out [nq, k].
// masks: null (wait for every rank) or [nq] owner bit masks (device)
cudaError_t launch_merge_wait(const uint64_t *gather, const uint32_t *flags, const uint8_t *masks, uint32_t world, uint32_t nq_cap,
                              uint32_t nq, uint32_t k, uint32_t seq, uint64_t *out, uint32_t *timeout, cudaStream_t s);

// ---- batched GEMM path (score_topk_gemm.cu) ----------------------------------------------------
constexpr int kGemmListK = 10;      // top-k kept per (thread, query) in registers; searches with k <= 10 qualify
constexpr int kGemmMT = 4;          // 128-query M-tiles resident per block (512 queries)
constexpr int kGemmTileRows = 128;  // chunk rows per B tile
struct GemmArgs {
    const uint32_t *seg;        // [rows] store-segment words
    const int8_t *q;            // [nq, 256] query vectors (device)
    const uint64_t *floors;     // [nq] per-query lower bound keys, or null
    uint64_t *out_lists;        // [n_slices * gemm_lists_per_slice(nq), nq, kGemmListK] sorted lists
    uint32_t scope[RF_SCOPE_MAX];
    uint32_t n_scope;
    uint32_t row_lo, row_hi;    // contiguous row range to score
    uint32_t nq;
    uint32_t id_base;
    uint32_t lists_per_slice;   // = gemm_lists_per_slice(nq)
    uint32_t group_max_mode;    // 1: floor-finding pass -- keep the top-k of per-32-chunk group maxima, not of chunks
    unsigned long long *debug;  // diagnostics: [block][8] cycle counters, or null
    uint32_t dbg_mode;          // diagnostics (RF_GEMM_DBG, debug instantiation of the pair kernel only; results are WRONG): the epilogue
                                // 1 = hands accumulators straight back (tensor-pipe pace alone), 2 = reads them but looks at nothing,
                                // 3 = + the maximum trees, 4 = + the threshold vote
};
// accumulator replicas per M-tile when a block holds fewer than four M-tiles (1 -> 4, 2 -> 2, else 1)
__host__ __device__ inline uint32_t gemm_replicas(uint32_t m_tiles) { return m_tiles == 1 ? 4u : m_tiles == 2 ? 2u : 1u; }
uint32_t gemm_lists_per_slice(uint32_t nq);   // lists each slice writes per query (= replicas)
size_t gemm_lists_bytes(uint32_t n_slices, uint32_t nq);
cudaError_t launch_score_topk_gemm(const GemmArgs &a, const int8_t *q_dev, const int8_t *F, uint64_t f_rows, uint32_t n_slices,
                                   cudaStream_t s);
// CTA-pair variant (score_topk_gemm_pair.cu, tcgen05 cta_group::2): 256-row chunk tiles, 512 queries per pair,
// kGemmPairLists lists per (slice, query)
constexpr int kGemmPairTileRows = 256;
constexpr int kGemmPairLists = 1;   // the four column-block lists of a CTA are merged in shared memory before they are written
size_t gemm_pair_lists_bytes(uint32_t n_slices, uint32_t nq);
cudaError_t launch_score_topk_gemm_pair(const GemmArgs &a, const int8_t *q_dev, const int8_t *F, uint64_t f_rows, uint32_t n_slices,
                                        cudaStream_t s);
// The pair kernel's floor pass (GemmArgs.group_max_mode = 1) writes gemm_pair_floor_groups(n_slices) group maxima per query
// ([groups][nq] u32, in out_lists); launch_kth_largest turns them into floors[q] = (k-th largest) << 32.
uint32_t gemm_pair_floor_groups(uint32_t n_slices);
cudaError_t launch_kth_largest(const uint32_t *vals, uint32_t n_groups, uint32_t nq, uint32_t k, uint64_t *floors, cudaStream_t s);
// 1024-feature rows (K = 1024): the pair kernel with the K dimension streamed; 256 queries per pair, one list per (slice, query);
// floor pass and k-th largest as the pair kernel's
size_t gemm_wide_lists_bytes(uint32_t n_slices, uint32_t nq);
cudaError_t launch_score_topk_gemm_wide(const GemmArgs &a, const int8_t *q_dev, const int8_t *F, uint64_t f_rows, uint32_t dim, uint32_t n_slices,
                                        cudaStream_t s);
// keys: [n_lists, nq, k_in] sorted lists -> out [nq, k_out] (k_out <= k_in <= 32, n_lists <= 1024)
cudaError_t launch_merge_lists(const uint64_t *keys, uint32_t n_lists, uint32_t nq, uint32_t k_in, uint32_t k_out, uint64_t *out,
                               cudaStream_t s, uint64_t *floors = nullptr, uint32_t k_floor = 0);

// zipf_bucket_dev: uint16[65536], bucket (< dim) of each table entry's token
cudaError_t launch_synth_rows(uint64_t seed, uint64_t start_counter, uint64_t n_rows, const uint16_t *zipf_bucket_dev, uint32_t dim,
                              int8_t *F, int32_t *ff, uint32_t *seg, uint32_t first_seg, uint64_t rows_per_store,
                              cudaStream_t s);

// host memcpy into a pinned staging buffer with non-temporal stores (hostcopy.cpp)
void stage_copy(void *dst, const void *src, size_t n);

// featurisation (featurize.cu)
constexpr uint32_t kFeatBlockBytes = 4096;
constexpr uint32_t kMaxDeferred = 64;    // tokens longer than a whole copy chunk, finished after the last copy
enum : uint32_t { kCtlTicket = 0, kCtlTokens = 1, kCtlDeferred = 2, kCtlStalled = 3, kCtlWords = 4 };   // control words (zeroed per document)
constexpr uint32_t kSpanMaxBytes = 12u << 10;   // most text bytes one tokeniser CTA keeps in shared memory (with its masks and token records: < 48 KB)
constexpr uint32_t kSpanMaxCtas = 4096;         // CTAs per tokeniser launch
struct TokenizeArgs {        // device scratch for one document
    const uint8_t *text;     // [>= n + 64] padded device copy of the document, 16-byte aligned
    size_t n;                // document bytes
    size_t avail_end;        // bytes [0, avail_end) have arrived (== n for the last / only launch)
    uint32_t n_blocks;       // 4 KB blocks of the whole document
    uint64_t *state;         // [n_blocks] per-CTA token counts, tagged with the launch number (zeroed per document)
    uint32_t *ctl;           // [kCtlWords]
    uint16_t *tok_bucket;    // [cap_tokens] bucket = hash & dim_mask
    uint32_t dim_mask;       // dim - 1
    uint32_t *tok_end;       // [cap_tokens]
    uint32_t *chunk_start;   // [cap_tokens / 112 + 2] byte offset of the first token of each chunk window
    uint32_t *deferred;      // [2 * kMaxDeferred] (ordinal, start) of parked tokens
    // one launch (filled in by launch_tokenize): CTA c of the launch takes bytes [byte_lo + c * span, + span) below byte_hi
    size_t byte_lo, byte_hi;
    uint32_t span;           // bytes per CTA, a multiple of 16, <= kSpanMaxBytes
    uint32_t seq;            // launch number within the document, from 1 (tags the state words)
    uint32_t ticket_base;    // CTAs launched for this document before this launch
};
// Tokenise the document's 4 KB blocks [blk_first, blk_first + n_blocks_here): as many launches as the range needs
// (one, unless the range is longer than kSpanMaxCtas spans).  Updates t.seq / t.ticket_base.
cudaError_t launch_tokenize(TokenizeArgs &t, uint32_t blk_first, uint32_t n_blocks_here, cudaStream_t s);
cudaError_t launch_hash_deferred(const TokenizeArgs &a, uint32_t count, cudaStream_t s);
cudaError_t launch_rows_from_tokens(const TokenizeArgs &w, uint32_t n_tokens, uint32_t n_chunks, int8_t *F,
                                    int32_t *ff, uint32_t *seg, uint32_t store_seg, int64_t *spans_dev,
                                    cudaStream_t s);
// weights: [dim] u8 device (RF-1w) or null (plain RF-1)
cudaError_t launch_featurize_query(const uint8_t *text_dev, uint32_t n_bytes, const uint8_t *weights, int8_t *q_out, uint32_t dim,
                                   cudaStream_t s);
// RF-1w statistics: out[d] += rows of the extents that are in scope and have F[row, d] > 0 (d < dim);
// out[dim] += rows in scope.  One streaming pass over the extents' rows.
constexpr uint32_t kDfMaxExtents = 64;
struct DfArgs {
    const int8_t *F;
    const uint32_t *seg;
    unsigned long long *out;            // [dim + 1] device, accumulated into (caller zeroes)
    uint32_t n_ext, n_scope, dim;
    uint32_t scope[RF_SCOPE_MAX];
    uint32_t lo[kDfMaxExtents];
    uint32_t prefix[kDfMaxExtents + 1]; // rows before extent i; prefix[n_ext] = total
};
cudaError_t launch_bucket_df(const DfArgs &a, int sm_count, cudaStream_t s);
// p[0 .. n) = value (segment words of rows that were written while still masked)
cudaError_t launch_fill_u32(uint32_t *p, uint64_t n, uint32_t value, cudaStream_t s);
// ff[r] = sum of squares of row r, seg[r] = store_seg (seg may be null), for rows appended as raw features
cudaError_t launch_row_meta(const int8_t *F, uint64_t n_rows, uint32_t dim, int32_t *ff, uint32_t *seg, uint32_t store_seg,
                            cudaStream_t s);

}  // namespace rf
