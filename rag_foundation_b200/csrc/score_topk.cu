// score_topk_scan: query-vs-chunk scoring fused with per-store top-k selection (sm_100a).
//
// Carries out the retrieval step of GeminiRag.ask_stream (reference
// backend/app/services/gemini_rag.py:517-551; the mock's canned citation is :704-718) as the
// RF-1 spec, steps 6-8 (oracle/SPEC.md).  HBM-bound integer work: 260 algorithmic bytes per chunk
// (256 B int8 features + 4 B store-segment word), one pass, scores never written back.
//
// Common to both variants.  A warp scores a tile of 32 consecutive rows (8 KB) per step: lane l
// holds 16 bytes of each of 16 row pairs (load i covers rows 2i, 2i+1; 512 contiguous bytes per
// warp-wide access), 4 dp4a per 16 bytes against the lane's 16-byte slice of the query held in
// registers, then a 15-shuffle transposing butterfly that leaves lane l with the finished int32
// score of row 2*(l&15) + (l>>4).  The lane checks its row's store-segment word against the
// query scope (tenant mask + tombstones), packs (score, id) into one u64 key and offers it to the
// warp's running top-k, which it enters only above the warp's bar: the larger of the warp's own
// k-th key and a per-query floor shared through global memory (atomicMax; any warp's k-th best key
// is a valid lower bound for the query's top-k).  The first tile is taken with one bitonic sort.
// Warps merge through shared memory, blocks through a per-query partial buffer, and the last
// block to finish (ticket counter) merges the partials and writes ids / scores / cosines: one
// launch per search.
//
// Variant "tma" (default): one producer warp streams tiles into a shared-memory ring with 1-D bulk
// copies (cp.async.bulk, completion on an mbarrier per stage); consumer warps wait on the stage's
// full barrier, read it with conflict-free 128-bit shared loads and release it on the empty
// barrier.  Bytes in flight per SM = ring size, independent of registers and of how long a
// consumer spends in the top-k path.
// Variant "ldg": every warp issues its 16 128-bit streaming global loads itself (L1 no-allocate).
#include "rf_device.cuh"
#include "rf_internal.h"

namespace rf {

namespace {

static_assert(kTileSubRows == static_cast<int>(kScanTileSubRows), "host and device disagree on the tile height");
constexpr int kTileBytes = kTileSubRows * kSubBytes;  // 8192
// Rows of dim = 256 * kM features (kM = 1, 2, 4) are kM consecutive 256-byte sub-rows; a tile is always 32
// sub-rows = 32 / kM whole rows, so the streaming side (ring, bulk copies, shared loads, dp4a, butterfly)
// is the same for every width.  What changes with kM: the lane's query slice(s), log2(kM) more shuffles
// that add a row's sub-row scores, and only every kM-th lane of the lower half-warp offers a key.
template <int kM>
struct RowGeom {
    static_assert(kM == 1 || kM == 2 || kM == 4, "rows are 256, 512 or 1024 features");
    static constexpr int kD = kSubDim * kM;                 // features (= bytes) per row
    static constexpr int kTileRows = kTileSubRows / kM;     // whole rows per tile
    static constexpr int kQ = kM >= 2 ? kM / 2 : 1;         // 16-byte query slices a lane holds
};
constexpr int kChunkTiles = 8;                     // tiles claimed per atomic by a producer lane (tma variant)
constexpr int kMaxExtSmem = 64;                    // extents staged in shared memory (engine caps plans at 64)

// Reduce 16 per-lane partial sums (one per load) across the 16 lanes of each half-warp, leaving
// lane l with the total of partial index (l & 15).  8 + 4 + 2 + 1 shuffles.
__device__ __forceinline__ int transpose_reduce16(int (&p)[16], int lane) {
    int v8[8];
    {
        const bool up = lane & 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int send = up ? p[j] : p[j + 8];
            const int keep = up ? p[j + 8] : p[j];
            v8[j] = keep + __shfl_xor_sync(kFull, send, 8);
        }
    }
    int v4[4];
    {
        const bool up = lane & 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int send = up ? v8[j] : v8[j + 4];
            const int keep = up ? v8[j + 4] : v8[j];
            v4[j] = keep + __shfl_xor_sync(kFull, send, 4);
        }
    }
    int v2[2];
    {
        const bool up = lane & 2;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int send = up ? v4[j] : v4[j + 2];
            const int keep = up ? v4[j + 2] : v4[j];
            v2[j] = keep + __shfl_xor_sync(kFull, send, 2);
        }
    }
    const bool up = lane & 1;
    const int send = up ? v2[0] : v2[1];
    const int keep = up ? v2[1] : v2[0];
    return keep + __shfl_xor_sync(kFull, send, 1);
}

// Load i of a tile covers sub-rows 2i (lanes 0-15) and 2i + 1 (lanes 16-31), i.e. part (2i + (lane >> 4)) % kM
// of a row: the lane multiplies it with slice i % kQ of its query registers (load_query below).
// Returns, on every lane, the score of row row_in_tile<kM>(lane) (valid there on the lanes key_lane<kM>).
template <int kM>
__device__ __forceinline__ int score_tile(const int4 (&x)[16], const int4 (&qv)[RowGeom<kM>::kQ], int lane) {
    constexpr int kQ = RowGeom<kM>::kQ;
    int p[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int4 &w = qv[i % kQ];
        int acc = __dp4a(x[i].x, w.x, 0);
        acc = __dp4a(x[i].y, w.y, acc);
        acc = __dp4a(x[i].z, w.z, acc);
        p[i] = __dp4a(x[i].w, w.w, acc);
    }
    int s = transpose_reduce16(p, lane);       // score of sub-row 2 * (lane & 15) + (lane >> 4)
    if (kM >= 2) s += __shfl_xor_sync(kFull, s, 16);
#pragma unroll
    for (int o = 1; o < kM / 2; o <<= 1) s += __shfl_xor_sync(kFull, s, o);
    return s;
}
template <int kM>
__device__ __forceinline__ int row_in_tile(int lane) {
    return kM == 1 ? 2 * (lane & 15) + (lane >> 4) : (lane & 15) / (kM / 2 > 0 ? kM / 2 : 1);
}
template <int kM>
__device__ __forceinline__ bool key_lane(int lane) {
    return kM == 1 ? true : (lane < 16 && ((lane & 15) % (kM / 2 > 0 ? kM / 2 : 1)) == 0);
}
// the lane's slice(s) of query vector q (RowGeom<kM>::kD bytes)
template <int kM>
__device__ __forceinline__ void load_query(const int8_t *q, int lane, int4 (&qv)[RowGeom<kM>::kQ]) {
#pragma unroll
    for (int j = 0; j < RowGeom<kM>::kQ; ++j) {
        const int part = kM == 1 ? 0 : 2 * j + (lane >> 4);
        qv[j] = *reinterpret_cast<const int4 *>(q + part * kSubBytes + (lane & 15) * 16);
    }
}
// this lane's share of ||q||^2 (summed over the warp it covers every query byte once)
template <int kM>
__device__ __forceinline__ int query_sq_lane(const int4 (&qv)[RowGeom<kM>::kQ], int lane) {
    int qq = 0;
#pragma unroll
    for (int j = 0; j < RowGeom<kM>::kQ; ++j) {
        qq = __dp4a(qv[j].x, qv[j].x, qq);
        qq = __dp4a(qv[j].y, qv[j].y, qq);
        qq = __dp4a(qv[j].z, qv[j].z, qq);
        qq = __dp4a(qv[j].w, qv[j].w, qq);
    }
    return (kM == 1 && lane >= 16) ? 0 : qq;
}

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// diagnostics: slot 0 entry, 1 plan staged, 2 first tile scored, 3 scan loop done (max over warps),
// 4 block merged + partial published, 5 last block done
__device__ __forceinline__ void stamp(const ScanArgs &a, int slot) {
    if (a.debug_ts) a.debug_ts[(static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 8 + slot] = gtime();
}
__device__ __forceinline__ void stamp_max(const ScanArgs &a, int slot) {
    if (a.debug_ts) atomicMax(a.debug_ts + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 8 + slot, gtime());
}

// Per-block view of one query's plan, staged in shared memory.
struct BlockPlan {
    uint32_t lo[kMaxExtSmem];
    uint32_t hi[kMaxExtSmem];
    uint32_t tile0[kMaxExtSmem + 1];
    uint32_t scope[RF_SCOPE_MAX];
    uint32_t n_ext, n_scope, t_lo, t_hi, total_tiles;
};

// Plan of query qi gathered from the device-resident store table: the extents of every store in the query's
// scope, concatenated (stores own disjoint rows, so the order does not matter: tiles are located through the
// prefix array).  Warp 0 does it; ends with __syncthreads().
template <int kTileRows>
__device__ __forceinline__ void stage_plan_from_table(const ScanArgs &a, int qi, BlockPlan &bp) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const uint32_t s0 = a.q_seg_off[qi], s1 = a.q_seg_off[qi + 1];
        const uint32_t n_scope = min(s1 - s0, static_cast<uint32_t>(RF_SCOPE_MAX));
        if (lane < RF_SCOPE_MAX) bp.scope[lane] = lane < static_cast<int>(n_scope) ? a.q_segs[s0 + lane] : kTombstone;
        uint32_t cnt = 0;
        for (uint32_t j = 0; j < n_scope; ++j) {
            const uint32_t sg = a.q_segs[s0 + j];
            bool dup = false;
            for (uint32_t i = 0; i < j; ++i) dup |= (a.q_segs[s0 + i] == sg);
            if (dup || sg >= a.st_n_stores) continue;
            const StoreEntry en = a.st_tbl[sg];
            const uint32_t take = min(en.n_ext, static_cast<uint32_t>(kMaxExtSmem) - cnt);
            for (uint32_t i = lane; i < take; i += 32) {
                bp.lo[cnt + i] = a.st_lo[en.ext_off + i];
                bp.hi[cnt + i] = a.st_hi[en.ext_off + i];
            }
            cnt += take;
        }
        __syncwarp();
        // exclusive prefix of the tile counts over <= 64 extents: two per lane
        const uint32_t ta = lane < static_cast<int>(cnt) ? (bp.hi[lane] - bp.lo[lane] + kTileRows - 1) / kTileRows : 0u;
        const uint32_t tb = lane + 32 < static_cast<int>(cnt) ? (bp.hi[lane + 32] - bp.lo[lane + 32] + kTileRows - 1) / kTileRows : 0u;
        uint32_t ia = ta, ib = tb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t xa = __shfl_up_sync(kFull, ia, o), xb = __shfl_up_sync(kFull, ib, o);
            if (lane >= o) { ia += xa; ib += xb; }
        }
        const uint32_t total_a = __shfl_sync(kFull, ia, 31), total = total_a + __shfl_sync(kFull, ib, 31);
        bp.tile0[lane] = ia - ta;
        bp.tile0[lane + 32] = total_a + ib - tb;
        if (lane == 0) {
            bp.tile0[kMaxExtSmem] = total;
            bp.n_ext = cnt;
            bp.total_tiles = total;
            bp.n_scope = n_scope;
            bp.t_lo = static_cast<uint32_t>(static_cast<uint64_t>(total) * blockIdx.x / gridDim.x);
            bp.t_hi = static_cast<uint32_t>(static_cast<uint64_t>(total) * (blockIdx.x + 1) / gridDim.x);
        }
        __syncwarp();
        if (lane == 0 && cnt < static_cast<uint32_t>(kMaxExtSmem)) bp.tile0[cnt] = total;   // the sentinel the tile cursor stops at
    }
    __syncthreads();
}

template <int kTileRows>
__device__ __forceinline__ void stage_plan(const ScanArgs &a, int qi, BlockPlan &bp) {
    if (a.st_tbl) {
        stage_plan_from_table<kTileRows>(a, qi, bp);
        return;
    }
    if (a.inline_plan) {
        // single query / shared scope with few extents: the plan rides in the kernel parameters
        // (constant bank), so no dependent global loads stand before the first feature load
        const uint32_t n_ext = a.plan0.n_ext;
        if (threadIdx.x < n_ext) {
            bp.lo[threadIdx.x] = a.inl_lo[threadIdx.x];
            bp.hi[threadIdx.x] = a.inl_hi[threadIdx.x];
        }
        if (threadIdx.x <= n_ext) bp.tile0[threadIdx.x] = a.inl_tile0[threadIdx.x];
        if (threadIdx.x < RF_SCOPE_MAX) bp.scope[threadIdx.x] = threadIdx.x < a.plan0.n_scope ? a.plan0.scope[threadIdx.x] : kTombstone;
        if (threadIdx.x == 0) {
            const uint32_t total = a.plan0.total_tiles;
            bp.n_ext = n_ext;
            bp.total_tiles = total;
            bp.n_scope = a.plan0.n_scope;
            bp.t_lo = static_cast<uint32_t>(static_cast<uint64_t>(total) * blockIdx.x / gridDim.x);
            bp.t_hi = static_cast<uint32_t>(static_cast<uint64_t>(total) * (blockIdx.x + 1) / gridDim.x);
        }
        __syncthreads();
        return;
    }
    const ScanPlan &plan = a.plans[a.shared_plan ? 0 : qi];
    const uint32_t n_ext = min(plan.n_ext, static_cast<uint32_t>(kMaxExtSmem));
    const uint32_t *g_lo = a.ext_lo + plan.ext_off;
    const uint32_t *g_hi = a.ext_hi + plan.ext_off;
    const uint32_t *g_t0 = a.ext_tile0 + plan.ext_off + (a.shared_plan ? 0 : qi);
    for (uint32_t i = threadIdx.x; i < n_ext; i += blockDim.x) {
        bp.lo[i] = g_lo[i];
        bp.hi[i] = g_hi[i];
    }
    for (uint32_t i = threadIdx.x; i <= n_ext; i += blockDim.x) bp.tile0[i] = g_t0[i];
    if (threadIdx.x < RF_SCOPE_MAX) bp.scope[threadIdx.x] = threadIdx.x < plan.n_scope ? plan.scope[threadIdx.x] : kTombstone;
    if (threadIdx.x == 0) {
        const uint32_t total = plan.total_tiles;
        bp.n_ext = n_ext;
        bp.total_tiles = total;
        bp.n_scope = plan.n_scope;
        bp.t_lo = static_cast<uint32_t>(static_cast<uint64_t>(total) * blockIdx.x / gridDim.x);
        bp.t_hi = static_cast<uint32_t>(static_cast<uint64_t>(total) * (blockIdx.x + 1) / gridDim.x);
    }
    __syncthreads();
}

// Extent cursor: tiles are visited in increasing order, so `e` only moves forward.
struct TileCursor {
    uint32_t e;
    __device__ __forceinline__ void seek(const BlockPlan &bp, uint32_t t) {
        uint32_t lo = 0, hi = bp.n_ext;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (bp.tile0[mid] <= t) lo = mid; else hi = mid;
        }
        e = lo;
    }
    template <int kTileRows>
    __device__ __forceinline__ void locate(const BlockPlan &bp, uint32_t t, uint32_t &row0, uint32_t &row_end) {
        while (t >= bp.tile0[e + 1]) ++e;
        row0 = bp.lo[e] + (t - bp.tile0[e]) * kTileRows;
        row_end = bp.hi[e];
    }
};

// Random access (work-stealing order): binary search over the <= 64 extents.
template <int kTileRows>
__device__ __forceinline__ void locate_tile(const BlockPlan &bp, uint32_t t, uint32_t &row0, uint32_t &row_end) {
    uint32_t lo = 0, hi = bp.n_ext;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (bp.tile0[mid] <= t) lo = mid; else hi = mid;
    }
    row0 = bp.lo[lo] + (t - bp.tile0[lo]) * kTileRows;
    row_end = bp.hi[lo];
}

__device__ __forceinline__ bool in_scope(const BlockPlan &bp, uint32_t seg) {
    bool ok = false;
    if (seg != kTombstone) {
        for (uint32_t j = 0; j < bp.n_scope; ++j) ok |= (seg == bp.scope[j]);
    }
    return ok;
}

// Offer one tile's keys to the warp list; share / learn the per-query floor.
__device__ __forceinline__ void offer_tile(WarpTopK &top, bool &first, uint64_t key, uint64_t floor_seen, uint64_t *g_floor,
                                           int k, int lane) {
    if (floor_seen > top.floor) top.floor = floor_seen;
    const uint64_t before = top.thr;
    if (first) {
        top.init_sorted(key, k, lane);
        first = false;
    } else {
        top.consume(key, k, lane);
    }
    if (top.thr > before && top.thr > top.floor) {  // warp-uniform
        if (lane == 0) atomicMax(reinterpret_cast<unsigned long long *>(g_floor), static_cast<unsigned long long>(top.thr));
    }
}

// ---- merges -----------------------------------------------------------------------------------
// Every list that is merged here is sorted descending and zero padded, so a merge is a tournament:
// lane l holds the head of list l; k rounds of a two-step warp max (score word, then id word among
// the lanes that tie on the score) pop the winners in rank order.  Cost is k rounds regardless of
// ties, instead of one serial insertion per candidate.
constexpr int kListCap = RF_TOPK_MAX;                      // keys per list slot
constexpr int kWarpArea = 32 * kListCap;                   // one warp's staging area: 32 lists
template <int kWarps>
struct MergeScratch {                                      // lives in dynamic shared memory
    static_assert(kWarps <= 32, "level 2 merges one list per lane");
    static constexpr int kL1Warps = kWarps < 8 ? kWarps : 8;   // warps that run level-1 tournaments
    uint64_t warp_area[kL1Warps][kWarpArea];               // level 1: 32 lists x k per warp
    uint64_t level2[32 * kListCap];                        // level 2: up to 32 merged lists
    uint32_t flag;
};

// lists[l * stride + j]: list l (l < n_lists <= 32), j < k.  Returns this lane's rank-th winner
// (rank = lane, 0 beyond k or when the lists run dry).
__device__ __forceinline__ uint64_t warp_tournament(const uint64_t *lists, int n_lists, int stride, int k, int lane) {
    int pos = 0;
    uint64_t head = lane < n_lists ? lists[lane * stride] : 0ull;
    uint64_t result = 0ull;
    for (int r = 0; r < k; ++r) {
        const uint32_t hi = static_cast<uint32_t>(head >> 32);
        const uint32_t lo = static_cast<uint32_t>(head);
        const uint32_t mhi = __reduce_max_sync(kFull, hi);
        const uint32_t mlo = __reduce_max_sync(kFull, hi == mhi ? lo : 0u);
        const uint64_t win = (static_cast<uint64_t>(mhi) << 32) | mlo;
        if (win == 0ull) break;                            // warp-uniform: nothing left anywhere
        if (lane == r) result = win;
        if (head == win) {                                 // duplicates across lists advance together
            ++pos;
            head = pos < k ? lists[lane * stride + pos] : 0ull;
        }
    }
    return result;
}

__device__ __forceinline__ void set_list(WarpTopK &top, uint64_t mine, int k) {
    top.mine = mine;
    top.thr = shfl_u64(mine, k - 1);
    top.floor = 0;
}

// Warps -> warp 0.  On return warp 0's `top` holds the block's top-k.
template <int kWarps>
__device__ __forceinline__ void block_merge(WarpTopK &top, MergeScratch<kWarps> &ms, int k, int warp, int lane) {
    constexpr int n_warps = kWarps;
    __syncthreads();
    if (lane < k) ms.level2[warp * k + lane] = top.mine;
    __syncthreads();
    if (warp == 0) set_list(top, warp_tournament(ms.level2, n_warps, k, k, lane), k);
}

// Block top-k -> partial buffer; the last block of the query merges all partials and writes the
// answer.  Called by every thread of the block.
// `n_blocks`: blocks of the launch that take part for this query (gridDim.x, or 1 when the query has nothing
// to scan here and only block 0 stays to write the empty answer).
template <int kWarps>
__device__ __forceinline__ void finish_query(const ScanArgs &a, int qi, WarpTopK &top, int qq_lane, MergeScratch<kWarps> &ms,
                                             int k, int warp, int lane, uint32_t n_blocks) {
    constexpr int n_warps = kWarps;
    block_merge(top, ms, k, warp, lane);
    uint64_t *part = a.partial + (static_cast<size_t>(qi) * gridDim.x) * k;
    if (warp == 0) {
        if (lane < k) __stcg(part + static_cast<size_t>(blockIdx.x) * k + lane, top.mine);
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            const uint32_t ticket = atomicAdd(a.tickets + qi, 1u);
            ms.flag = (ticket == n_blocks - 1) ? 1u : 0u;
            stamp(a, 4);
        }
    }
    __syncthreads();
    if (!ms.flag) return;

    // ---- last block of this query: two tournament levels over the gridDim.x partial lists
    __threadfence();
    const int n_lists = static_cast<int>(n_blocks);        // engine keeps this <= 1024
    const int n_groups = (n_lists + 31) / 32;
    constexpr int l1_warps = MergeScratch<kWarps>::kL1Warps;
    const int total_keys = n_lists * k;
    if (total_keys <= l1_warps * kWarpArea) {
        // All partial lists fit the level-1 scratch: the whole block copies them in ONE sweep (every
        // thread has its loads in flight together -- this is the tail of a lone query's latency), then
        // the warps play the level-1 tournaments out of shared memory.
        uint64_t *flat = &ms.warp_area[0][0];               // list l at flat[l * k]
        constexpr int kThreads = kWarps * 32;
        for (int base = static_cast<int>(threadIdx.x); base < total_keys; base += kThreads * 8) {
            uint64_t t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * kThreads;
                t[u] = idx < total_keys ? __ldcg(part + idx) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * kThreads;
                if (idx < total_keys) flat[idx] = t[u];
            }
        }
        __syncthreads();
        for (int g = warp; g < n_groups; g += n_warps) {
            const int lists_here = min(32, n_lists - g * 32);
            const uint64_t w = warp_tournament(flat + static_cast<size_t>(g) * 32 * k, lists_here, k, k, lane);
            if (lane < k) ms.level2[g * k + lane] = w;
        }
    } else {
        for (int g = warp; g < n_groups && warp < l1_warps; g += l1_warps) {
            const int lists_here = min(32, n_lists - g * 32);
            const uint64_t *src = part + static_cast<size_t>(g) * 32 * k;
            uint64_t *area = ms.warp_area[warp];
            for (int i = lane; i < lists_here * k; i += 32) area[i] = __ldcg(src + i);
            __syncwarp();
            const uint64_t w = warp_tournament(area, lists_here, k, k, lane);
            if (lane < k) ms.level2[g * k + lane] = w;
            __syncwarp();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) stamp(a, 6);
    if (warp != 0) return;
    set_list(top, warp_tournament(ms.level2, n_groups, k, k, lane), k);

    // ||q||^2 for the reported cosine (the lanes' shares cover the query bytes once: query_sq_lane)
    int qq = qq_lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(kFull, qq, o);

    const uint64_t key = lane < k ? top.mine : 0ull;
    const unsigned found = __ballot_sync(kFull, key != 0ull);
    if (lane < k) {
        const size_t o = static_cast<size_t>(qi) * k + lane;
        a.out_keys[o] = key;
        const uint32_t gid = key_gid(key);
        const int32_t sc = key_score(key);
        if (a.out_ids) a.out_ids[o] = key ? static_cast<uint64_t>(gid) : ~0ull;
        if (a.out_scores) a.out_scores[o] = key ? sc : 0;
        if (a.out_cos) {
            float c = 0.0f;
            if (key && a.ff) {
                const float nq = sqrtf(static_cast<float>(qq));
                const float nf = sqrtf(static_cast<float>(__ldg(a.ff + (gid - a.id_base))));
                const float den = __fmul_rn(nq, nf);
                c = den == 0.0f ? 0.0f : __fdiv_rn(static_cast<float>(sc), den);
            }
            a.out_cos[o] = c;
        }
    }
    if (a.px_world > 1) {
        // ---- fused exchange (sharded search): compute + collective in this one kernel.  This
        // rank's top-k of query qi goes straight into every rank's gather buffer over NVLink
        // (peer-mapped stores), one release store per peer publishes it; then lane r acquires rank
        // r's flag (the peers' scan kernels run concurrently on their own GPUs; the wait is
        // bounded) and the warp plays the tournament over the `world` lists in the local buffer.
        const uint32_t gq = a.q_index ? a.q_index[qi] : static_cast<uint32_t>(qi);    // the query's number in the whole batch
        if (lane < k) {
            const size_t slot = (static_cast<size_t>(a.px_rank) * a.px_nq_cap + gq) * k + lane;
            for (uint32_t p = 0; p < a.px_world; ++p) a.px_keys[p][slot] = key;
        }
        __threadfence_system();
        __syncwarp();
        if (lane < static_cast<int>(a.px_world)) {
            uint32_t *flag = a.px_flags[lane] + static_cast<size_t>(a.px_rank) * a.px_nq_cap + gq;
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(a.px_seq) : "memory");
        }
        if (a.px_publish_only) goto px_done;     // merge_wait_kernel (next in the stream) acquires and merges
        {
        bool arrived = true;
        if (lane < static_cast<int>(a.px_world)) {
            const uint32_t *flag = a.px_flags[a.px_rank] + static_cast<size_t>(lane) * a.px_nq_cap + gq;
            const long long t0 = clock64();
            uint32_t v;
            while (true) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                if (v == a.px_seq) break;
                if (clock64() - t0 > (4ll << 30)) { arrived = false; break; }   // ~2 s: fail loudly, never hang
            }
        }
        arrived = __all_sync(kFull, arrived);
        uint64_t *lists = ms.level2;
        const uint64_t *gather = a.px_keys[a.px_rank];
        for (int i = lane; i < static_cast<int>(a.px_world) * k; i += 32) {
            const int r = i / k, j = i % k;
            lists[i] = __ldcg(gather + (static_cast<size_t>(r) * a.px_nq_cap + gq) * k + j);   // peers wrote into L2
        }
        __syncwarp();
        const uint64_t merged = warp_tournament(lists, static_cast<int>(a.px_world), k, k, lane);
        if (lane < k) a.px_out[static_cast<size_t>(qi) * k + lane] = arrived ? merged : 0ull;
        if (!arrived && lane == 0) atomicExch(a.px_timeout, 1u);
        }
    }
px_done:
    if (lane == 0) {
        if (a.out_counts) a.out_counts[qi] = __popc(found);
        a.tickets[qi] = 0;  // ready for the next launch that uses this sync set
        a.floors[qi] = 0;
        a.tile_ctr[qi] = 0;
        stamp(a, 5);
    }
    if (a.done_flag) {   // host-side searches spin on this word instead of a stream synchronise
        __threadfence_system();   // this query's results are visible to the host
        __syncwarp();
        if (lane == 0) {
            if (gridDim.y == 1) {
                // one query per launch (the single-caller path): the fence above already ordered its
                // results before this store -- no counter, no second fence (~1.7 us of the caller's latency)
                *reinterpret_cast<volatile uint32_t *>(a.done_flag) = a.done_seq;
            } else {
                // the per-launch counter lives in device memory: an atomic on mapped host memory is a
                // PCIe round trip per query and serialises a 1024-query batch
                const uint32_t finished = atomicAdd(a.done_count, 1u) + 1u;
                if (finished == gridDim.y) {
                    *a.done_count = 0;
                    __threadfence_system();
                    *reinterpret_cast<volatile uint32_t *>(a.done_flag) = a.done_seq;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Variant "ldg": direct global loads.
// ------------------------------------------------------------------------------------------------
constexpr int kLdgWarps = 8;

__global__ void __launch_bounds__(kLdgWarps * 32, 2) score_topk_scan_ldg_kernel(const ScanArgs a) {
    __shared__ BlockPlan bp;
    extern __shared__ __align__(16) uint8_t ldg_dyn_smem[];
    MergeScratch<kLdgWarps> &ms = *reinterpret_cast<MergeScratch<kLdgWarps> *>(ldg_dyn_smem);

    const int qi = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int k = static_cast<int>(a.k);
    if (threadIdx.x == 0) stamp(a, 0);
    constexpr int kTileRows = kTileSubRows;     // this variant exists for 256-feature rows only
    constexpr int kRowBytes = kSubBytes;
    const uint32_t q_row = a.q_index ? a.q_index[qi] : static_cast<uint32_t>(qi);
    int4 qv[1];
    load_query<1>(a.q ? a.q + static_cast<size_t>(q_row) * kSubDim : a.q_inline, lane, qv);
    stage_plan<kTileRows>(a, qi, bp);
    if (threadIdx.x == 0) stamp(a, 1);
    const int my_row_in_tile = 2 * (lane & 15) + (lane >> 4);
    uint64_t *g_floor = a.floors + qi;

    WarpTopK top;
    top.reset();
    bool first = true;
    TileCursor cur;
    if (bp.t_lo + warp < bp.t_hi) cur.seek(bp, bp.t_lo + warp);

    for (uint32_t t = bp.t_lo + warp; t < bp.t_hi; t += kLdgWarps) {
        uint32_t row0, row_end;
        cur.locate<kTileRows>(bp, t, row0, row_end);
        const int4 *src = reinterpret_cast<const int4 *>(a.F + static_cast<size_t>(row0) * kRowBytes) + lane;
        int4 x[16];
        uint32_t seg;
        const uint32_t my_row = row0 + my_row_in_tile;
        if (row0 + kTileRows <= row_end) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = ld_stream_v4(src + i * 32);
            seg = __ldg(a.seg + my_row);
        } else {  // ragged last tile of an extent
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint32_t r = row0 + 2 * i + (lane >> 4);
                x[i] = r < row_end ? ld_stream_v4(src + i * 32) : make_int4(0, 0, 0, 0);
            }
            seg = my_row < row_end ? __ldg(a.seg + my_row) : kTombstone;
        }
        const uint64_t floor_seen = __ldcg(g_floor);
        const int score = score_tile<1>(x, qv, lane);
        const uint64_t key = in_scope(bp, seg) ? pack_key(score, a.id_base + my_row) : 0ull;
        if (first && threadIdx.x == 0) stamp(a, 2);
        offer_tile(top, first, key, floor_seen, g_floor, k, lane);
    }
    if (lane == 0) stamp_max(a, 3);
    finish_query(a, qi, top, query_sq_lane<1>(qv, lane), ms, k, warp, lane, gridDim.x);
}

// ------------------------------------------------------------------------------------------------
// Variant "tma": producer warp + shared-memory ring filled by bulk copies.
// ------------------------------------------------------------------------------------------------
template <int kConsumers, int kStages>
struct TmaSmem {
    alignas(128) uint8_t stage[kStages][kTileBytes];
    alignas(8) uint64_t full[kStages];
    alignas(8) uint64_t empty[kStages];
    alignas(16) uint32_t st_seg[kStages][kTileSubRows];   // store-segment words of the tile (when bulk-copied)
    uint32_t st_row0[kStages];   // tile held by each stage (written by its producer lane)
    uint32_t st_rows[kStages];   // rows | 0x100 if st_seg is valid; 0 = sentinel: no more tiles on this stage
    BlockPlan bp;
};

template <int kConsumers, int kStages, int kM>
__global__ void __launch_bounds__((kConsumers + 1) * 32, 1) score_topk_scan_tma_kernel(const ScanArgs a) {
    static_assert(kStages % kConsumers == 0, "a stage is always drained by the same consumer warp");
    using Geom = RowGeom<kM>;
    constexpr int kTileRows = Geom::kTileRows;
    constexpr int kRowBytes = Geom::kD;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    using Smem = TmaSmem<kConsumers, kStages>;
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);

    const int qi = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;  // 0 = producer, 1..kConsumers = consumers
    const int k = static_cast<int>(a.k);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], 1);
        }
        mbar_fence_init();
    }
    if (threadIdx.x == 0) stamp(a, 0);
    const uint32_t q_row = a.q_index ? a.q_index[qi] : static_cast<uint32_t>(qi);
    int4 qv[Geom::kQ];
    load_query<kM>(a.q ? a.q + static_cast<size_t>(q_row) * Geom::kD : a.q_inline, lane, qv);
    const int qq_lane = query_sq_lane<kM>(qv, lane);
    stage_plan<kTileRows>(a, qi, sm.bp);  // ends with __syncthreads()
    if (threadIdx.x == 0) stamp(a, 1);
    const BlockPlan &bp = sm.bp;

    WarpTopK top;
    top.reset();

    const uint32_t total_tiles = bp.total_tiles;
    uint32_t *tile_ctr = a.tile_ctr + qi;
    if (total_tiles == 0) {
        // Nothing of this query's scope lives here (store-sharded batches: most queries on most ranks).  Block 0
        // alone writes -- and, in a sharded launch, publishes -- the empty answer; the others are done.
        if (blockIdx.x != 0) return;
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        using Scratch0 = MergeScratch<kConsumers + 1>;
        finish_query(a, qi, top, qq_lane, *reinterpret_cast<Scratch0 *>(&sm.stage[0][0]), k, warp, lane, 1u);
        return;
    }

    if (warp == 0) {
        // ===== producer warp: lane s owns ring stage s and refills it as soon as it is released,
        // so up to kStages bulk copies are in flight per SM.  Tiles are claimed kChunkTiles at a
        // time from a per-query counter (work stealing): SMs that stream faster take more, every
        // block finishes within a few tiles of the others, and a block that starts late (the next
        // query overlapping this one's tail under programmatic dependent launch) just takes less.
        static_assert(kStages <= 32, "one producer lane per stage");
        if (lane < kStages) {
            // The first round is static (tile = block * kStages + lane), so no atomic stands before
            // the first copy; the counter hands out tiles from gridDim.x * kStages on.  Chunks
            // shrink as the query runs out (guided self-scheduling) so the tail is one tile deep.
            const uint32_t static_tiles = gridDim.x * kStages;
            const uint32_t lanes_total = static_tiles;
            uint32_t round = 0;
            uint32_t next = blockIdx.x * kStages + lane, chunk_end = next + 1;
            uint32_t seen = static_tiles;                       // last counter value this lane saw
            while (true) {
                if (round) mbar_wait(&sm.empty[lane], (round - 1) & 1);
                if (next == chunk_end && (a.dbg_flags & 2u)) {   // diagnostics: static round-robin
                    next += static_tiles - 1;
                    chunk_end = next + 1;
                } else if (next == chunk_end) {
                    const uint32_t left = seen < total_tiles ? total_tiles - seen : 0u;
                    const uint32_t want = max(1u, min(static_cast<uint32_t>(kChunkTiles), left / (2u * lanes_total)));
                    seen = static_tiles + atomicAdd(tile_ctr, want);
                    next = seen;
                    chunk_end = min(next + want, total_tiles);
                }
                if (next >= total_tiles) {              // sentinel: this stage is done
                    sm.st_rows[lane] = 0;
                    mbar_arrive(&sm.full[lane]);
                    break;
                }
                uint32_t row0, row_end;
                locate_tile<kTileRows>(bp, next, row0, row_end);
                ++next;
                const uint32_t rows = min(static_cast<uint32_t>(kTileRows), row_end - row0);
                // the tile's store-segment words ride along when their bytes are 16-byte aligned
                constexpr uint32_t kSegBytes = kTileRows * 4u;                      // 128 / 64 / 32: multiples of 16
                const bool seg_copy = rows == kTileRows && (row0 & 3u) == 0u && !(a.dbg_flags & 1u);
                sm.st_row0[lane] = row0;
                sm.st_rows[lane] = rows | (seg_copy ? 0x100u : 0u);
                mbar_arrive_expect_tx(&sm.full[lane], rows * kRowBytes + (seg_copy ? kSegBytes : 0u));   // release: publishes st_*
                bulk_g2s(sm.stage[lane], a.F + static_cast<size_t>(row0) * kRowBytes, rows * kRowBytes, &sm.full[lane]);
                if (seg_copy) bulk_g2s(sm.st_seg[lane], a.seg + row0, kSegBytes, &sm.full[lane]);
                if (round == 0 && lane == 0) stamp(a, 7);
                ++round;
            }
        }
        __syncwarp();
    } else {
        // ===== consumers: warp c drains stages c, c + kConsumers, ... in ring order =====
        const int cw = warp - 1;
        const int my_row_in_tile = row_in_tile<kM>(lane);
        const bool offers = key_lane<kM>(lane);        // wider rows: one lane per row carries the key
        uint64_t *g_floor = a.floors + qi;
        bool first = true;
        constexpr int kMine = kStages / kConsumers;    // stages per consumer warp
        uint32_t done_mask = 0;
        for (uint32_t it = 0; done_mask != (1u << kMine) - 1u; ++it) {
            const uint32_t slot = it % kMine;
            if (done_mask & (1u << slot)) continue;
            const uint32_t s = cw + slot * kConsumers;
            const uint64_t floor_seen = __ldcg(g_floor);
            mbar_wait(&sm.full[s], (it / kMine) & 1);
            const uint32_t rows_word = sm.st_rows[s];
            if (rows_word == 0) {                      // warp-uniform
                done_mask |= 1u << slot;
                continue;
            }
            const uint32_t rows = rows_word & 0xFFu;
            const uint32_t row0 = sm.st_row0[s];
            const uint32_t my_row = row0 + my_row_in_tile;
            uint32_t seg = kTombstone;
            if (offers) {
                if (rows_word & 0x100u) seg = sm.st_seg[s][my_row_in_tile];
                else seg = my_row_in_tile < static_cast<int>(rows) ? __ldg(a.seg + my_row) : kTombstone;
            }
            const int4 *src = reinterpret_cast<const int4 *>(sm.stage[s]) + lane;
            int4 x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = src[j * 32];
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[s]);   // stage may be refilled while we reduce
            const int score = score_tile<kM>(x, qv, lane);
            const uint64_t key = in_scope(bp, seg) ? pack_key(score, a.id_base + my_row) : 0ull;
            if (first && threadIdx.x == 32) stamp(a, 2);
            offer_tile(top, first, key, floor_seen, g_floor, k, lane);
        }
        if (lane == 0) stamp_max(a, 3);
    }
    // Programmatic dependent launch: wait until the previous kernel in the stream has completed
    // (it may still be merging while we scanned; it shares the partial/ticket buffers), then let
    // the next kernel start filling SMs as our blocks retire.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // warp 0 (the producer) becomes the merging warp: its list is empty so far
    // every tile has been consumed: the ring is idle and becomes the merge scratch
    using Scratch = MergeScratch<kConsumers + 1>;
    static_assert(sizeof(Scratch) <= sizeof(sm.stage), "ring too small to double as merge scratch");
    __syncthreads();
    finish_query(a, qi, top, qq_lane, *reinterpret_cast<Scratch *>(&sm.stage[0][0]), k, warp, lane, gridDim.x);
}

// k-way merge of n_lists sorted top-k lists per query (after the all-gather of the sharded path,
// and after the batched GEMM path's per-slice lists): one block per query, one warp per group of
// 32 lists (level-1 tournaments run in parallel), warp 0 plays the final over the group winners.
constexpr int kMergeWarps = 8;
__global__ void __launch_bounds__(kMergeWarps * 32) merge_lists_kernel(const uint64_t *__restrict__ keys, uint32_t n_lists, uint32_t nq,
                                                                       uint32_t k_in, uint32_t k_out, uint64_t *__restrict__ out,
                                                                       uint64_t *__restrict__ floors, uint32_t k_floor) {
    extern __shared__ __align__(16) uint8_t merge_smem[];   // kMergeWarps areas of 32*k_in keys + level 2 of 32*k_in keys
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t qi = blockIdx.x;
    const int kin = static_cast<int>(k_in);
    const int n_groups = static_cast<int>((n_lists + 31) / 32);
    uint64_t *area = reinterpret_cast<uint64_t *>(merge_smem) + static_cast<size_t>(warp) * 32 * kin;
    uint64_t *level2 = reinterpret_cast<uint64_t *>(merge_smem) + static_cast<size_t>(kMergeWarps) * 32 * kin;
    for (int g = warp; g < n_groups; g += kMergeWarps) {
        const int lists_here = min(32, static_cast<int>(n_lists) - g * 32);
        for (int i = lane; i < lists_here * kin; i += 32) {
            const int l = g * 32 + i / kin, j = i % kin;
            area[i] = keys[(static_cast<size_t>(l) * nq + qi) * kin + j];
        }
        __syncwarp();
        const uint64_t w = warp_tournament(area, lists_here, kin, kin, lane);
        if (lane < kin) level2[g * kin + lane] = w;
        __syncwarp();
    }
    __syncthreads();
    if (warp != 0) return;
    const uint64_t w = n_groups == 1 ? level2[lane < kin ? lane : 0] : warp_tournament(level2, n_groups, kin, kin, lane);
    if (lane < static_cast<int>(k_out)) out[static_cast<size_t>(qi) * k_out + lane] = (lane < kin) ? w : 0ull;
    // optional: the k_floor-th best score as a lower-bound key for a following pass (score word only:
    // the low word of a group-maximum key is a group tag, not a chunk id)
    if (floors && lane == static_cast<int>(k_floor) - 1) floors[qi] = lane < kin ? (w & 0xFFFFFFFF00000000ull) : 0ull;
}

// Second half of the publish-only exchange (store-sharded batches): one warp per query.  Lane r acquires rank
// r's flag for the query in this rank's buffer, then the warp plays the tournament over the `world` lists.
__global__ void __launch_bounds__(128) merge_wait_kernel(const uint64_t *__restrict__ gather, const uint32_t *__restrict__ flags,
                                                         const uint8_t *__restrict__ masks, uint32_t world, uint32_t nq_cap, uint32_t nq,
                                                         uint32_t k, uint32_t seq, uint64_t *__restrict__ out,
                                                         uint32_t *__restrict__ timeout) {
    __shared__ uint64_t lists[4][8 * kListCap];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t qi = blockIdx.x * 4 + w;
    if (qi >= nq) return;
    const uint32_t mask = masks ? masks[qi] : 0xFFu;      // ranks that hold rows of this query's scope
    bool arrived = true;
    if (lane < static_cast<int>(world) && ((mask >> lane) & 1u)) {
        const uint32_t *flag = flags + static_cast<size_t>(lane) * nq_cap + qi;
        const long long t0 = clock64();
        uint32_t v;
        while (true) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v == seq) break;
            if (clock64() - t0 > (4ll << 30)) { arrived = false; break; }   // ~2 s: fail loudly, never hang
        }
    }
    arrived = __all_sync(kFull, arrived);
    const int kk = static_cast<int>(k);
    for (int i = lane; i < static_cast<int>(world) * kk; i += 32) {
        const int r = i / kk, j = i % kk;
        lists[w][i] = ((mask >> r) & 1u) ? __ldcg(gather + (static_cast<size_t>(r) * nq_cap + qi) * kk + j) : 0ull;
    }
    __syncwarp();
    const uint64_t merged = warp_tournament(lists[w], static_cast<int>(world), kk, kk, lane);
    if (lane < kk) out[static_cast<size_t>(qi) * kk + lane] = arrived ? merged : 0ull;
    if (!arrived && lane == 0) atomicExch(timeout, 1u);
}

// Packed keys -> the result arrays a host caller gets (same arithmetic as finish_query: RF-1 step 8
// for the cosine).  One warp per query.
__global__ void __launch_bounds__(128) unpack_keys_kernel(const uint64_t *__restrict__ keys, const int8_t *__restrict__ q, uint32_t dim,
                                                          const int32_t *__restrict__ ff, uint32_t id_base, uint32_t nq, uint32_t k,
                                                          uint64_t *__restrict__ out_ids, int32_t *__restrict__ out_scores,
                                                          float *__restrict__ out_cos, uint32_t *__restrict__ out_counts) {
    const int lane = threadIdx.x & 31;
    const uint32_t qi = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (qi >= nq) return;
    int qq = 0;
    for (uint32_t part = 0; part < dim; part += kSubDim) {
        const int2 qv = *reinterpret_cast<const int2 *>(q + static_cast<size_t>(qi) * dim + part + lane * 8);
        qq = __dp4a(qv.x, qv.x, qq);
        qq = __dp4a(qv.y, qv.y, qq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(kFull, qq, o);
    const uint64_t key = lane < static_cast<int>(k) ? keys[static_cast<size_t>(qi) * k + lane] : 0ull;
    const unsigned found = __ballot_sync(kFull, key != 0ull);
    if (lane < static_cast<int>(k)) {
        const size_t o = static_cast<size_t>(qi) * k + lane;
        const uint32_t gid = key_gid(key);
        const int32_t sc = key_score(key);
        out_ids[o] = key ? static_cast<uint64_t>(gid) : ~0ull;
        out_scores[o] = key ? sc : 0;
        if (out_cos) {
            float c = 0.0f;
            if (key && ff) {
                const float nq_ = sqrtf(static_cast<float>(qq));
                const float nf = sqrtf(static_cast<float>(__ldg(ff + (gid - id_base))));
                const float den = __fmul_rn(nq_, nf);
                c = den == 0.0f ? 0.0f : __fdiv_rn(static_cast<float>(sc), den);
            }
            out_cos[o] = c;
        }
    }
    if (lane == 0 && out_counts) out_counts[qi] = __popc(found);
}

template <int kConsumers, int kStages, int kM = 1>
cudaError_t launch_tma(const ScanArgs &a, dim3 grid, cudaStream_t s, bool overlap) {
    using Smem = TmaSmem<kConsumers, kStages>;
    auto kern = score_topk_scan_tma_kernel<kConsumers, kStages, kM>;
    if (cudaError_t e = ensure_dynamic_smem(kern, static_cast<int>(sizeof(Smem))); e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3((kConsumers + 1) * 32, 1, 1);
    cfg.dynamicSmemBytes = sizeof(Smem);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: see griddepcontrol in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap ? 1 : 0;   // without the attribute griddepcontrol.wait returns at once: a fully serialised launch
    return cudaLaunchKernelEx(&cfg, kern, a);
}

}  // namespace

uint32_t scan_default_blocks_per_query(int sm_count, int variant) {
    // blocks resident per SM: the 96 KB-ring variants (and ldg) fit two, the big rings one
    const bool two = variant == kScanVariantLdg || variant == kScanVariantTma6x12 || variant == kScanVariantTma12x12 ||
                     variant == kScanVariantTma4x12;
    if (variant == kScanVariantTma4x8) return static_cast<uint32_t>(sm_count) * 3u;   // 64 KB rings: three per SM
    return static_cast<uint32_t>(sm_count) * (two ? 2u : 1u);
}

cudaError_t launch_score_topk_scan(const ScanArgs &a, uint32_t nq, uint32_t blocks_per_query, int variant, uint32_t dim, cudaStream_t s,
                                   bool overlap) {
    dim3 grid(blocks_per_query, nq, 1);
    // wider rows: the default ring (6 consumers x 12 stages, two blocks per SM) only
    if (dim == 512) return variant == kScanVariantTma6x12 ? launch_tma<6, 12, 2>(a, grid, s, overlap) : cudaErrorInvalidValue;
    if (dim == 1024) return variant == kScanVariantTma6x12 ? launch_tma<6, 12, 4>(a, grid, s, overlap) : cudaErrorInvalidValue;
    if (dim != 256) return cudaErrorInvalidValue;
    switch (variant) {
        case kScanVariantLdg: {
            if (cudaError_t e = ensure_dynamic_smem(score_topk_scan_ldg_kernel, static_cast<int>(sizeof(MergeScratch<kLdgWarps>)));
                e != cudaSuccess)
                return e;
            score_topk_scan_ldg_kernel<<<grid, kLdgWarps * 32, sizeof(MergeScratch<kLdgWarps>), s>>>(a);
            return cudaGetLastError();
        }
        case kScanVariantTma8x24: return launch_tma<8, 24>(a, grid, s, overlap);
        case kScanVariantTma12x24: return launch_tma<12, 24>(a, grid, s, overlap);
        case kScanVariantTma8x16: return launch_tma<8, 16>(a, grid, s, overlap);
        case kScanVariantTma6x12: return launch_tma<6, 12>(a, grid, s, overlap);
        case kScanVariantTma12x12: return launch_tma<12, 12>(a, grid, s, overlap);
        case kScanVariantTma4x12: return launch_tma<4, 12>(a, grid, s, overlap);
        case kScanVariantTma4x8: return launch_tma<4, 8>(a, grid, s, overlap);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_merge_lists(const uint64_t *keys, uint32_t n_lists, uint32_t nq, uint32_t k_in, uint32_t k_out, uint64_t *out,
                               cudaStream_t s, uint64_t *floors, uint32_t k_floor) {
    if (n_lists == 0 || n_lists > 1024 || k_in == 0 || k_in > RF_TOPK_MAX || k_out > k_in) return cudaErrorInvalidValue;
    const size_t smem = (static_cast<size_t>(kMergeWarps) + 1) * 32 * k_in * 8;   // <= 72 KB at k_in = 32
    if (cudaError_t e = ensure_dynamic_smem(merge_lists_kernel, static_cast<int>((kMergeWarps + 1) * 32 * RF_TOPK_MAX * 8)); e != cudaSuccess)
        return e;
    if (floors && (k_floor == 0 || k_floor > k_out)) return cudaErrorInvalidValue;
    merge_lists_kernel<<<nq, kMergeWarps * 32, smem, s>>>(keys, n_lists, nq, k_in, k_out, out, floors, k_floor);
    return cudaGetLastError();
}

cudaError_t launch_unpack_keys(const uint64_t *keys, const int8_t *q, uint32_t dim, const int32_t *ff, uint32_t id_base, uint32_t nq, uint32_t k,
                               uint64_t *out_ids, int32_t *out_scores, float *out_cos, uint32_t *out_counts, cudaStream_t s) {
    unpack_keys_kernel<<<(nq + 3) / 4, 128, 0, s>>>(keys, q, dim, ff, id_base, nq, k, out_ids, out_scores, out_cos, out_counts);
    return cudaGetLastError();
}

cudaError_t launch_merge_wait(const uint64_t *gather, const uint32_t *flags, const uint8_t *masks, uint32_t world, uint32_t nq_cap,
                              uint32_t nq, uint32_t k, uint32_t seq, uint64_t *out, uint32_t *timeout, cudaStream_t s) {
    if (world == 0 || world > 8 || k == 0 || k > RF_TOPK_MAX || nq == 0) return cudaErrorInvalidValue;
    merge_wait_kernel<<<(nq + 3) / 4, 128, 0, s>>>(gather, flags, masks, world, nq_cap, nq, k, seq, out, timeout);
    return cudaGetLastError();
}

cudaError_t launch_merge_topk(const uint64_t *keys, uint32_t n_lists, uint32_t nq, uint32_t k, uint64_t *out_keys,
                              cudaStream_t s) {
    return launch_merge_lists(keys, n_lists, nq, k, k, out_keys, s);
}

}  // namespace rf
