// EXPERIMENT, NOT BUILT: TS-mode variant of score_topk_gemm (queries resident in TMEM, four 64-column
// accumulators, 6-stage feature ring).  Parity-green, and the MMA side drops from ~4100 to ~2950 cycles
// per tile (46 cycles per N=64 MMA), but with only 256 TMEM columns left for accumulators the MMA <->
// epilogue hand-off is too shallow: 4677 cycles per tile against 4107 for the SS-mode kernel in csrc/.
// Kept for the next round (the epilogue's ~665 cycles per 32-column group is the thing to fix first).
// score_topk_gemm: batched query-vs-chunk scoring as a real GEMM on the 5th-generation tensor
// cores (tcgen05.mma kind::i8, S8 x S8 -> S32, accumulators in TMEM), with the per-query top-k
// selection fused into the TMEM epilogue -- the 1024 x 1M score matrix (4 GB) never exists.
// BASELINE.json configs[2]; RF-1 steps 6-7 (oracle/SPEC.md); the retrieval step of
// GeminiRag.ask_stream (reference backend/app/services/gemini_rag.py:517-551) for many queries.
//
// Orientation.  A = queries (M = 128 rows per MMA), B = chunk rows (N = 128 per MMA), K = 256
// int8 = 8 MMAs of K = 32.  D[query, chunk] lands in TMEM with lane = query, column = chunk, so an
// epilogue thread owns ONE query per M-tile and walks its row of scores against that query's own
// threshold held in a register: no cross-lane reduction.  A block keeps kGemmMT M-tiles (512
// queries) of Q resident in shared memory (SW128 K-major, loaded once by TMA) and streams its
// slice of the feature arena through a 3-stage TMA ring; it owns all 512 TMEM columns as four
// 128-column accumulators (one per M-tile), so the MMA warp runs up to four tiles ahead of the
// epilogue.  Grid = (chunk slices, query groups of 512).
//
// Warp roles (18 warps): 0 = TMA producer + TMEM allocator, 1 = MMA issuer (one elected thread),
// 2..17 = epilogue, four warps per accumulator: warp w serves M-tile (w-2)/4 and reads TMEM lane
// quarter w % 4 (hardware rule), so the four accumulators drain concurrently.
// Per 32-column load an epilogue thread takes the max of its 32 scores; only when it reaches the
// query's threshold score does the warp enter the candidate path, which walks the union of the
// lanes' candidate columns (a warp-uniform register pick each; the tenant mask of the tile's 128
// columns was fetched one tile ahead as four ballots) and inserts into the thread's sorted top-10,
// a compare-exchange chain held entirely in registers.  Each thread ends with one list for its
// query; the lists of all slices are merged by merge_lists_kernel (warp tournaments).
//
// Two passes (engine.cu:search_gemm).  Pass A (group_max_mode) runs the same GEMM over a sample of
// the rows but keeps, per query, the top-k of per-32-chunk GROUP MAXIMA: its k-th value is a valid
// lower bound of the query's final k-th best score (k distinct chunks reach it) and costs one
// insertion per group.  Pass B scores every row with those floors, so the candidate path is rare.
// What bounds pass B (profiles/gemm_timeline_r01.txt): the MMA's operand fetch -- both operands
// come from shared memory (SS mode, 8 KB per 128x128x32 MMA), which paces each MMA at ~128 cycles
// against 64 cycles of tensor-pipe time; TMEM read-back sustains 468 B/clk/SM and is not the limit.
#include <algorithm>

#include <cuda.h>

#include "rf_device.cuh"
#include "rf_internal.h"

namespace rf {

namespace {

constexpr int kGemmK = kGemmListK;          // list length kept per (thread, query)
constexpr int kMT = kGemmMT;                // M-tiles (of 128 queries) resident per block
constexpr int kBN = kGemmTileRows;          // chunk rows per B tile / MMA N
constexpr int kStagesB = 6;                // 192 KB feature ring (the queries live in TMEM, not in shared memory)
constexpr int kAccs = 4;                   // 64-column accumulators at tmem columns 0, 64, 128, 192
constexpr int kAccN = 64;                  // MMA N: half a chunk tile per accumulator
constexpr int kQCol0 = 256;                // queries: 4 M-tiles x 64 columns at tmem columns 256..511
constexpr int kEpiWarpsPerAcc = 4;         // one warp per TMEM lane quarter
constexpr int kKBlockBytes = 128;           // one SW128 swizzle row: 128 int8 of K
constexpr int kTileKBlock = 128 * kKBlockBytes;   // 16 KB: 128 rows x 128 B
constexpr int kEpiWarps = kAccs * kEpiWarpsPerAcc;  // 16
constexpr int kGemmThreads = (2 + kEpiWarps) * 32;  // 576

struct GemmSmem {
    alignas(1024) uint8_t b[kStagesB][2][kTileKBlock];      // 192 KB: feature ring (SW128 K-major tiles)
    alignas(8) uint64_t full[kStagesB], empty[kStagesB];
    uint64_t tmem_full[kAccs], tmem_empty[kAccs];
    uint32_t tmem_base;
};

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor: 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(const void *smem) {
    const uint64_t addr = (smem_u32(smem) & 0x3FFFFu) >> 4;
    return addr | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <bool kAccumulate>
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "n"(kAccumulate ? 1 : 0), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
// TS mode: A from tensor memory (lane = row, 4 int8 per 32-bit column), B from shared memory.
template <bool kAccumulate>
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "n"(kAccumulate ? 1 : 0), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// v[j] for a warp-uniform j: a 32-way uniform switch over registers (a single-column TMEM reload
// would queue behind the other warps' 4 KB accumulator loads).
__device__ __forceinline__ uint32_t pick32(const uint32_t (&v)[32], int j) {
    switch (j) {
#define RF_PICK(i) case i: return v[i];
        RF_PICK(0) RF_PICK(1) RF_PICK(2) RF_PICK(3) RF_PICK(4) RF_PICK(5) RF_PICK(6) RF_PICK(7)
        RF_PICK(8) RF_PICK(9) RF_PICK(10) RF_PICK(11) RF_PICK(12) RF_PICK(13) RF_PICK(14) RF_PICK(15)
        RF_PICK(16) RF_PICK(17) RF_PICK(18) RF_PICK(19) RF_PICK(20) RF_PICK(21) RF_PICK(22) RF_PICK(23)
        RF_PICK(24) RF_PICK(25) RF_PICK(26) RF_PICK(27) RF_PICK(28) RF_PICK(29) RF_PICK(30)
#undef RF_PICK
        default: return v[31];
    }
}

// Sorted (descending) top-kGemmK list in registers.
struct RegList {
    uint64_t e[kGemmK];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < kGemmK; ++i) e[i] = 0ull;
    }
    // x > e[kGemmK-1] is the caller's business; a compare-exchange chain bubbles x into place
    __device__ __forceinline__ void insert(uint64_t x) {
        e[kGemmK - 1] = x;
#pragma unroll
        for (int i = kGemmK - 1; i > 0; --i) {
            const uint64_t hi = e[i] > e[i - 1] ? e[i] : e[i - 1];
            const uint64_t lo = e[i] > e[i - 1] ? e[i - 1] : e[i];
            e[i - 1] = hi;
            e[i] = lo;
        }
    }
};

__global__ void __launch_bounds__(kGemmThreads, 1) score_topk_gemm_kernel(const __grid_constant__ CUtensorMap map_f, const GemmArgs a) {
    extern __shared__ __align__(1024) uint8_t gemm_smem_raw[];
    GemmSmem &sm = *reinterpret_cast<GemmSmem *>((reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t slice = blockIdx.x, n_slices = gridDim.x, qgroup = blockIdx.y;

    // this block's slice of the row range, in whole tiles
    const uint32_t total_tiles = (a.row_hi - a.row_lo + kBN - 1) / kBN;
    const uint32_t t_lo = static_cast<uint32_t>(static_cast<uint64_t>(total_tiles) * slice / n_slices);
    const uint32_t t_hi = static_cast<uint32_t>(static_cast<uint64_t>(total_tiles) * (slice + 1) / n_slices);
    const uint32_t n_tiles = t_hi - t_lo;
    const uint32_t q_base = qgroup * (kMT * 128);
    const uint32_t q_here = min(static_cast<uint32_t>(kMT * 128), a.nq - q_base);
    // M-tiles this block runs: 1, 2 or 4 (3 is padded to 4; with two lists per slice at least 2), so
    // that a unit's accumulator -- and with it the epilogue group that owns the query -- is fixed.
    uint32_t m_eff = (q_here + 127) / 128;
    if (m_eff == 3) m_eff = 4;
    if (a.lists_per_slice == 2 && m_eff == 1) m_eff = 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesB; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        for (int i = 0; i < kAccs; ++i) { mbar_init(&sm.tmem_full[i], 1); mbar_init(&sm.tmem_empty[i], kEpiWarpsPerAcc); }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    // ---- queries -> TMEM (the A operand lives there: TS-mode MMA).  Lane = query row of the M-tile,
    // 4 int8 per 32-bit column, 64 columns per M-tile at tmem column kQCol0 + 64 m.  Rows past nq
    // are zero.  Sixteen warps: warp w writes M-tile (w-2)/4, lane quarter w % 4.
    if (warp >= 2) {
        const uint32_t mt = static_cast<uint32_t>(warp - 2) >> 2, lq = warp & 3;
        const uint32_t q = q_base + mt * 128 + lq * 32 + lane;
        const uint4 *src = reinterpret_cast<const uint4 *>(a.q + static_cast<size_t>(q) * kDim);
        const uint32_t ta = tmem + ((lq * 32u) << 16) + kQCol0 + mt * 64;
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
            uint4 w0 = make_uint4(0, 0, 0, 0), w1 = make_uint4(0, 0, 0, 0);
            if (q < a.nq) { w0 = __ldg(src + c); w1 = __ldg(src + c + 1); }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(ta + c * 4), "r"(w0.x), "r"(w0.y),
                         "r"(w0.z), "r"(w0.w), "r"(w1.x), "r"(w1.y), "r"(w1.z), "r"(w1.w)
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (uint32_t t = 0; t < n_tiles; ++t) {
                const uint32_t s = t % kStagesB;
                if (t >= kStagesB) mbar_wait(&sm.empty[s], ((t / kStagesB) - 1) & 1);
                const uint32_t row0 = a.row_lo + (t_lo + t) * kBN;
                mbar_arrive_expect_tx(&sm.full[s], 2 * kTileKBlock);
                for (int kb = 0; kb < 2; ++kb) tma_load_2d(sm.b[s][kb], &map_f, kb * kKBlockBytes, static_cast<int>(row0), &sm.full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the loop (uniform control flow keeps descriptors in
        // uniform registers); one elected lane issues the tcgen05 instructions.  Unit (t, m) uses
        // accumulator acc(t, m); units are issued in (t, m) order and each accumulator's epilogue
        // group drains them in the same order =====
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kAccN >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
        constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);        // SBO = 1024 B, version 1, SWIZZLE_128B
        const uint32_t b_lo0 = ((smem_u32(&sm.b[0][0][0]) & 0x3FFFFu) >> 4) | (1u << 16);
        constexpr uint32_t kKBlockStep = kTileKBlock >> 4;                   // descriptor units (16 B) per K-block
        constexpr uint32_t kHalfStep = (kAccN * kKBlockBytes) >> 4;          // rows 64..127 of a K-block
        uint32_t uses0 = 0, uses1 = 0, uses2 = 0, uses3 = 0;
        long long w_full = 0, w_empty = 0;
        const long long c_start = clock64();
        for (uint32_t t = 0; t < n_tiles; ++t) {
            const uint32_t s = t % kStagesB;
            long long c0 = clock64();
            mbar_wait(&sm.full[s], (t / kStagesB) & 1);
            w_full += clock64() - c0;
            tc_fence_after();
            const uint32_t b_lo = b_lo0 + s * 2 * kKBlockStep;
            for (uint32_t m = 0; m < m_eff; ++m) {
#pragma unroll
                for (uint32_t half = 0; half < 2; ++half) {
                    const uint32_t acc = (((m_eff == 1 ? t : m) & 1u) << 1) | half;
                    const uint32_t n_used = acc == 0 ? uses0 : acc == 1 ? uses1 : acc == 2 ? uses2 : uses3;
                    c0 = clock64();
                    if (n_used) mbar_wait(&sm.tmem_empty[acc], (n_used - 1) & 1);   // epilogue drained this accumulator
                    w_empty += clock64() - c0;
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_tmem = tmem + kQCol0 + m * 64;
#pragma unroll
                        for (int kb = 0; kb < (a.group_max_mode & 4u ? 0 : 2); ++kb) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t db = (static_cast<uint64_t>(kDescHi) << 32) | (b_lo + kb * kKBlockStep + half * kHalfStep + 2u * k);
                                if (kb | k) umma_i8_ts<true>(tmem + acc * kAccN, a_tmem + (kb * 4 + k) * 8, db, idesc);
                                else umma_i8_ts<false>(tmem + acc * kAccN, a_tmem + (kb * 4 + k) * 8, db, idesc);
                            }
                        }
                        umma_commit(&sm.tmem_full[acc]);
                    }
                    __syncwarp();
                    if (acc == 0) ++uses0; else if (acc == 1) ++uses1; else if (acc == 2) ++uses2; else ++uses3;
                }
            }
            if (elect_one()) umma_commit(&sm.empty[s]);   // the stage is free once every MMA that reads it has retired
            __syncwarp();
        }
        if (a.debug && lane == 0 && n_tiles) {
            unsigned long long *d = a.debug + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 8;
            d[0] = clock64() - c_start; d[1] = w_full; d[2] = w_empty; d[3] = n_tiles;
        }
    } else {
        // ===== epilogue: four warps (one per TMEM lane quarter, hardware rule: quarter = w % 4) per
        // accumulator.  Group g = (w-2)/4 drains accumulator g = 2*gi + ch: column half ch of the
        // chunk tile, parity class gi.  A thread owns query row lq*32+lane of the M-tiles its
        // accumulator serves: m_eff == 4 -> M-tiles gi and gi+2 (two lists), m_eff == 2 -> M-tile
        // gi, m_eff == 1 -> M-tile 0 on the chunk tiles with t % 2 == gi =====
        const uint32_t g = static_cast<uint32_t>(warp - 2) >> 2;
        const uint32_t gi = g >> 1;
        const uint32_t ch = g & 1u;
        const uint32_t lq = warp & 3;
        const uint32_t row_in_tile = lq * 32 + lane;
        const uint32_t n_scope = a.n_scope;
        const uint32_t n_mine = m_eff == 4 ? 2u : 1u;                   // M-tiles (lists) of this thread
        RegList list[2];
        uint64_t thr[2];
        bool live[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            list[i].clear();
            const uint32_t m = m_eff == 1 ? 0u : gi + 2u * i;
            const uint32_t q = q_base + m * 128 + row_in_tile;
            live[i] = static_cast<uint32_t>(i) < n_mine && m < m_eff && q < a.nq;   // padding rows never produce candidates
            thr[i] = (a.floors && live[i]) ? a.floors[q] : 0ull;
        }
        long long w_tfull = 0, w_cand = 0, n_cand = 0;
        const long long e_start = clock64();
        uint32_t n_done = 0;                                            // units this group has drained
        const uint32_t t_first = m_eff == 1 ? gi : 0u, t_step = m_eff == 1 ? 2u : 1u;
        uint32_t seg_next[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t row = a.row_lo + (t_lo + t_first) * kBN + ch * 64 + h * 32 + lane;
            seg_next[h] = (t_first < n_tiles && row < a.row_hi) ? __ldg(a.seg + row) : kTombstone;
        }
        for (uint32_t t = t_first; t < n_tiles; t += t_step) {
            const uint32_t row0 = a.row_lo + (t_lo + t) * kBN;
            // Tenant mask of this warp's 64 chunk columns, one bit per column, from words loaded one
            // tile ahead (the candidate path below never touches global memory): bit j of
            // ok_mask[h] <=> row row0 + 64 ch + 32 h + j is in scope, live and inside the range.
            uint32_t ok_mask[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t row = row0 + ch * 64 + h * 32 + lane;
                const uint32_t sg = seg_next[h];
                bool ok = false;
                if (row < a.row_hi && sg != kTombstone)
                    for (uint32_t x = 0; x < n_scope; ++x) ok |= (sg == a.scope[x]);
                ok_mask[h] = __ballot_sync(kFull, ok);
            }
            if (t + t_step < n_tiles) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t row = row0 + t_step * kBN + ch * 64 + h * 32 + lane;
                    seg_next[h] = row < a.row_hi ? __ldg(a.seg + row) : kTombstone;
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (static_cast<uint32_t>(i) >= n_mine) break;
                long long c0 = clock64();
                mbar_wait(&sm.tmem_full[g], n_done & 1);
                w_tfull += clock64() - c0;
                tc_fence_after();
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    const uint32_t col0 = ch * 64 + h * 32;                      // chunk column within the tile
                    const uint32_t taddr = tmem + ((lq * 32u) << 16) + g * kAccN + h * 32;
                    const uint32_t okm = h ? ok_mask[1] : ok_mask[0];
                    uint32_t v[32];
                    tmem_ld32(taddr, v);
                    if (a.group_max_mode & 1u) {
                        // Floor-finding pass: the k-th largest of per-group maxima (a group = these
                        // 32 chunks of this query) is a valid lower bound of the query's k-th best
                        // score -- k distinct chunks reach it -- and costs one insertion per group
                        // instead of one per chunk.  Out-of-scope chunks must not raise the bound.
                        int gm = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) gm = max(gm, ((okm >> j) & 1u) ? static_cast<int>(v[j]) : 0);
                        const uint64_t key = pack_key(gm, a.id_base + row0 + col0);   // low word only makes groups distinct
                        if (live[i] && okm != 0u && key > thr[i]) {
                            list[i].insert(key);
                            const uint64_t kth = list[i].e[kGemmK - 1];
                            if (kth > thr[i]) thr[i] = kth;
                        }
                        continue;
                    }
                    int mx = static_cast<int>(v[0]);
#pragma unroll
                    for (int j = 1; j < 32; ++j) mx = max(mx, static_cast<int>(v[j]));
                    const uint32_t thr_s = static_cast<uint32_t>(thr[i] >> 32);
                    // scores are >= 0 and < 2^31, so the unsigned compare is exact
                    if (__any_sync(kFull, live[i] && static_cast<uint32_t>(mx) >= thr_s)) {
                        const long long cc = clock64();
                        ++n_cand;
                        uint32_t cand = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) cand |= (v[j] >= thr_s ? 1u : 0u) << j;
                        cand &= okm;
                        if (!live[i]) cand = 0;
                        uint32_t uni = __reduce_or_sync(kFull, cand);
                        while (uni) {
                            const int j = __ffs(uni) - 1;
                            uni &= uni - 1;
                            const uint32_t sc = pick32(v, j);                    // j is warp-uniform: a jump, not a reload
                            if ((cand >> j) & 1u) {
                                const uint64_t key = pack_key(static_cast<int32_t>(sc), a.id_base + row0 + col0 + j);
                                if (key > thr[i]) {
                                    list[i].insert(key);
                                    const uint64_t kth = list[i].e[kGemmK - 1];
                                    if (kth > thr[i]) thr[i] = kth;
                                }
                            }
                        }
                        w_cand += clock64() - cc;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.tmem_empty[g]);
                ++n_done;
            }
        }
        if (a.debug && warp == 2 && lane == 0 && n_tiles) {
            unsigned long long *d = a.debug + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 8;
            d[4] = clock64() - e_start; d[5] = w_tfull; d[6] = w_cand; d[7] = n_cand;
        }
        // lists: per (slice, list slot, query); slot = column half (and group when one M-tile is
        // spread over both accumulators)
        const uint32_t L = a.lists_per_slice;
        const uint32_t slot = L == 4 ? gi * 2 + ch : ch;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const uint32_t m = m_eff == 1 ? 0u : gi + 2u * i;
            const uint32_t q = q_base + m * 128 + row_in_tile;
            if (live[i]) {
                uint64_t *dst = a.out_lists + ((static_cast<size_t>(slice) * L + slot) * a.nq + q) * kGemmK;
#pragma unroll
                for (int e = 0; e < kGemmK; ++e) dst[e] = list[i].e[e];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// floors[q] = (k-th best score of query q in `keys` ([nq, k_src] sorted lists)) << 32, for the next pass.
__global__ void floors_from_keys_kernel(const uint64_t *__restrict__ keys, uint32_t nq, uint32_t k_src, uint32_t k, uint64_t *__restrict__ floors) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    // score word only: the low word of a group-maximum key is a group tag, not a chunk id
    if (q < nq) floors[q] = k <= k_src ? (keys[static_cast<size_t>(q) * k_src + (k - 1)] & 0xFFFFFFFF00000000ull) : 0ull;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows, 256] int8 row-major -> boxes of 128 rows x 128 bytes, 128-byte swizzle
bool make_map(CUtensorMap *map, const void *base, uint64_t rows) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {256, rows};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {128, 128};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// lists each slice writes per query: one per column half, and with a single M-tile (nq <= 128) one
// per (accumulator, column half) because both accumulators then serve that M-tile on alternate tiles
uint32_t gemm_lists_per_slice(uint32_t nq) { return nq <= 128 ? 4u : 2u; }
size_t gemm_lists_bytes(uint32_t n_slices, uint32_t nq) { return static_cast<size_t>(n_slices) * gemm_lists_per_slice(nq) * nq * kGemmK * 8; }

cudaError_t launch_score_topk_gemm(const GemmArgs &a, const int8_t *q_dev, const int8_t *F, uint64_t f_rows, uint32_t n_slices,
                                   cudaStream_t s) {
    CUtensorMap map_f;
    if (!make_map(&map_f, F, f_rows)) return cudaErrorNotSupported;
    (void)q_dev;   // the kernel reads the queries through GemmArgs::q
    const int smem = static_cast<int>(sizeof(GemmSmem)) + 1024;
    if (cudaError_t e = ensure_dynamic_smem(score_topk_gemm_kernel, smem); e != cudaSuccess) return e;
    dim3 grid(n_slices, (a.nq + kMT * 128 - 1) / (kMT * 128), 1);
    score_topk_gemm_kernel<<<grid, kGemmThreads, smem, s>>>(map_f, a);
    return cudaGetLastError();
}

cudaError_t launch_floors_from_keys(const uint64_t *keys, uint32_t nq, uint32_t k_src, uint32_t k, uint64_t *floors, cudaStream_t s) {
    floors_from_keys_kernel<<<(nq + 127) / 128, 128, 0, s>>>(keys, nq, k_src, k, floors);
    return cudaGetLastError();
}

}  // namespace rf
