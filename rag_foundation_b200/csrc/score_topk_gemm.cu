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
#include "rf_gemm_device.cuh"
#include "rf_internal.h"

namespace rf {

namespace {

using namespace gemm;

constexpr int kMT = kGemmMT;                // M-tiles (of 128 queries) resident per block
constexpr int kBN = kGemmTileRows;          // chunk rows per B tile / MMA N
constexpr int kStagesB = 3;
constexpr int kEpiWarpsPerTile = 4;          // one warp per TMEM lane quarter
constexpr int kEpiWarps = kMT * kEpiWarpsPerTile;   // 16: one group of four warps per accumulator
constexpr int kGemmThreads = (2 + kEpiWarps) * 32;  // 576

struct GemmSmem {
    alignas(1024) uint8_t q[kMT][2][kTileKBlock];           // 128 KB: resident query tiles
    alignas(1024) uint8_t b[kStagesB][2][kTileKBlock];      //  96 KB: feature ring
    alignas(8) uint64_t q_full;
    uint64_t full[kStagesB], empty[kStagesB];
    uint64_t tmem_full[kMT], tmem_empty[kMT];
    uint32_t tmem_base;
};

// kDebug: in-kernel cycle counters (RF_SCAN_DEBUG=1); the production instantiation has no clock reads in its loops
template <bool kDebug>
__global__ void __launch_bounds__(kGemmThreads, 1)
score_topk_gemm_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_f, const GemmArgs a) {
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
    const uint32_t m_tiles = (q_here + 127) / 128;
    // Fewer than four M-tiles: the spare accumulators take alternate chunk tiles of the same M-tile,
    // so all four epilogue groups (and the MMA look-ahead) stay busy for small batches.
    const uint32_t reps = a.lists_per_slice;   // host-chosen so every block of the launch agrees: > 1 only with one query group

    if (threadIdx.x == 0) {
        mbar_init(&sm.q_full, 1);
        for (int s = 0; s < kStagesB; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        for (int m = 0; m < kMT; ++m) { mbar_init(&sm.tmem_full[m], 1); mbar_init(&sm.tmem_empty[m], kEpiWarpsPerTile); }
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

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0 && n_tiles) {
            mbar_arrive_expect_tx(&sm.q_full, m_tiles * 2 * kTileKBlock);
            for (uint32_t m = 0; m < m_tiles; ++m)
                for (int kb = 0; kb < 2; ++kb)
                    tma_load_2d(sm.q[m][kb], &map_q, kb * kKBlockBytes, static_cast<int>(q_base + m * 128), &sm.q_full);
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
        // uniform registers); one elected lane issues the tcgen05 instructions =====
        if (n_tiles) {
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kBN >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
            constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);        // SBO = 1024 B, version 1, SWIZZLE_128B
            const uint32_t q_lo0 = ((smem_u32(&sm.q[0][0][0]) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t b_lo0 = ((smem_u32(&sm.b[0][0][0]) & 0x3FFFFu) >> 4) | (1u << 16);
            constexpr uint32_t kKBlockStep = kTileKBlock >> 4;                   // descriptor units (16 B) per K-block
            mbar_wait(&sm.q_full, 0);
            long long w_full = 0, w_empty = 0;
            const long long c_start = kDebug ? clock64() : 0;
            for (uint32_t t = 0; t < n_tiles; ++t) {
                const uint32_t s = t % kStagesB;
                long long c0 = kDebug ? clock64() : 0;
                mbar_wait(&sm.full[s], (t / kStagesB) & 1);
                if (kDebug) w_full += clock64() - c0;
                tc_fence_after();
                const uint32_t b_lo = b_lo0 + s * 2 * kKBlockStep;
                for (uint32_t m = 0; m < m_tiles; ++m) {
                    const uint32_t acc = m + m_tiles * (t % reps);      // accumulator (and epilogue group) of this unit
                    if (kDebug) c0 = clock64();
                    if (t >= reps) mbar_wait(&sm.tmem_empty[acc], ((t / reps) - 1) & 1);   // epilogue drained this accumulator
                    if (kDebug) w_empty += clock64() - c0;
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t q_lo = q_lo0 + m * 2 * kKBlockStep;
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t da = (static_cast<uint64_t>(kDescHi) << 32) | (q_lo + kb * kKBlockStep + 2u * k);
                                const uint64_t db = (static_cast<uint64_t>(kDescHi) << 32) | (b_lo + kb * kKBlockStep + 2u * k);
                                if (kb | k) umma_i8<true>(tmem + acc * kBN, da, db, idesc);
                                else umma_i8<false>(tmem + acc * kBN, da, db, idesc);
                            }
                        }
                        umma_commit(&sm.tmem_full[acc]);
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(&sm.empty[s]);   // the stage is free once every MMA that reads it has retired
                __syncwarp();
            }
            if (kDebug && a.debug && lane == 0) {
                unsigned long long *d = a.debug + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 8;
                d[0] = clock64() - c_start; d[1] = w_full; d[2] = w_empty; d[3] = n_tiles;
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: group g = (warp - 2) / 4 owns accumulator g, i.e. M-tile g % m_tiles and
        // the chunk tiles t with t % reps == g / m_tiles =====
        const uint32_t g = static_cast<uint32_t>(warp - 2) >> 2;
        const uint32_t m = reps > 1 ? g % m_tiles : g;
        const uint32_t rep = reps > 1 ? g / m_tiles : 0;
        const uint32_t lq = warp & 3;                      // TMEM lane quarter this warp may touch
        const uint32_t row_in_tile = lq * 32 + lane;       // query row within the M-tile
        const uint32_t q = q_base + m * 128 + row_in_tile;
        const uint32_t n_scope = a.n_scope;
        const uint32_t sc0 = a.scope[0], sc1 = a.scope[1], sc2 = a.scope[2], sc3 = a.scope[3];
        RegList list;
        list.clear();
        uint64_t thr = (a.floors && q < a.nq) ? a.floors[q] : 0ull;
        const bool live = q < a.nq;                        // padding rows never produce candidates
        long long w_tfull = 0, w_cand = 0, n_cand = 0;
        const long long e_start = kDebug ? clock64() : 0;
        if (g < m_tiles * reps) {
            uint32_t seg_next[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const uint32_t row = a.row_lo + (t_lo + rep) * kBN + h * 32 + lane;
                seg_next[h] = (rep < n_tiles && row < a.row_hi) ? __ldg(a.seg + row) : kTombstone;
            }
            for (uint32_t t = rep; t < n_tiles; t += reps) {
                const uint32_t row0 = a.row_lo + (t_lo + t) * kBN;
                // Tenant mask of the tile's 128 chunk columns, one bit per column, fetched with four
                // coalesced loads before the accumulator is awaited (the candidate path below never
                // touches global memory): bit j of ok_mask[h] <=> row row0 + 32 h + j is in scope,
                // not tombstoned and inside the row range.
                uint32_t ok_mask[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const uint32_t row = row0 + h * 32 + lane;
                    const uint32_t sg = seg_next[h];                 // loaded one tile ahead
                    // first four scope words from registers (unused entries hold the tombstone value):
                    // an indexed parameter load per comparison is a dependent constant-cache round trip
                    bool ok = (sg == sc0) | (sg == sc1) | (sg == sc2) | (sg == sc3);
                    if (n_scope > 4)
                        for (uint32_t x = 4; x < n_scope; ++x) ok |= (sg == a.scope[x]);
                    ok_mask[h] = __ballot_sync(kFull, ok && row < a.row_hi && sg != kTombstone);
                }
                if (t + reps < n_tiles) {
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const uint32_t row = row0 + reps * kBN + h * 32 + lane;
                        seg_next[h] = row < a.row_hi ? __ldg(a.seg + row) : kTombstone;
                    }
                }
                long long c0 = kDebug ? clock64() : 0;
                mbar_wait(&sm.tmem_full[g], (t / reps) & 1);
                if (kDebug) w_tfull += clock64() - c0;
                tc_fence_after();
#pragma unroll 1
                for (int h = 0; h < 4; ++h) {
                    const uint32_t col0 = h * 32;                                // chunk column within the tile
                    const uint32_t taddr = tmem + ((lq * 32u) << 16) + g * kBN + col0;
                    const uint32_t okm = h == 0 ? ok_mask[0] : h == 1 ? ok_mask[1] : h == 2 ? ok_mask[2] : ok_mask[3];
                    uint32_t v[32];
                    tmem_ld32(taddr, v);
                    if (a.group_max_mode) {
                        // Floor-finding pass: the k-th largest of per-group maxima (a group = these
                        // 32 chunks of this query) is a valid lower bound of the query's k-th best
                        // score -- k distinct chunks reach it -- and costs one insertion per group
                        // instead of one per chunk.  Out-of-scope chunks must not raise the bound.
                        int gm = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) gm = max(gm, ((okm >> j) & 1u) ? static_cast<int>(v[j]) : 0);
                        const uint64_t key = pack_key(gm, a.id_base + row0 + col0);   // low word only makes groups distinct
                        if (live && okm != 0u && key > thr) {
                            list.insert(key);
                            const uint64_t kth = list.e[kGemmK - 1];
                            if (kth > thr) thr = kth;
                        }
                        continue;
                    }
                    const int mx = max32(v);
                    const uint32_t thr_s = static_cast<uint32_t>(thr >> 32);
                    // scores are >= 0 and < 2^31, so the unsigned compare is exact
                    if (__any_sync(kFull, live && static_cast<uint32_t>(mx) >= thr_s)) {
                        const long long cc = kDebug ? clock64() : 0;
                        if (kDebug) ++n_cand;
                        uint32_t cand = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) cand |= (v[j] >= thr_s ? 1u : 0u) << j;
                        cand &= okm;
                        if (!live) cand = 0;
                        uint32_t uni = __reduce_or_sync(kFull, cand);
                        while (uni) {
                            const int j = __ffs(uni) - 1;
                            uni &= uni - 1;
                            const uint32_t sc = pick32(v, j);                    // j is warp-uniform: a jump, not a reload
                            if ((cand >> j) & 1u) {
                                const uint64_t key = pack_key(static_cast<int32_t>(sc), a.id_base + row0 + col0 + j);
                                if (key > thr) {
                                    list.insert(key);
                                    const uint64_t kth = list.e[kGemmK - 1];
                                    if (kth > thr) thr = kth;
                                }
                            }
                        }
                        if (kDebug) w_cand += clock64() - cc;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.tmem_empty[g]);
            }
            if (kDebug && a.debug && warp == 2 && lane == 0) {
                unsigned long long *d = a.debug + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 8;
                d[4] = clock64() - e_start; d[5] = w_tfull; d[6] = w_cand; d[7] = n_cand;
            }
            // one list per (slice, replica, query)
            if (live) {
                uint64_t *dst = a.out_lists + ((static_cast<size_t>(slice) * reps + rep) * a.nq + q) * kGemmK;
#pragma unroll
                for (int i = 0; i < kGemmK; ++i) dst[i] = list.e[i];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace

uint32_t gemm_lists_per_slice(uint32_t nq) {
    const uint32_t m_tiles_max = std::min<uint32_t>((nq + 127) / 128, kGemmMT);   // every query group of a launch has this many, or fewer in the last
    return nq > static_cast<uint32_t>(kGemmMT) * 128 ? 1u : gemm_replicas(m_tiles_max);
}
size_t gemm_lists_bytes(uint32_t n_slices, uint32_t nq) { return static_cast<size_t>(n_slices) * gemm_lists_per_slice(nq) * nq * kGemmK * 8; }

cudaError_t launch_score_topk_gemm(const GemmArgs &a, const int8_t *q_dev, const int8_t *F, uint64_t f_rows, uint32_t n_slices,
                                   cudaStream_t s) {
    CUtensorMap map_q, map_f;
    if (!make_map(&map_q, q_dev, a.nq) || !make_map(&map_f, F, f_rows)) return cudaErrorNotSupported;
    const int smem = static_cast<int>(sizeof(GemmSmem)) + 1024;
    dim3 grid(n_slices, (a.nq + kMT * 128 - 1) / (kMT * 128), 1);
    if (a.debug) {
        if (cudaError_t e = ensure_dynamic_smem(score_topk_gemm_kernel<true>, smem); e != cudaSuccess) return e;
        score_topk_gemm_kernel<true><<<grid, kGemmThreads, smem, s>>>(map_q, map_f, a);
    } else {
        if (cudaError_t e = ensure_dynamic_smem(score_topk_gemm_kernel<false>, smem); e != cudaSuccess) return e;
        score_topk_gemm_kernel<false><<<grid, kGemmThreads, smem, s>>>(map_q, map_f, a);
    }
    return cudaGetLastError();
}

}  // namespace rf
