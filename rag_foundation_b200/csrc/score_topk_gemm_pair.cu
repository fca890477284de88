// score_topk_gemm_pair: the batched scoring GEMM on CTA PAIRS (tcgen05 cta_group::2), sm_100a.
//
// Same job as score_topk_gemm.cu (S = Q . F^T in int8 -> int32 with the per-query top-k taken in the
// epilogue; BASELINE.json configs[2]; retrieval step of GeminiRag.ask_stream, reference
// backend/app/services/gemini_rag.py:517-551), laid out for the tensor pipe's full rate:
// a single-CTA kind::i8 MMA of 128 x 128 x 32 is paced at ~130-150 cycles whatever its operands'
// source (tools/probe/umma_contention_probe.cu), a CTA-pair MMA of 256 x 256 x 32 runs in exactly
// 128 cycles = 64 cycles per 128 x 128 x 32 per SM, even with 16 warps per SM reading accumulators
// back at the same time (tools/probe/umma_pair_probe.cu; operand split checked by umma_pair_check.cu).
//
// A cluster of two CTAs (two SMs of one TPC) owns 512 queries and a slice of the chunk rows:
//   * queries: M-group g (g = 0, 1) is 256 queries; CTA r keeps rows [256 g + 128 r, +128) of it
//     resident in shared memory (64 KB per CTA);
//   * chunks: a tile is 256 rows; CTA r streams rows [128 r, +128) of it through its own ring
//     (32 KB per stage) -- the half of the MMA's B operand it contributes;
//   * one thread of the leader CTA issues, per tile and M-group, 8 MMAs (K = 8 x 32) of
//     M = 256 x N = 256 into accumulator g: in EACH CTA's tensor memory 128 lanes (its queries) x
//     256 columns (the tile's chunks); the two accumulators (2 x 256 = all 512 columns) alternate,
//     so the MMAs of one M-group overlap the epilogue of the other;
//   * TMA loads of both CTAs complete on the leader's "full" barrier (cp.async.bulk.tensor
//     .cta_group::2 + a remote arrive.expect_tx); tcgen05.commit multicasts "stage free" and
//     "accumulator ready" to both CTAs; the 2 x 16 epilogue warps hand an accumulator back with
//     remote arrives on the leader's barrier.
// Epilogue: all 16 warps of a CTA drain the accumulator that is ready -- warp w reads TMEM lane
// quarter w % 4 (its 32 queries) x 64 of the 256 columns in two tcgen05.ld.32x32b.x32; a thread
// owns one query per M-group and keeps two sorted top-10 lists in registers.  The tenant mask, the
// candidate path, the floor (group-maximum) pass and the list merge are those of score_topk_gemm.cu;
// the four column-block lists of a query are merged in shared memory at the end, so a pair writes one
// list per query and slice.
// Floor pass (kFloorPass): no lists at all.  A thread keeps ONE running maximum per M-group over every column
// it reads (its 64 columns of each tile of the slice) and writes it at the end: n_slices x 4 group maxima per
// query, of disjoint chunk groups, whose k-th largest (kth_largest_kernel) is a valid lower bound of the
// query's final k-th best score.  Branch-free (two LDTM + two maximum trees per accumulator), so this pass
// runs at the tensor pipe's pace and can afford a sample four times the size the list-keeping version could:
// the main pass then meets a quarter of the candidates, and the candidates are what slows it down (a
// candidate costs its warp ~1000 cycles, and the next tile's MMAs wait for the slowest of the pair's 32 warps).
#include <algorithm>
#include <cstddef>

#include <cuda.h>

#include "rf_device.cuh"
#include "rf_gemm_device.cuh"
#include "rf_internal.h"

namespace rf {

namespace {

using namespace gemm;

constexpr int kPN = kGemmPairTileRows;      // chunk rows per pair tile = MMA N
constexpr int kPStages = 4;                 // 32 KB per stage per CTA
constexpr int kPGroups = 2;                 // M-groups (of 256 queries) per pair = accumulators per CTA
constexpr int kPEpiWarps = 16;
constexpr int kPThreads = (kPEpiWarps + 2) * 32;   // 576: 16 epilogue warps, then the TMA producer and the MMA issuer
// Warp roles.  The warp scheduler prefers the HIGHEST warp id among the eligible warps of its partition, and the epilogue
// works in bursts (all 16 warps wake on the same "accumulator ready"): an MMA-issuing warp with a low id waits behind
// four busy epilogue warps for its ~30 issue slots per accumulator while the tensor pipe idles.  So the epilogue takes
// warps 0..15 (lane quarter = warp % 4 as the hardware requires) and the producer and the MMA issuer the two highest ids.
constexpr int kPProducerWarp = kPEpiWarps, kPMmaWarp = kPEpiWarps + 1;
constexpr int kPColBlocks = 4;              // blocks of 64 accumulator columns, one epilogue warp each per lane quarter

struct PairSmem {
    alignas(1024) uint8_t q[kPGroups][2][kTileKBlock];     // 64 KB: this CTA's 128 queries of each M-group
    alignas(1024) uint8_t b[kPStages][2][kTileKBlock];     // 128 KB: this CTA's half of the chunk tiles
    alignas(8) uint64_t q_full;                            // leader's copy is the one in use (2 arrivals + both CTAs' bytes)
    uint64_t full[kPStages];                               // leader's copy in use
    uint64_t empty[kPStages];                              // own copy (multicast commit)
    uint64_t tmem_full[kPGroups];                          // own copy (multicast commit)
    uint64_t tmem_empty[kPGroups];                         // leader's copy in use (2 x 16 warp arrivals)
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void *p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void remote_arrive_expect_tx(uint32_t cluster_bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void remote_arrive(uint32_t cluster_bar) {
    // CTA-scope release (the PTX default), as in the single-CTA kernel: what must be ordered before the
    // arrive is the accumulator read, which tcgen05.wait::ld + tcgen05.fence::before_thread_sync already
    // completed; a cluster-scope release also waits for the warp's in-flight global loads (the tenant
    // words prefetched for the next tile) and was measured at ~1300 cycles per arrive
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, int c0, int c1, uint32_t cluster_bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1)
                 : "memory");
}
template <bool kAccumulate>
__device__ __forceinline__ void umma2_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "n"(kAccumulate ? 1 : 0), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
// arrive on the barrier at this shared offset in BOTH CTAs once every MMA issued so far has retired
__device__ __forceinline__ void umma2_commit_both(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
}

// Barrier operations on 32-bit shared-window addresses computed once per thread: the shared structure
// sits behind a manually aligned pointer, so the compiler converts generic -> shared again at every
// use (two special-register reads and a dozen instructions per mbarrier wait inside the hot loops).
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta_a(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void umma2_commit_both_a(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
}

// One 32-column group of one query's scores, in two steps: the step that needs the values can run while
// the accumulator is still held, the step that touches the list after it has been handed back.
//
// group_candidate: the lane's single candidate key of the group (0 = none) -- in normal mode the one score
// that reaches the query's threshold, in group-maximum mode (floor pass) the group's maximum.  Out-of-scope
// columns are zeroed first (uniform branch: a tile that lies wholly inside the scope skips it), so the
// maximum is itself a candidate's score and "exactly one score >= threshold" means that score is the
// maximum.  `multi` (warp-uniform) reports that some lane has several candidates in this group: the caller
// then runs take_group_multi on the same values (rare: two chunks of one 32-chunk group above a query's floor).
__device__ __forceinline__ uint64_t group_candidate(uint32_t (&v)[32], uint32_t okm, bool live, bool group_max_mode, uint32_t id0,
                                                    uint64_t thr, bool &multi) {
    multi = false;
    if (okm != kFull) {                                  // warp-uniform
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = ((okm >> j) & 1u) ? v[j] : 0u;
    }
    const int mx = max32(v);
    if (group_max_mode) return (live && okm != 0u) ? pack_key(mx, id0) : 0ull;      // low word only makes groups distinct
    const uint32_t thr_s = static_cast<uint32_t>(thr >> 32);
    const bool hit = live && static_cast<uint32_t>(mx) >= thr_s;     // scores are in [0, 2^31): unsigned compare is exact
    uint64_t key = 0ull;
    if (__any_sync(kFull, hit)) {
        uint32_t cand = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) cand |= (v[j] >= thr_s ? 1u : 0u) << j;
        cand = hit ? (cand & okm) : 0u;
        const int cnt = __popc(cand);
        multi = __any_sync(kFull, cnt > 1);
        if (cnt == 1) key = pack_key(mx, id0 + (__ffs(cand) - 1));
    }
    return key;
}
__device__ __forceinline__ void offer_key(uint64_t key, RegList &list, uint64_t &thr) {
    if (key > thr) {
        list.insert(key);
        const uint64_t kth = list.e[kGemmK - 1];
        if (kth > thr) thr = kth;
    }
}
// several candidates per lane: every score that reaches the threshold, one warp-uniform column at a time
__device__ __forceinline__ void take_group_multi(const uint32_t (&v)[32], uint32_t okm, bool live, uint32_t id0, RegList &list, uint64_t &thr) {
    const uint32_t thr_s = static_cast<uint32_t>(thr >> 32);
    uint32_t cand = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) cand |= (v[j] >= thr_s ? 1u : 0u) << j;
    cand &= okm;
    if (!live) cand = 0;
    uint32_t uni = __reduce_or_sync(kFull, cand);
    while (uni) {
        const int j = __ffs(uni) - 1;
        uni &= uni - 1;
        const uint32_t sc = pick32(v, j);        // j is warp-uniform: a jump, not a TMEM reload
        if ((cand >> j) & 1u) offer_key(pack_key(static_cast<int32_t>(sc), id0 + j), list, thr);
    }
}

// kDebug: in-kernel cycle counters (RF_SCAN_DEBUG=1, tools/gemm_timeline.py); the production instantiation
// carries no clock reads in its loops (they cost 5 % of the batch time)
template <bool kDebug, bool kFloorPass>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPThreads, 1)
score_topk_gemm_pair_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_f, const GemmArgs a) {
    extern __shared__ __align__(1024) uint8_t pair_smem_raw[];
    // the dynamic window starts at the same offset in both CTAs, so aligned addresses agree too
    PairSmem &sm = *reinterpret_cast<PairSmem *>((reinterpret_cast<uintptr_t>(pair_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t slice = blockIdx.x >> 1, n_slices = gridDim.x >> 1, qgroup = blockIdx.y;

    const uint32_t total_tiles = (a.row_hi - a.row_lo + kPN - 1) / kPN;
    const uint32_t t_lo = static_cast<uint32_t>(static_cast<uint64_t>(total_tiles) * slice / n_slices);
    const uint32_t t_hi = static_cast<uint32_t>(static_cast<uint64_t>(total_tiles) * (slice + 1) / n_slices);
    const uint32_t n_tiles = t_hi - t_lo;
    const uint32_t q_base = qgroup * (kPGroups * 256);
    const uint32_t q_here = min(static_cast<uint32_t>(kPGroups * 256), a.nq - q_base);
    const uint32_t m_groups = (q_here + 255) / 256;      // 1 only in a ragged last query group

    if (threadIdx.x == 0) {
        mbar_init(&sm.q_full, 2);
        for (int s = 0; s < kPStages; ++s) { mbar_init(&sm.full[s], 2); mbar_init(&sm.empty[s], 1); }
        for (int g = 0; g < kPGroups; ++g) { mbar_init(&sm.tmem_full[g], 1); mbar_init(&sm.tmem_empty[g], 2 * kPEpiWarps); }
        mbar_fence_init();
    }
    __syncthreads();
    cluster_sync_all();                                   // both CTAs' barriers exist before anyone signals across
    if (warp == kPProducerWarp) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    // shared-window addresses of the barriers (one conversion per thread)
    uint32_t sm_a;
    {
        // derived from the shared array itself (a symbol the compiler knows to be in shared memory; the
        // window base is far more aligned than 1024 B), and pinned in a register by an opaque move so it
        // is not re-derived from the generic pointer at each use
        const uint32_t raw_a = smem_u32(pair_smem_raw);
        const uint32_t aligned = (raw_a + 1023u) & ~1023u;
        asm volatile("mov.u32 %0, %1;" : "=r"(sm_a) : "r"(aligned));
    }
    const uint32_t a_q_full = sm_a + static_cast<uint32_t>(offsetof(PairSmem, q_full));
    const uint32_t a_full = sm_a + static_cast<uint32_t>(offsetof(PairSmem, full));
    const uint32_t a_empty = sm_a + static_cast<uint32_t>(offsetof(PairSmem, empty));
    const uint32_t a_tmem_full = sm_a + static_cast<uint32_t>(offsetof(PairSmem, tmem_full));
    const uint32_t a_tmem_empty = sm_a + static_cast<uint32_t>(offsetof(PairSmem, tmem_empty));

    if (warp == kPProducerWarp) {
        // ===== TMA producer (both CTAs): own rows, bytes counted on the leader's barriers =====
        if (lane == 0 && n_tiles) {
            const uint32_t leader_q_full = map_to_cta_a(a_q_full, 0);
            remote_arrive_expect_tx(leader_q_full, m_groups * 2 * kTileKBlock);
            for (uint32_t g = 0; g < m_groups; ++g)
                for (int kb = 0; kb < 2; ++kb)
                    tma_load_2d_pair(sm.q[g][kb], &map_q, kb * kKBlockBytes, static_cast<int>(q_base + g * 256 + rank * 128), leader_q_full);
            for (uint32_t t = 0; t < n_tiles; ++t) {
                const uint32_t s = t % kPStages;
                if (t >= kPStages) mbar_wait_a(a_empty + 8 * s, ((t / kPStages) - 1) & 1);
                const uint32_t row0 = a.row_lo + (t_lo + t) * kPN + rank * 128;
                const uint32_t leader_full = map_to_cta_a(a_full + 8 * s, 0);
                remote_arrive_expect_tx(leader_full, 2 * kTileKBlock);
                for (int kb = 0; kb < 2; ++kb) tma_load_2d_pair(sm.b[s][kb], &map_f, kb * kKBlockBytes, static_cast<int>(row0), leader_full);
            }
        }
        __syncwarp();
    } else if (warp == kPMmaWarp) {
        // ===== MMA issuer: leader CTA only; the whole warp walks the loop, one elected lane issues =====
        if (rank == 0 && n_tiles) {
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kPN >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
            constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);        // SBO = 1024 B, version 1, SWIZZLE_128B
            const uint32_t q_lo0 = ((smem_u32(&sm.q[0][0][0]) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t b_lo0 = ((smem_u32(&sm.b[0][0][0]) & 0x3FFFFu) >> 4) | (1u << 16);
            constexpr uint32_t kKBlockStep = kTileKBlock >> 4;                   // descriptor units (16 B) per K-block
            mbar_wait_a(a_q_full, 0);
            long long w_full = 0, w_empty = 0;
            const long long c_start = kDebug ? clock64() : 0;
            for (uint32_t t = 0; t < n_tiles; ++t) {
                const uint32_t s = t % kPStages;
                long long c0 = kDebug ? clock64() : 0;
                mbar_wait_a(a_full + 8 * s, (t / kPStages) & 1);
                if (kDebug) w_full += clock64() - c0;
                tc_fence_after();
                const uint32_t b_lo = b_lo0 + s * 2 * kKBlockStep;
                for (uint32_t g = 0; g < m_groups; ++g) {
                    if (kDebug) c0 = clock64();
                    if (t >= 1) mbar_wait_a(a_tmem_empty + 8 * g, (t - 1) & 1);      // both CTAs' epilogues drained accumulator g
                    if (kDebug) w_empty += clock64() - c0;
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t q_lo = q_lo0 + g * 2 * kKBlockStep;
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t da = (static_cast<uint64_t>(kDescHi) << 32) | (q_lo + kb * kKBlockStep + 2u * k);
                                const uint64_t db = (static_cast<uint64_t>(kDescHi) << 32) | (b_lo + kb * kKBlockStep + 2u * k);
                                if (kb | k) umma2_i8<true>(tmem + g * kPN, da, db, idesc);
                                else umma2_i8<false>(tmem + g * kPN, da, db, idesc);
                            }
                        }
                        umma2_commit_both_a(a_tmem_full + 8 * g);
                        // timeline (debug build, first cluster): MMAs of (tile, group) issued
                        if (kDebug && a.debug && blockIdx.x == 0 && blockIdx.y == 0 && t >= 16 && t < 24)
                            a.debug[2048 + ((t - 16) * 2 + g) * 8 + 0] = clock64();
                    }
                    __syncwarp();
                }
                if (elect_one()) umma2_commit_both_a(a_empty + 8 * s);   // both CTAs' halves of the stage are free once these MMAs retire
                __syncwarp();
            }
            if (kDebug && a.debug && lane == 0) {
                unsigned long long *d = a.debug + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 8;
                d[0] = clock64() - c_start; d[1] = w_full; d[2] = w_empty; d[3] = n_tiles;
            }
        }
        __syncwarp();
    } else if (n_tiles) {
        // ===== epilogue (both CTAs): warp -> lane quarter lq, column block cb; thread -> one query per M-group =====
        const uint32_t lq = warp & 3;
        const uint32_t cb = static_cast<uint32_t>(warp) >> 2;
        const uint32_t n_scope = a.n_scope;
        const uint32_t q0 = q_base + rank * 128 + lq * 32 + lane;          // + 256 g
        const bool live0 = q0 < a.nq, live1 = q0 + 256 < a.nq;
        RegList list0, list1;
        list0.clear();
        list1.clear();
        uint64_t thr0 = (a.floors && live0) ? a.floors[q0] : 0ull;
        uint64_t thr1 = (a.floors && live1) ? a.floors[q0 + 256] : 0ull;
        const uint32_t leader_empty0 = map_to_cta_a(a_tmem_empty, 0), leader_empty1 = map_to_cta_a(a_tmem_empty + 8, 0);
        constexpr bool gmm = false;                        // (the floor pass has its own branch below)
        int run0 = 0, run1 = 0;                            // floor pass: running maxima of this thread's columns
        long long w_tfull = 0;
        const long long e_start = kDebug ? clock64() : 0;
        // The scope test runs once per tile and column half: keep the first four scope words in
        // registers (read with immediate offsets from the parameter bank; unused entries hold the
        // tombstone value) -- an indexed parameter load per comparison costs a dependent constant-cache
        // round trip, which made this the most expensive part of the tile loop.
        const uint32_t sc0 = a.scope[0], sc1 = a.scope[1], sc2 = a.scope[2], sc3 = a.scope[3];
        const uint32_t row_hi = a.row_hi;
        const uint32_t *__restrict__ seg_words = a.seg;
        auto in_scope = [&](uint32_t sg) {
            bool ok = (sg == sc0) | (sg == sc1) | (sg == sc2) | (sg == sc3);
            if (n_scope > 4)
                for (uint32_t x = 4; x < n_scope; ++x) ok |= (sg == a.scope[x]);
            return ok && sg != kTombstone;
        };
        uint32_t seg_next[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t row = a.row_lo + t_lo * kPN + cb * 64 + h * 32 + lane;
            seg_next[h] = row < row_hi ? __ldg(seg_words + row) : kTombstone;
        }
        for (uint32_t t = 0; t < n_tiles; ++t) {
            const uint32_t row0 = a.row_lo + (t_lo + t) * kPN + cb * 64;   // chunk row of this warp's first column
            // tenant mask of this warp's 64 columns as two ballots (words loaded one tile ahead)
            uint32_t ok_mask[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t row = row0 + h * 32 + lane;
                ok_mask[h] = __ballot_sync(kFull, row < row_hi && in_scope(seg_next[h]));
            }
            if (t + 1 < n_tiles) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t row = row0 + kPN + h * 32 + lane;
                    seg_next[h] = row < row_hi ? __ldg(seg_words + row) : kTombstone;
                }
            }
#pragma unroll
            for (int g = 0; g < kPGroups; ++g) {
                if (static_cast<uint32_t>(g) >= m_groups) break;
                const long long c0 = kDebug ? clock64() : 0;
                mbar_wait_a(a_tmem_full + 8 * g, t & 1);
                if (kDebug) w_tfull += clock64() - c0;
                tc_fence_after();
                const uint32_t taddr = tmem + ((lq * 32u) << 16) + g * kPN + cb * 64;
                const bool live = g ? live1 : live0;
                RegList &list = g ? list1 : list0;
                uint64_t &thr = g ? thr1 : thr0;
                // timeline (debug build): accumulator seen ready / handed back / values done, first and last epilogue warp
                const bool tl = kDebug && a.debug && blockIdx.x == 0 && blockIdx.y == 0 && t >= 16 && t < 24 && lane == 0 && (warp == 0 || warp == 15);
                unsigned long long *tl_d = a.debug + 2048 + ((t - 16) * 2 + g) * 8 + (warp == 0 ? 1 : 4);
                if (tl) tl_d[0] = clock64();
                // The next tile's MMAs into this accumulator wait for the SLOWEST of the pair's 32 epilogue warps, and in
                // nearly every tile some warp meets a candidate in its first 32 columns.  So while the accumulator is held
                // only the candidate is FOUND (one key per lane); the list work (~1000 cycles for the warp) waits until the
                // second read is in registers and the accumulator has been handed back.
                if (kDebug && a.dbg_mode) {                  // pace probes (wrong results): see GemmArgs.dbg_mode
                    uint32_t w[32];
                    int mm = 0;
                    if (a.dbg_mode >= 2) { tmem_ld32(taddr, w); if (a.dbg_mode >= 3) mm = max32(w); tmem_ld32(taddr + 32, w); }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) remote_arrive(g ? leader_empty1 : leader_empty0);
                    if (a.dbg_mode == 2 && w[lane] == 0xDEADBEEFu) run0 += 1;      // keeps the loads alive
                    if (a.dbg_mode >= 3) {
                        mm = max(mm, max32(w));
                        run0 = max(run0, mm);
                        if (a.dbg_mode >= 4 && __any_sync(kFull, static_cast<uint32_t>(mm) >= static_cast<uint32_t>((g ? thr1 : thr0) >> 32))) run1 += 1;
                    }
                    continue;
                }
                uint32_t v[32];
                tmem_ld32(taddr, v);
                if (kFloorPass) {
                    if (ok_mask[0] != kFull) {               // warp-uniform: a tile wholly inside the scope skips the masking
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = ((ok_mask[0] >> j) & 1u) ? v[j] : 0u;
                    }
                    const int m0 = max32(v);
                    tmem_ld32(taddr + 32, v);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) remote_arrive(g ? leader_empty1 : leader_empty0);
                    if (ok_mask[1] != kFull) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = ((ok_mask[1] >> j) & 1u) ? v[j] : 0u;
                    }
                    const int m1 = max32(v);
                    if (g) run1 = __vimax3_s32(run1, m0, m1);
                    else run0 = __vimax3_s32(run0, m0, m1);
                    continue;
                }
                bool multi;
                uint64_t first_key = group_candidate(v, ok_mask[0], live, gmm, a.id_base + row0, thr, multi);
                if (multi) {                                   // rare: list work while holding (it covers every lane's candidates)
                    take_group_multi(v, ok_mask[0], live, a.id_base + row0, list, thr);
                    first_key = 0ull;
                }
                tmem_ld32(taddr + 32, v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) remote_arrive(g ? leader_empty1 : leader_empty0);
                if (tl) tl_d[1] = clock64();
                offer_key(first_key, list, thr);
                const uint64_t second_key = group_candidate(v, ok_mask[1], live, gmm, a.id_base + row0 + 32, thr, multi);
                if (multi) take_group_multi(v, ok_mask[1], live, a.id_base + row0 + 32, list, thr);
                else offer_key(second_key, list, thr);
                if (tl) tl_d[2] = clock64();
            }
        }
        if (kDebug && a.debug && warp == 0 && lane == 0) {
            unsigned long long *d = a.debug + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 8;
            d[4] = clock64() - e_start; d[5] = w_tfull; 
        }
        if (kFloorPass) {
            // one group maximum per (slice, column block, query): [n_slices * 4][nq] u32 in the list buffer
            uint32_t *vals = reinterpret_cast<uint32_t *>(a.out_lists);
            const size_t grp = static_cast<size_t>(slice) * kPColBlocks + cb;
            if (live0) vals[grp * a.nq + q0] = static_cast<uint32_t>(run0);
            if (live1) vals[grp * a.nq + q0 + 256] = static_cast<uint32_t>(run1);
        } else {
        // ---- the four column-block lists of a query -> one list, through the (now idle) feature ring:
        // every tile's MMAs have retired (the last "accumulator ready" was awaited above), so neither
        // the tensor core nor the TMA unit touches the ring any more
        uint64_t *stage = reinterpret_cast<uint64_t *>(&sm.b[0][0][0]);           // [g][cb][128 rows][k]: 80 KB of the 128 KB ring
        const uint32_t row_in_cta = lq * 32 + lane;
#pragma unroll
        for (int i = 0; i < kGemmK; ++i) {
            stage[((0 * kPColBlocks + cb) * 128 + row_in_cta) * kGemmK + i] = list0.e[i];
            stage[((1 * kPColBlocks + cb) * 128 + row_in_cta) * kGemmK + i] = list1.e[i];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kPEpiWarps * 32) : "memory");       // the 16 epilogue warps only
        if (cb < kPGroups) {
            // warps with cb = g merge M-group g: a 4-way merge of sorted lists, heads in registers
            const uint32_t g = cb;
            const uint32_t q = q0 + 256 * g;
            if (q < a.nq) {
                const uint64_t *src = stage + (static_cast<size_t>(g) * kPColBlocks * 128 + row_in_cta) * kGemmK;
                uint32_t pos[kPColBlocks];
                uint64_t head[kPColBlocks];
#pragma unroll
                for (int c = 0; c < kPColBlocks; ++c) { pos[c] = 0; head[c] = src[static_cast<size_t>(c) * 128 * kGemmK]; }
                uint64_t *dst = a.out_lists + (static_cast<size_t>(slice) * a.nq + q) * kGemmK;
#pragma unroll
                for (int i = 0; i < kGemmK; ++i) {
                    uint64_t best = head[0];
                    int bc = 0;
#pragma unroll
                    for (int c = 1; c < kPColBlocks; ++c)
                        if (head[c] > best) { best = head[c]; bc = c; }
                    dst[i] = best;                                                 // 0 once every list has run dry
#pragma unroll
                    for (int c = 0; c < kPColBlocks; ++c)
                        if (c == bc) {
                            ++pos[c];
                            head[c] = pos[c] < static_cast<uint32_t>(kGemmK) ? src[static_cast<size_t>(c) * 128 * kGemmK + pos[c]] : 0ull;
                        }
                }
            }
        }
        }   // !kFloorPass
    }
    if (n_tiles == 0 && warp < 4 * kPGroups) {
        // a slice without tiles still owes its (empty) lists / group maxima
        const uint32_t lq = warp & 3, g = static_cast<uint32_t>(warp) >> 2;
        const uint32_t q = q_base + rank * 128 + lq * 32 + lane + 256 * g;
        if (q < a.nq) {
            if (kFloorPass) {
                uint32_t *vals = reinterpret_cast<uint32_t *>(a.out_lists);
                for (int c = 0; c < kPColBlocks; ++c) vals[(static_cast<size_t>(slice) * kPColBlocks + c) * a.nq + q] = 0u;
            } else {
                uint64_t *dst = a.out_lists + (static_cast<size_t>(slice) * a.nq + q) * kGemmK;
                for (int i = 0; i < kGemmK; ++i) dst[i] = 0ull;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();     // the peer's shared memory and barriers stay valid until both CTAs are done
    if (warp == kPProducerWarp) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ---- wider rows (K = 512, 1024): the same pair kernel with the K dimension streamed ------------------------------------
// A row is kSlabs 256-byte K-slabs (2 or 4).  A CTA keeps its 128 queries of ONE M-group resident with all of K (64 / 128
// KB), so a pair owns 256 queries; a chunk tile's B operand arrives as kSlabs ring stages (one K-slab of this CTA's 128
// rows each, 32 KB, three stages), the tile's kSlabs x 8 MMAs accumulate into one 256-column accumulator and the two
// accumulators alternate by TILE.  Per tile the epilogue has one accumulator to drain per 2048 / 4096 cycles of MMA time, so
// it is never the limit here; with 256 queries per pair the pass is HBM-bound instead (128 / 256 KB of features per pair and
// tile): a batch of up to 256 queries costs one streaming pass over the scope, larger batches one pass per 256 queries
// that mostly hits L2 (the query groups of a slice run side by side).
constexpr int kWMaxSlabs = 4;
constexpr int kWStages = 3;
template <int kSlabs>
struct WideSmemT {
    alignas(1024) uint8_t q[kSlabs][2][kTileKBlock];       // 64 / 128 KB: this CTA's 128 queries, every K-slab
    alignas(1024) uint8_t b[kWStages][2][kTileKBlock];     // 96 KB: K-slabs of this CTA's half of the chunk tiles
    alignas(8) uint64_t q_full;                            // leader's copy in use
    uint64_t full[kWStages];                               // leader's copy in use
    uint64_t empty[kWStages];                              // own copy (multicast commit)
    uint64_t tmem_full[2];                                 // own copy (multicast commit)
    uint64_t tmem_empty[2];                                // leader's copy in use (2 x 16 warp arrivals)
    uint32_t tmem_base;
};

template <bool kFloorPass, int kWSlabs>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPThreads, 1)
score_topk_gemm_wide_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_f, const GemmArgs a) {
    static_assert(kWSlabs == 2 || kWSlabs == 4, "rows of 512 or 1024 features");
    using WideSmem = WideSmemT<kWSlabs>;
    extern __shared__ __align__(1024) uint8_t pair_smem_raw[];
    WideSmem &sm = *reinterpret_cast<WideSmem *>((reinterpret_cast<uintptr_t>(pair_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t slice = blockIdx.x >> 1, n_slices = gridDim.x >> 1, qgroup = blockIdx.y;
    const uint32_t total_tiles = (a.row_hi - a.row_lo + kPN - 1) / kPN;
    const uint32_t t_lo = static_cast<uint32_t>(static_cast<uint64_t>(total_tiles) * slice / n_slices);
    const uint32_t t_hi = static_cast<uint32_t>(static_cast<uint64_t>(total_tiles) * (slice + 1) / n_slices);
    const uint32_t n_tiles = t_hi - t_lo;
    const uint32_t q_base = qgroup * 256;

    if (threadIdx.x == 0) {
        mbar_init(&sm.q_full, 2);
        for (int s = 0; s < kWStages; ++s) { mbar_init(&sm.full[s], 2); mbar_init(&sm.empty[s], 1); }
        for (int g = 0; g < 2; ++g) { mbar_init(&sm.tmem_full[g], 1); mbar_init(&sm.tmem_empty[g], 2 * kPEpiWarps); }
        mbar_fence_init();
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == kPProducerWarp) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    uint32_t sm_a;
    {
        const uint32_t raw_a = smem_u32(pair_smem_raw);
        const uint32_t aligned = (raw_a + 1023u) & ~1023u;
        asm volatile("mov.u32 %0, %1;" : "=r"(sm_a) : "r"(aligned));
    }
    const uint32_t a_q_full = sm_a + static_cast<uint32_t>(offsetof(WideSmem, q_full));
    const uint32_t a_full = sm_a + static_cast<uint32_t>(offsetof(WideSmem, full));
    const uint32_t a_empty = sm_a + static_cast<uint32_t>(offsetof(WideSmem, empty));
    const uint32_t a_tmem_full = sm_a + static_cast<uint32_t>(offsetof(WideSmem, tmem_full));
    const uint32_t a_tmem_empty = sm_a + static_cast<uint32_t>(offsetof(WideSmem, tmem_empty));

    if (warp == kPProducerWarp) {
        // ===== TMA producer (both CTAs) =====
        if (lane == 0 && n_tiles) {
            const uint32_t leader_q_full = map_to_cta_a(a_q_full, 0);
            remote_arrive_expect_tx(leader_q_full, kWSlabs * 2 * kTileKBlock);
            for (int sl = 0; sl < kWSlabs; ++sl)
                for (int kb = 0; kb < 2; ++kb)
                    tma_load_2d_pair(sm.q[sl][kb], &map_q, sl * 256 + kb * kKBlockBytes, static_cast<int>(q_base + rank * 128), leader_q_full);
            for (uint32_t i = 0; i < n_tiles * kWSlabs; ++i) {            // i: (tile, K-slab) in issue order
                const uint32_t t = i / kWSlabs, sl = i % kWSlabs, s = i % kWStages;
                if (i >= kWStages) mbar_wait_a(a_empty + 8 * s, ((i / kWStages) - 1) & 1);
                const uint32_t row0 = a.row_lo + (t_lo + t) * kPN + rank * 128;
                const uint32_t leader_full = map_to_cta_a(a_full + 8 * s, 0);
                remote_arrive_expect_tx(leader_full, 2 * kTileKBlock);
                for (int kb = 0; kb < 2; ++kb)
                    tma_load_2d_pair(sm.b[s][kb], &map_f, static_cast<int>(sl * 256 + kb * kKBlockBytes), static_cast<int>(row0), leader_full);
            }
        }
        __syncwarp();
    } else if (warp == kPMmaWarp) {
        // ===== MMA issuer (leader CTA) =====
        if (rank == 0 && n_tiles) {
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kPN >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
            constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);        // SBO = 1024 B, version 1, SWIZZLE_128B
            const uint32_t q_lo0 = ((smem_u32(&sm.q[0][0][0]) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t b_lo0 = ((smem_u32(&sm.b[0][0][0]) & 0x3FFFFu) >> 4) | (1u << 16);
            constexpr uint32_t kKBlockStep = kTileKBlock >> 4;
            mbar_wait_a(a_q_full, 0);
            for (uint32_t t = 0; t < n_tiles; ++t) {
                const uint32_t acc = t & 1u;
                if (t >= 2) mbar_wait_a(a_tmem_empty + 8 * acc, ((t >> 1) - 1) & 1);    // both CTAs' epilogues drained this accumulator
                for (uint32_t sl = 0; sl < kWSlabs; ++sl) {
                    const uint32_t i = t * kWSlabs + sl, s = i % kWStages;
                    mbar_wait_a(a_full + 8 * s, (i / kWStages) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t q_lo = q_lo0 + sl * 2 * kKBlockStep, b_lo = b_lo0 + s * 2 * kKBlockStep;
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t da = (static_cast<uint64_t>(kDescHi) << 32) | (q_lo + kb * kKBlockStep + 2u * k);
                                const uint64_t db = (static_cast<uint64_t>(kDescHi) << 32) | (b_lo + kb * kKBlockStep + 2u * k);
                                if (sl | kb | k) umma2_i8<true>(tmem + acc * kPN, da, db, idesc);
                                else umma2_i8<false>(tmem + acc * kPN, da, db, idesc);
                            }
                        }
                        umma2_commit_both_a(a_empty + 8 * s);              // the K-slab's stage is free once these MMAs retire
                        if (sl == kWSlabs - 1) umma2_commit_both_a(a_tmem_full + 8 * acc);
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    } else if (n_tiles) {
        // ===== epilogue (both CTAs): warp -> lane quarter lq, column block cb; thread -> one query =====
        const uint32_t lq = warp & 3;
        const uint32_t cb = static_cast<uint32_t>(warp) >> 2;
        const uint32_t n_scope = a.n_scope;
        const uint32_t q0 = q_base + rank * 128 + lq * 32 + lane;
        const bool live = q0 < a.nq;
        RegList list;
        list.clear();
        uint64_t thr = (a.floors && live) ? a.floors[q0] : 0ull;
        int run = 0;
        const uint32_t leader_empty[2] = {map_to_cta_a(a_tmem_empty, 0), map_to_cta_a(a_tmem_empty + 8, 0)};
        const uint32_t sc0 = a.scope[0], sc1 = a.scope[1], sc2 = a.scope[2], sc3 = a.scope[3];
        const uint32_t row_hi = a.row_hi;
        const uint32_t *__restrict__ seg_words = a.seg;
        auto in_scope = [&](uint32_t sg) {
            bool ok = (sg == sc0) | (sg == sc1) | (sg == sc2) | (sg == sc3);
            if (n_scope > 4)
                for (uint32_t x = 4; x < n_scope; ++x) ok |= (sg == a.scope[x]);
            return ok && sg != kTombstone;
        };
        uint32_t seg_next[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t row = a.row_lo + t_lo * kPN + cb * 64 + h * 32 + lane;
            seg_next[h] = row < row_hi ? __ldg(seg_words + row) : kTombstone;
        }
        for (uint32_t t = 0; t < n_tiles; ++t) {
            const uint32_t row0 = a.row_lo + (t_lo + t) * kPN + cb * 64;
            uint32_t ok_mask[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t row = row0 + h * 32 + lane;
                ok_mask[h] = __ballot_sync(kFull, row < row_hi && in_scope(seg_next[h]));
            }
            if (t + 1 < n_tiles) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t row = row0 + kPN + h * 32 + lane;
                    seg_next[h] = row < row_hi ? __ldg(seg_words + row) : kTombstone;
                }
            }
            const uint32_t acc = t & 1u;
            mbar_wait_a(a_tmem_full + 8 * acc, (t >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + ((lq * 32u) << 16) + acc * kPN + cb * 64;
            uint32_t v[32];
            tmem_ld32(taddr, v);
            if (kFloorPass) {
                if (ok_mask[0] != kFull) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = ((ok_mask[0] >> j) & 1u) ? v[j] : 0u;
                }
                const int m0 = max32(v);
                tmem_ld32(taddr + 32, v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) remote_arrive(leader_empty[acc]);
                if (ok_mask[1] != kFull) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = ((ok_mask[1] >> j) & 1u) ? v[j] : 0u;
                }
                run = __vimax3_s32(run, m0, max32(v));
                continue;
            }
            bool multi;
            uint64_t first_key = group_candidate(v, ok_mask[0], live, false, a.id_base + row0, thr, multi);
            if (multi) {
                take_group_multi(v, ok_mask[0], live, a.id_base + row0, list, thr);
                first_key = 0ull;
            }
            tmem_ld32(taddr + 32, v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) remote_arrive(leader_empty[acc]);
            offer_key(first_key, list, thr);
            const uint64_t second_key = group_candidate(v, ok_mask[1], live, false, a.id_base + row0 + 32, thr, multi);
            if (multi) take_group_multi(v, ok_mask[1], live, a.id_base + row0 + 32, list, thr);
            else offer_key(second_key, list, thr);
        }
        if (kFloorPass) {
            uint32_t *vals = reinterpret_cast<uint32_t *>(a.out_lists);
            if (live) vals[(static_cast<size_t>(slice) * kPColBlocks + cb) * a.nq + q0] = static_cast<uint32_t>(run);
        } else {
            // the four column-block lists of a query -> one list, through the (now idle) feature ring
            uint64_t *stage = reinterpret_cast<uint64_t *>(&sm.b[0][0][0]);       // [cb][128 rows][k]: 40 KB of the 96 KB ring
            const uint32_t row_in_cta = lq * 32 + lane;
#pragma unroll
            for (int i = 0; i < kGemmK; ++i) stage[(cb * 128 + row_in_cta) * kGemmK + i] = list.e[i];
            asm volatile("bar.sync 1, %0;" ::"n"(kPEpiWarps * 32) : "memory");
            if (cb == 0 && live) {
                const uint64_t *src = stage + static_cast<size_t>(row_in_cta) * kGemmK;
                uint32_t pos[kPColBlocks];
                uint64_t head[kPColBlocks];
#pragma unroll
                for (int c = 0; c < kPColBlocks; ++c) { pos[c] = 0; head[c] = src[static_cast<size_t>(c) * 128 * kGemmK]; }
                uint64_t *dst = a.out_lists + (static_cast<size_t>(slice) * a.nq + q0) * kGemmK;
#pragma unroll
                for (int i = 0; i < kGemmK; ++i) {
                    uint64_t best = head[0];
                    int bc = 0;
#pragma unroll
                    for (int c = 1; c < kPColBlocks; ++c)
                        if (head[c] > best) { best = head[c]; bc = c; }
                    dst[i] = best;
#pragma unroll
                    for (int c = 0; c < kPColBlocks; ++c)
                        if (c == bc) {
                            ++pos[c];
                            head[c] = pos[c] < static_cast<uint32_t>(kGemmK) ? src[static_cast<size_t>(c) * 128 * kGemmK + pos[c]] : 0ull;
                        }
                }
            }
        }
    }
    if (n_tiles == 0 && warp < 4) {
        // a slice without tiles still owes its (empty) list / group maxima
        const uint32_t q = q_base + rank * 128 + (warp & 3) * 32 + lane;
        if (q < a.nq) {
            if (kFloorPass) {
                uint32_t *vals = reinterpret_cast<uint32_t *>(a.out_lists);
                for (int c = 0; c < kPColBlocks; ++c) vals[(static_cast<size_t>(slice) * kPColBlocks + c) * a.nq + q] = 0u;
            } else {
                uint64_t *dst = a.out_lists + (static_cast<size_t>(slice) * a.nq + q) * kGemmK;
                for (int i = 0; i < kGemmK; ++i) dst[i] = 0ull;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == kPProducerWarp) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// floors[q] = (k-th largest of vals[0 .. n_groups)[q]) << 32: the floor pass's group maxima -> one lower-bound key per
// query (score word only).  One warp per query; the values sit in shared memory, k rounds of lane maximum ->
// warp maximum -> the lowest lane holding it retires ONE instance (equal maxima of different groups count separately:
// each names a different chunk).
constexpr int kKthWarps = 4;
__global__ void __launch_bounds__(kKthWarps * 32) kth_largest_kernel(const uint32_t *__restrict__ vals, uint32_t n_groups, uint32_t nq, uint32_t k,
                                                                  uint64_t *__restrict__ floors) {
    extern __shared__ uint32_t kth_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * kKthWarps + w;
    if (q >= nq) return;
    uint32_t *mine = kth_smem + static_cast<size_t>(w) * n_groups;
    for (uint32_t i = lane; i < n_groups; i += 32) mine[i] = vals[static_cast<size_t>(i) * nq + q];
    __syncwarp();
    uint32_t kth = 0;
    for (uint32_t r = 0; r < k; ++r) {
        uint32_t best = 0, at = 0;
        for (uint32_t i = lane; i < n_groups; i += 32) {
            const uint32_t x = mine[i];
            if (x > best) { best = x; at = i; }
        }
        kth = __reduce_max_sync(kFull, best);
        if (kth == 0u) break;                                        // fewer than k groups with a positive maximum
        const uint32_t owners = __ballot_sync(kFull, best == kth);
        if (lane == __ffs(owners) - 1) mine[at] = 0u;
        __syncwarp();
    }
    if (lane == 0) floors[q] = static_cast<uint64_t>(kth) << 32;
}

}  // namespace

cudaError_t launch_kth_largest(const uint32_t *vals, uint32_t n_groups, uint32_t nq, uint32_t k, uint64_t *floors, cudaStream_t s) {
    if (n_groups == 0 || n_groups > 8192 || k == 0 || nq == 0) return cudaErrorInvalidValue;
    const size_t smem = static_cast<size_t>(kKthWarps) * n_groups * 4;
    if (smem > 48 * 1024)
        if (cudaError_t e = ensure_dynamic_smem(kth_largest_kernel, static_cast<int>(smem)); e != cudaSuccess) return e;
    kth_largest_kernel<<<(nq + kKthWarps - 1) / kKthWarps, kKthWarps * 32, smem, s>>>(vals, n_groups, nq, k, floors);
    return cudaGetLastError();
}
uint32_t gemm_pair_floor_groups(uint32_t n_slices) { return n_slices * kPColBlocks; }

size_t gemm_pair_lists_bytes(uint32_t n_slices, uint32_t nq) { return static_cast<size_t>(n_slices) * kGemmPairLists * nq * kGemmK * 8; }

cudaError_t launch_score_topk_gemm_pair(const GemmArgs &a, const int8_t *q_dev, const int8_t *F, uint64_t f_rows, uint32_t n_slices,
                                        cudaStream_t s) {
    CUtensorMap map_q, map_f;
    if (!make_map(&map_q, q_dev, a.nq) || !make_map(&map_f, F, f_rows)) return cudaErrorNotSupported;
    const int smem = static_cast<int>(sizeof(PairSmem)) + 1024;
    dim3 grid(2 * n_slices, (a.nq + kPGroups * 256 - 1) / (kPGroups * 256), 1);   // x: CTA pairs (cluster dims 2 x 1 x 1)
    auto go = [&](auto kern) -> cudaError_t {
        if (cudaError_t e = ensure_dynamic_smem(kern, smem); e != cudaSuccess) return e;
        kern<<<grid, kPThreads, smem, s>>>(map_q, map_f, a);
        return cudaGetLastError();
    };
    // floor pass: group maxima only ([n_slices * 4][nq] u32 in out_lists) / main pass: one sorted list per (slice, query)
    if (a.group_max_mode) return a.debug ? go(score_topk_gemm_pair_kernel<true, true>) : go(score_topk_gemm_pair_kernel<false, true>);
    return a.debug ? go(score_topk_gemm_pair_kernel<true, false>) : go(score_topk_gemm_pair_kernel<false, false>);
}

// wider rows: [rows, dim] int8 -> boxes of 128 rows x 128 bytes, 128-byte swizzle
static bool make_map_wide(CUtensorMap *map, const void *base, uint64_t rows, uint32_t dim) {
    gemm::EncodeTiledFn enc = gemm::encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {dim, rows};
    cuuint64_t strides[1] = {dim};
    cuuint32_t box[2] = {128, 128};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t gemm_wide_lists_bytes(uint32_t n_slices, uint32_t nq) { return static_cast<size_t>(n_slices) * nq * gemm::kGemmK * 8; }

cudaError_t launch_score_topk_gemm_wide(const GemmArgs &a, const int8_t *q_dev, const int8_t *F, uint64_t f_rows, uint32_t dim, uint32_t n_slices,
                                        cudaStream_t s) {
    if (dim != 512 && dim != 1024) return cudaErrorInvalidValue;
    CUtensorMap map_q, map_f;
    if (!make_map_wide(&map_q, q_dev, a.nq, dim) || !make_map_wide(&map_f, F, f_rows, dim)) return cudaErrorNotSupported;
    dim3 grid(2 * n_slices, (a.nq + 255) / 256, 1);
    auto go = [&](auto kern, int smem) -> cudaError_t {
        if (cudaError_t e = ensure_dynamic_smem(kern, smem); e != cudaSuccess) return e;
        kern<<<grid, kPThreads, smem, s>>>(map_q, map_f, a);
        return cudaGetLastError();
    };
    if (dim == 512) {
        const int smem = static_cast<int>(sizeof(WideSmemT<2>)) + 1024;
        return a.group_max_mode ? go(score_topk_gemm_wide_kernel<true, 2>, smem) : go(score_topk_gemm_wide_kernel<false, 2>, smem);
    }
    const int smem = static_cast<int>(sizeof(WideSmemT<4>)) + 1024;
    return a.group_max_mode ? go(score_topk_gemm_wide_kernel<true, 4>, smem) : go(score_topk_gemm_wide_kernel<false, 4>, smem);
}

}  // namespace rf
