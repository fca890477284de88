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
constexpr int kPThreads = (2 + kPEpiWarps) * 32;   // 576: TMA producer, MMA issuer, 16 epilogue warps
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

// One 32-column group of one query's scores -> the query's list.  Same logic as score_topk_gemm.cu.
__device__ __forceinline__ void take_group(const uint32_t (&v)[32], uint32_t okm, bool live, bool group_max_mode, uint32_t id0,
                                           RegList &list, uint64_t &thr) {
    if (group_max_mode) {
        int gm = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) gm = max(gm, ((okm >> j) & 1u) ? static_cast<int>(v[j]) : 0);
        const uint64_t key = pack_key(gm, id0);      // low word only makes groups distinct
        if (live && okm != 0u && key > thr) {
            list.insert(key);
            const uint64_t kth = list.e[kGemmK - 1];
            if (kth > thr) thr = kth;
        }
        return;
    }
    const int mx = max32(v);
    const uint32_t thr_s = static_cast<uint32_t>(thr >> 32);
    if (__any_sync(kFull, live && static_cast<uint32_t>(mx) >= thr_s)) {     // scores are in [0, 2^31): unsigned compare is exact
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
            if ((cand >> j) & 1u) {
                const uint64_t key = pack_key(static_cast<int32_t>(sc), id0 + j);
                if (key > thr) {
                    list.insert(key);
                    const uint64_t kth = list.e[kGemmK - 1];
                    if (kth > thr) thr = kth;
                }
            }
        }
    }
}

// ---- 16-bit packed epilogue -------------------------------------------------------------------------
// When every score of the batch fits 16 bits (sum of a query's counts x 127 <= 65535 for every query:
// query_bound_kernel below decides it on the device), the accumulator is read with tcgen05.ld ... .pack::16b:
// ONE load brings a warp's 64 columns as 32 registers, register i = (column 2i+1) << 16 | (column 2i)
// (tools/probe/tmem_pack_probe.cu), the accumulator goes back right after that load, and the maxima run on
// both halves at once (VIMNMX3.U16x2): half the loads and half the ALU work per score.
__device__ __forceinline__ void tmem_ld32_pack16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// per-half maxima: low half over the even columns, high half over the odd ones; m[i] = registers 3i .. 3i+2
__device__ __forceinline__ uint32_t max32_u16x2(const uint32_t (&v)[32], uint32_t (&m)[11]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) m[i] = __vimax3_u16x2(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    m[10] = __vimax3_u16x2(v[30], v[31], 0u);
    const uint32_t a = __vimax3_u16x2(m[0], m[1], m[2]), b = __vimax3_u16x2(m[3], m[4], m[5]), c = __vimax3_u16x2(m[6], m[7], m[8]);
    return __vimax3_u16x2(__vimax3_u16x2(a, b, c), m[9], m[10]);
}
__device__ __forceinline__ uint32_t max_half(uint32_t x) { return max(x & 0xFFFFu, x >> 16); }
// candidate bits of three packed registers (warp-uniform triple index -> compile-time register numbers)
#define RF_CAND(j)                                            \
    ce |= ((v[j] & 0xFFFFu) >= thr_s ? 1u : 0u) << (j);      \
    co |= ((v[j] >> 16) >= thr_s ? 1u : 0u) << (j);
#define RF_TRIPLE(i) \
    case i: RF_CAND(3 * (i)) RF_CAND(3 * (i) + 1) RF_CAND(3 * (i) + 2) break;
// 64 columns of one query: v packed as above; ok_e / ok_o bit i <=> column 2i / 2i+1 is in scope; id0 = chunk
// id of column 0.  thr of a padding row is all ones.
template <bool kGroupMax, bool kDbg = false>
__device__ __forceinline__ void take_packed(const uint32_t (&v)[32], uint32_t ok_e, uint32_t ok_o, uint32_t id0, RegList &list,
                                            uint64_t &thr, long long *dbg = nullptr) {
    uint32_t parts[11];
    if (kGroupMax) {
        // two groups per load: the even and the odd columns (disjoint chunk sets, which is all the floor
        // argument needs); out-of-scope chunks must not raise the bound
        uint32_t m;
        if ((ok_e & ok_o) == 0xFFFFFFFFu) {
            m = max32_u16x2(v, parts);
        } else {
            uint32_t w[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) w[i] = v[i] & (((ok_e >> i) & 1u) * 0xFFFFu + ((ok_o >> i) & 1u) * 0xFFFF0000u);
            m = max32_u16x2(w, parts);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const uint64_t key = pack_key(static_cast<int32_t>(half ? m >> 16 : m & 0xFFFFu), id0 + half);   // low word only makes groups distinct
            if ((half ? ok_o : ok_e) != 0u && key > thr) {
                list.insert(key);
                const uint64_t kth = list.e[kGemmK - 1];
                if (kth > thr) thr = kth;
            }
        }
        return;
    }
    const uint32_t m = max32_u16x2(v, parts);
    const uint32_t thr_s = static_cast<uint32_t>(thr >> 32);      // > 0xFFFF (padding row, or unreachable): the test fails
    if (__any_sync(kFull, max_half(m) >= thr_s)) {
        const long long dbg0 = kDbg ? clock64() : 0;
        // which register triples hold a candidate in some lane: 11 tests on the partial maxima, then only those
        // triples are compared element by element (warp-uniform walk)
        uint32_t t1 = 0;
#pragma unroll
        for (int i = 0; i < 11; ++i) t1 |= (max_half(parts[i]) >= thr_s ? 1u : 0u) << i;
        uint32_t uni_t = __reduce_or_sync(kFull, t1);
        uint32_t ce = 0, co = 0;
        while (uni_t) {
            const int tr = __ffs(uni_t) - 1;
            uni_t &= uni_t - 1;
            switch (tr) {
                RF_TRIPLE(0) RF_TRIPLE(1) RF_TRIPLE(2) RF_TRIPLE(3) RF_TRIPLE(4) RF_TRIPLE(5)
                RF_TRIPLE(6) RF_TRIPLE(7) RF_TRIPLE(8) RF_TRIPLE(9)
                default: RF_CAND(30) RF_CAND(31) break;
            }
        }
        ce &= ok_e;
        co &= ok_o;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const uint32_t mine = half ? co : ce;
            uint32_t uni = __reduce_or_sync(kFull, mine);
            while (uni) {
                const int i = __ffs(uni) - 1;
                uni &= uni - 1;
                const uint32_t x = pick32(v, i);                 // i is warp-uniform: a jump, not a TMEM reload
                if ((mine >> i) & 1u) {
                    const uint64_t key = pack_key(static_cast<int32_t>(half ? x >> 16 : x & 0xFFFFu), id0 + 2 * i + half);
                    if (key > thr) {
                        list.insert(key);
                        const uint64_t kth = list.e[kGemmK - 1];
                        if (kth > thr) thr = kth;
                    }
                }
            }
        }
        if (kDbg) { dbg[0] += clock64() - dbg0; dbg[1] += 1; }
    }
}
#undef RF_TRIPLE
#undef RF_CAND

// *flag |= 1 when some query's score bound does not fit 16 bits (or a query has a negative feature)
__global__ void __launch_bounds__(256) query_bound_kernel(const int8_t *__restrict__ q, uint32_t nq, uint32_t *__restrict__ flag) {
    const uint32_t qi = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (qi >= nq) return;
    const int2 x = *reinterpret_cast<const int2 *>(q + static_cast<size_t>(qi) * 256 + lane * 8);
    int sum = __dp4a(x.x, 0x01010101, __dp4a(x.y, 0x01010101, 0));
    const bool neg = ((static_cast<uint32_t>(x.x) | static_cast<uint32_t>(x.y)) & 0x80808080u) != 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
    const bool bad = __any_sync(kFull, neg) || sum * 127 > 0xFFFF;   // every score is a sum of q[d] * F[c, d], 0 <= F <= 127
    if (bad && lane == 0) atomicOr(flag, 1u);
}

// kDebug: in-kernel cycle counters (RF_SCAN_DEBUG=1, tools/gemm_timeline.py); the production instantiation
// carries no clock reads in its loops (they cost 5 % of the batch time)
// kPacked: the 16-bit packed epilogue.  Both instantiations are launched for a search; the one that does not
// match the batch's bound flag (a.pack_flag, written by query_bound_kernel) returns at once.
template <bool kDebug, bool kPacked>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPThreads, 1)
score_topk_gemm_pair_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_f, const GemmArgs a) {
    if (a.pack_flag && ((*a.pack_flag == 0u) != kPacked)) return;      // uniform over the whole grid
    if (!a.pack_flag && kPacked) return;
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
    if (warp == 0) {
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

    if (warp == 0) {
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
    } else if (warp == 1) {
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
        const uint32_t cb = static_cast<uint32_t>(warp - 2) >> 2;
        const uint32_t n_scope = a.n_scope;
        const uint32_t q0 = q_base + rank * 128 + lq * 32 + lane;          // + 256 g
        const bool live0 = q0 < a.nq, live1 = q0 + 256 < a.nq;
        RegList list0, list1;
        list0.clear();
        list1.clear();
        uint64_t thr0 = (a.floors && live0) ? a.floors[q0] : 0ull;
        uint64_t thr1 = (a.floors && live1) ? a.floors[q0 + 256] : 0ull;
        const uint32_t leader_empty0 = map_to_cta_a(a_tmem_empty, 0), leader_empty1 = map_to_cta_a(a_tmem_empty + 8, 0);
        const bool gmm = a.group_max_mode != 0;
        long long w_tfull = 0;
        long long dbgc[3] = {0, 0, 0};
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
        // tenant words of this warp's 64 columns, loaded one tile ahead: lane l holds columns l and 32 + l
        // (plain reads: two 32-column masks) or 2 l and 2 l + 1 (packed reads: even- and odd-column masks)
        const uint32_t c_a = kPacked ? 2 * lane : lane, c_b = kPacked ? 2 * lane + 1 : 32 + lane;
        uint32_t seg_next[2];
        {
            const uint32_t row = a.row_lo + t_lo * kPN + cb * 64;
            seg_next[0] = row + c_a < row_hi ? __ldg(seg_words + row + c_a) : kTombstone;
            seg_next[1] = row + c_b < row_hi ? __ldg(seg_words + row + c_b) : kTombstone;
        }
        if (kPacked) {       // padding rows never produce candidates: an unreachable threshold instead of a flag
            if (!live0) thr0 = ~0ull;
            if (!live1) thr1 = ~0ull;
        }
        long long w_mask = 0, w_arr = 0, w_ldx = 0;
        for (uint32_t t = 0; t < n_tiles; ++t) {
            const long long cm = kDebug ? clock64() : 0;
            const uint32_t row0 = a.row_lo + (t_lo + t) * kPN + cb * 64;   // chunk row of this warp's first column
            uint32_t ok_mask[2];
            ok_mask[0] = __ballot_sync(kFull, row0 + c_a < row_hi && in_scope(seg_next[0]));
            ok_mask[1] = __ballot_sync(kFull, row0 + c_b < row_hi && in_scope(seg_next[1]));
            if (t + 1 < n_tiles) {
                const uint32_t row = row0 + kPN;
                seg_next[0] = row + c_a < row_hi ? __ldg(seg_words + row + c_a) : kTombstone;
                seg_next[1] = row + c_b < row_hi ? __ldg(seg_words + row + c_b) : kTombstone;
            }
            if (kDebug) w_mask += clock64() - cm;
#pragma unroll
            for (int g = 0; g < kPGroups; ++g) {
                if (static_cast<uint32_t>(g) >= m_groups) break;
                const long long c0 = kDebug ? clock64() : 0;
                mbar_wait_a(a_tmem_full + 8 * g, t & 1);
                if (kDebug) w_tfull += clock64() - c0;
                tc_fence_after();
                const uint32_t taddr = tmem + ((lq * 32u) << 16) + g * kPN + cb * 64;
                uint32_t v[32];
                if (kPacked) {
                    const long long cl = kDebug ? clock64() : 0;
                    tmem_ld32_pack16(taddr, v);
                    const long long ca = kDebug ? clock64() : 0;
                    // the accumulator's only read is in registers: hand it back before working on the values
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) remote_arrive(g ? leader_empty1 : leader_empty0);
                    if (kDebug) { w_ldx += ca - cl; w_arr += clock64() - ca; }
                    if (gmm) take_packed<true>(v, ok_mask[0], ok_mask[1], a.id_base + row0, g ? list1 : list0, g ? thr1 : thr0);
                    else {
                        const long long k1 = kDebug ? clock64() : 0;
                        take_packed<false, kDebug>(v, ok_mask[0], ok_mask[1], a.id_base + row0, g ? list1 : list0, g ? thr1 : thr0, dbgc);
                        if (kDebug) dbgc[2] += clock64() - k1;
                    }
                } else {
                    tmem_ld32(taddr, v);
                    take_group(v, ok_mask[0], g ? live1 : live0, gmm, a.id_base + row0, g ? list1 : list0, g ? thr1 : thr0);
                    tmem_ld32(taddr + 32, v);
                    // the accumulator's last read is in registers: hand it back before working on the values
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) remote_arrive(g ? leader_empty1 : leader_empty0);
                    take_group(v, ok_mask[1], g ? live1 : live0, gmm, a.id_base + row0 + 32, g ? list1 : list0, g ? thr1 : thr0);
                }
            }
        }
        if (kDebug && a.debug && warp == 2 && lane == 0) {
            unsigned long long *d = a.debug + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 8;
            d[4] = clock64() - e_start; d[5] = w_tfull; d[6] = w_mask; d[7] = w_arr; d[1] = rank ? dbgc[2] : d[1]; d[2] = rank ? w_ldx : d[2];
        }
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
    }
    if (n_tiles == 0 && warp >= 2 && warp < 2 + 4 * kPGroups) {
        // a slice without tiles still owes its (empty) lists
        const uint32_t lq = warp & 3, g = static_cast<uint32_t>(warp - 2) >> 2;
        const uint32_t q = q_base + rank * 128 + lq * 32 + lane + 256 * g;
        if (q < a.nq) {
            uint64_t *dst = a.out_lists + (static_cast<size_t>(slice) * a.nq + q) * kGemmK;
            for (int i = 0; i < kGemmK; ++i) dst[i] = 0ull;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();     // the peer's shared memory and barriers stay valid until both CTAs are done
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace

size_t gemm_pair_lists_bytes(uint32_t n_slices, uint32_t nq) { return static_cast<size_t>(n_slices) * kGemmPairLists * nq * kGemmK * 8; }

cudaError_t launch_score_topk_gemm_pair(const GemmArgs &a, const int8_t *q_dev, const int8_t *F, uint64_t f_rows, uint32_t n_slices,
                                        cudaStream_t s) {
    CUtensorMap map_q, map_f;
    if (!make_map(&map_q, q_dev, a.nq) || !make_map(&map_f, F, f_rows)) return cudaErrorNotSupported;
    const int smem = static_cast<int>(sizeof(PairSmem)) + 1024;
    dim3 grid(2 * n_slices, (a.nq + kPGroups * 256 - 1) / (kPGroups * 256), 1);   // x: CTA pairs (cluster dims 2 x 1 x 1)
    auto go = [&](auto kernel) -> cudaError_t {
        if (cudaError_t e = ensure_dynamic_smem(kernel, smem); e != cudaSuccess) return e;
        kernel<<<grid, kPThreads, smem, s>>>(map_q, map_f, a);
        return cudaGetLastError();
    };
    // both epilogue variants are enqueued; the one that does not match the batch's bound flag returns at once
    if (cudaError_t e = a.debug ? go(score_topk_gemm_pair_kernel<true, false>) : go(score_topk_gemm_pair_kernel<false, false>); e != cudaSuccess)
        return e;
    if (a.pack_flag)
        if (cudaError_t e = a.debug ? go(score_topk_gemm_pair_kernel<true, true>) : go(score_topk_gemm_pair_kernel<false, true>); e != cudaSuccess)
            return e;
    return cudaGetLastError();
}

cudaError_t launch_query_bound(const int8_t *q_dev, uint32_t nq, uint32_t *flag, cudaStream_t s) {
    if (cudaError_t e = cudaMemsetAsync(flag, 0, 4, s); e != cudaSuccess) return e;
    query_bound_kernel<<<(nq + 7) / 8, 256, 0, s>>>(q_dev, nq, flag);
    return cudaGetLastError();
}

}  // namespace rf
