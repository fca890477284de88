// Device-side helpers shared by the RF-1 kernels (sm_100a).
//
// Key order (oracle/SPEC.md step 7): one unsigned 64-bit key per chunk,
//   (uint64(score) << 32) | (0xFFFFFFFF - global_chunk_id),
// so "score desc, chunk id asc" is a plain unsigned max and the selection is independent of how
// rows are spread over lanes, warps, blocks or GPUs.  Key 0 means "no result".
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rf {

// A chunk row is `dim` int8 features, dim = 256 * m (m = 1, 2, 4: rf_config.dim): m consecutive 256-byte
// SUB-ROWS.  The streaming kernels are written over sub-rows (a warp step always covers 32 of them = 8 KB)
// and are templated on m where the row boundary matters.
constexpr int kSubDim = 256;       // int8 features per sub-row
constexpr int kSubBytes = 256;
constexpr int kTileSubRows = 32;   // sub-rows one warp scores per step (8 KB) = 32 / m rows
constexpr int kMaxSub = 4;         // widest row: 1024 features
constexpr uint32_t kTombstone = 0xFFFFFFFFu;
constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ uint64_t pack_key(int32_t score, uint32_t gid) {
    return (static_cast<uint64_t>(static_cast<uint32_t>(score)) << 32) |
           static_cast<uint64_t>(0xFFFFFFFFu - gid);
}
__device__ __forceinline__ int32_t key_score(uint64_t key) { return static_cast<int32_t>(key >> 32); }
__device__ __forceinline__ uint32_t key_gid(uint64_t key) {
    return 0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFu);
}

// 128-bit streaming load: read-only path, do not allocate in L1 (each feature byte is used once).
__device__ __forceinline__ int4 ld_stream_v4(const int4 *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(kFull, static_cast<uint32_t>(v), src);
    uint32_t hi = __shfl_sync(kFull, static_cast<uint32_t>(v >> 32), src);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int delta) {
    uint32_t lo = __shfl_up_sync(kFull, static_cast<uint32_t>(v), delta);
    uint32_t hi = __shfl_up_sync(kFull, static_cast<uint32_t>(v >> 32), delta);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int mask) {
    uint32_t lo = __shfl_xor_sync(kFull, static_cast<uint32_t>(v), mask);
    uint32_t hi = __shfl_xor_sync(kFull, static_cast<uint32_t>(v >> 32), mask);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

// A warp keeps its running top-32 as one key per lane, sorted descending by lane index.  `thr` is
// the key at lane k-1 (0 until k keys have been seen); `floor` is a bound learnt from other warps
// (a published k-th best key of some other list of the same query): a key below either cannot be
// in the query's top-k, so it is dropped before it costs anything.
struct WarpTopK {
    uint64_t mine;   // this lane's entry of the sorted list
    uint64_t thr;    // warp-uniform: own k-th best key
    uint64_t floor;  // warp-uniform: external lower bound (monotone)

    __device__ __forceinline__ void reset() { mine = 0; thr = 0; floor = 0; }
    __device__ __forceinline__ uint64_t bar() const { return thr > floor ? thr : floor; }

    // First tile: take all 32 keys at once with a bitonic sort (descending) instead of 32 inserts.
    __device__ __forceinline__ void init_sorted(uint64_t key, int k, int lane) {
        uint64_t v = key;
#pragma unroll
        for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
            for (int j = k2 >> 1; j > 0; j >>= 1) {
                const uint64_t other = shfl_xor_u64(v, j);
                const bool desc_block = (lane & k2) == 0;
                const bool lower = (lane & j) == 0;
                const bool keep_max = (lower == desc_block);
                const uint64_t mx = v > other ? v : other;
                const uint64_t mn = v > other ? other : v;
                v = keep_max ? mx : mn;
            }
        }
        mine = v;
        thr = shfl_u64(mine, k - 1);
    }

    // Every lane offers one key (0 = nothing).  Rare path: scanned data is mostly below the bar.
    __device__ __forceinline__ void consume(uint64_t key, int k, int lane) {
        unsigned pending = __ballot_sync(kFull, key > bar());
        while (pending) {
            const int src = __ffs(pending) - 1;
            const uint64_t cand = shfl_u64(key, src);
            pending &= pending - 1;
            const unsigned dup = __ballot_sync(kFull, mine == cand);
            if (cand > bar() && dup == 0) {
                const int pos = __popc(__ballot_sync(kFull, mine > cand));  // sorted => a lane prefix
                const uint64_t up = shfl_up_u64(mine, 1);
                if (lane == pos) mine = cand;
                else if (lane > pos) mine = up;
                thr = shfl_u64(mine, k - 1);
                pending &= __ballot_sync(kFull, key > bar());
            }
        }
    }
};

// ---- mbarrier / bulk-copy (TMA engine, 1-D) helpers -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy executed by the TMA unit; completion is signalled on `bar` (bytes).
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace rf
