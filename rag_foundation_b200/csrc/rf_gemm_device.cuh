// Device and host helpers shared by the batched tensor-core kernels (score_topk_gemm.cu: one CTA
// per tile, cta_group::1; score_topk_gemm_pair.cu: CTA pairs, cta_group::2).
#pragma once
#include <cuda.h>

#include "rf_device.cuh"
#include "rf_internal.h"

namespace rf {
namespace gemm {

constexpr int kGemmK = kGemmListK;          // list length kept per (thread, query)
constexpr int kKBlockBytes = 128;           // one SW128 swizzle row: 128 int8 of K
constexpr int kTileKBlock = 128 * kKBlockBytes;   // 16 KB: 128 rows x 128 B

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

// max of 32 non-negative scores as a depth-4 tree of three-input maxima (VIMNMX3): 16 operations like
// the running maximum, but no 16-long dependency chain in front of the candidate test
__device__ __forceinline__ int max32(const uint32_t (&v)[32]) {
    int m[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) m[i] = __vimax3_s32(static_cast<int>(v[3 * i]), static_cast<int>(v[3 * i + 1]), static_cast<int>(v[3 * i + 2]));
    m[10] = max(static_cast<int>(v[30]), static_cast<int>(v[31]));
    const int a = __vimax3_s32(m[0], m[1], m[2]), b = __vimax3_s32(m[3], m[4], m[5]), c = __vimax3_s32(m[6], m[7], m[8]);
    return __vimax3_s32(__vimax3_s32(a, b, c), m[9], m[10]);
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows, 256] int8 row-major -> boxes of 128 rows x 128 bytes, 128-byte swizzle
inline bool make_map(CUtensorMap *map, const void *base, uint64_t rows) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {256, rows};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {128, 128};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace gemm
}  // namespace rf
