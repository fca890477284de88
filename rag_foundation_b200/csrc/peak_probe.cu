// rf_probe_int8_peak: the achievable dense int8 rate of this GPU's tensor pipe, measured the way
// MEASURED_PEAKS.json measures bf16 -- a kernel that does nothing but issue tcgen05.mma kind::i8 --
// so that the batched scoring kernel (BASELINE.json configs[2]; score_topk_gemm_pair.cu) has a
// roofline denominator recorded in the same run ("measure the achievable figure at build
// time").  One CTA pair per two SMs; the pair's leader issues 256 x 256 x 32 MMAs (cta_group::2, both
// operands in shared memory, SWIZZLE_128B descriptors -- the shape the scoring kernel uses) back to back
// into alternating TMEM accumulators and commits every 8; operand bytes are arbitrary (rates only).
// Timed with CUDA events around the launch; ops = 2 * M * N * K per MMA.
#include <cstdint>
#include <cuda_runtime.h>

#include "rf_device.cuh"
#include "rf_internal.h"

namespace rf {
namespace {

struct ProbeSmem {
    alignas(1024) uint8_t a[2][128 * 128];    // this CTA's 128 rows of two M = 256 groups, one 128-byte K-block
    alignas(1024) uint8_t b[2][128 * 128];    // this CTA's half of B, two stages
    alignas(8) uint64_t bar[4];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t probe_desc_sw128(const void *smem) {
    const uint64_t addr = (smem_u32(smem) & 0x3FFFFu) >> 4;
    return addr | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) int8_peak_probe_kernel(int n_batches) {
    extern __shared__ __align__(1024) uint8_t probe_raw[];
    ProbeSmem &s = *reinterpret_cast<ProbeSmem *>((reinterpret_cast<uintptr_t>(probe_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&s.bar[i], 1);
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(s.a) + sizeof(s.b)) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(s.a)[i] = 0x01010101u * (i & 3);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem_base;
    if (rank == 0 && warp == 0 && lane == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(256 >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
        for (int bt = 0; bt < n_batches; ++bt) {
            if (bt >= 4) mbar_wait(&s.bar[bt & 3], ((bt >> 2) - 1) & 1);
            const uint32_t d = tmem + static_cast<uint32_t>((bt & 1) * 256);
            for (int k = 0; k < 8; ++k) {
                const uint64_t da = probe_desc_sw128(s.a[bt & 1]) + static_cast<uint64_t>((k & 3) * 2);
                const uint64_t db = probe_desc_sw128(s.b[k >> 2]) + static_cast<uint64_t>((k & 3) * 2);
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n}" ::"r"(d),
                    "l"(da), "l"(db), "r"(idesc), "r"(k ? 1u : 0u), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0)
                    : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                             smem_u32(&s.bar[bt & 3])),
                         "h"(static_cast<uint16_t>(1))
                         : "memory");
        }
        for (int bt = n_batches - 4 > 0 ? n_batches - 4 : 0; bt < n_batches; ++bt) mbar_wait(&s.bar[bt & 3], (bt >> 2) & 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace
}  // namespace rf

extern "C" int rf_probe_int8_peak(int device, uint32_t n_batches, double *ops_per_s, double *ms) {
    if (!ops_per_s || n_batches < 8) return RF_EINVAL;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return RF_ENODEVICE; }
    if (prop.major != 10) return RF_ENODEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return RF_ECUDA;
    const int blocks = prop.multiProcessorCount & ~1;
    const int smem = static_cast<int>(sizeof(rf::ProbeSmem)) + 1024;
    if (rf::ensure_dynamic_smem(rf::int8_peak_probe_kernel, smem) != cudaSuccess) return RF_ECUDA;
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return RF_ECUDA;
    float best = 0.0f;
    for (int rep = 0; rep < 4; ++rep) {      // first launch warms up; best of the other three
        cudaEventRecord(e0, nullptr);
        rf::int8_peak_probe_kernel<<<blocks, 128, smem>>>(static_cast<int>(n_batches));
        cudaEventRecord(e1, nullptr);
        if (cudaEventSynchronize(e1) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            return RF_ECUDA;
        }
        float t = 0.0f;
        cudaEventElapsedTime(&t, e0, e1);
        if (rep && (best == 0.0f || t < best)) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double ops = static_cast<double>(blocks / 2) * n_batches * 8.0 * 2.0 * 256.0 * 256.0 * 32.0;
    *ops_per_s = ops / (best * 1e-3);
    if (ms) *ms = best;
    return RF_OK;
}
