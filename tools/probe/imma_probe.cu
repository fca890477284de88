// Probe: peak rate of the legacy warp-level tensor path (mma.sync m16n8k32 s8) on sm_100a, and of
// tcgen05.ld in several shapes -- numbers that decide whether a register-accumulator epilogue
// could beat the TMEM read-back wall of score_topk_gemm.  Development tool.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(256) imma_loop(int iters, int *sink) {
    int a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = threadIdx.x ^ 5, b1 = 11;
    int c[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    if (s == 123456789) *sink = s;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int kShape>
__global__ void __launch_bounds__(512) ldtm_loop(int iters, int *sink, long long *cycles) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 128u;
    uint32_t acc = 0;
    const long long c0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t v[32];
        if (kShape == 0) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(t + (uint32_t)((i & 3) * 32)));
        } else {
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(t + (uint32_t)((i & 3) * 32)));
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(t + (uint32_t)((i & 3) * 32) + (16u << 16)));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j];
    }
    const long long c1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = c1 - c0;
    if (acc == 0x12345678u) *sink = (int)acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

int main() {
    int *sink; long long *cyc;
    CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&cyc, 148 * 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps_per_sm = 8; warps_per_sm <= 32; warps_per_sm *= 2) {
        const int iters = 20000, blocks = 148 * (warps_per_sm / 8);
        imma_loop<<<blocks, 256>>>(100, sink);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        imma_loop<<<blocks, 256>>>(iters, sink);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double ops = 2.0 * 16 * 8 * 32 * 8.0 * iters * (double)blocks * 8;
        printf("mma.sync m16n8k32 s8, %2d warps/SM: %.1f TOP/s\n", warps_per_sm, ops / (ms * 1e-3) / 1e12);
    }
    for (int shape = 0; shape < 2; ++shape)
        for (int threads = 128; threads <= 512; threads *= 2) {
            const int iters = 20000;
            if (shape == 0) ldtm_loop<0><<<148, threads>>>(iters, sink, cyc); else ldtm_loop<1><<<148, threads>>>(iters, sink, cyc);
            CK(cudaDeviceSynchronize());
            long long h[148]; CK(cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost));
            const double bytes = (double)iters * (threads / 32) * 32 * 32 * 4;
            printf("tcgen05.ld %s, %2d warps/SM: %.1f B/clk/SM\n", shape == 0 ? "32x32b.x32" : "16x256b.x4 (x2)", threads / 32, bytes / (double)h[0]);
        }
    return 0;
}
