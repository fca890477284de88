// Probe: layout of tcgen05.ld.32x32b.x32.pack::16b -- which TMEM columns land in which register halves.
// Each thread writes 64 columns of its lane with tcgen05.st (value = 0xABCD0000 + 256 * (lane & 63) + column),
// reads them back packed, and the host prints the mapping.  Development tool.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128) probe(uint32_t *out) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 64; c += 8) {
        uint32_t w[8];
        for (int j = 0; j < 8; ++j) w[j] = 0xABCD0000u + 256u * (uint32_t)(lane) + (uint32_t)(c + j);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(t + c), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(t));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) out[threadIdx.x * 32 + i] = v[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64) : "memory");
}
int main() {
    uint32_t *d, h[128 * 32];
    CK(cudaMalloc(&d, sizeof h));
    probe<<<1, 128>>>(d);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
    for (int thr : {0, 37}) {
        printf("thread %d (lane %d):", thr, thr & 31);
        for (int i = 0; i < 32; ++i) printf(" r%d=%08x", i, h[thr * 32 + i]);
        printf("\n");
    }
    long bad = 0;
    for (int thr = 0; thr < 128; ++thr)
        for (int i = 0; i < 32; ++i) {
            const uint32_t lo = 256u * (thr & 31) + 2 * i, hi = 256u * (thr & 31) + 2 * i + 1;
            if (h[thr * 32 + i] != ((hi & 0xFFFF) << 16 | (lo & 0xFFFF))) ++bad;
        }
    printf("hypothesis r[i] = (col 2i+1 low16) << 16 | (col 2i low16): %ld mismatches of 4096\n", bad);
    return 0;
}
