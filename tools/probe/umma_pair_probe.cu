// Probe: rate of tcgen05.mma.cta_group::2 kind::i8 (M = 256 over a CTA pair, N = 128 / 256, K = 32)
// with both operands in shared memory, alone and while 16 warps per CTA loop tcgen05.ld over the
// accumulators.  Rates only (operand data is arbitrary).  Development tool; see
// umma_contention_probe.cu for the single-CTA numbers.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw128(const void *smem) {
    const uint64_t addr = (smem_u32(smem) & 0x3FFFFu) >> 4;
    return addr | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma2_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}
__device__ __forceinline__ void commit2(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

struct Smem {
    alignas(1024) uint8_t a[2][128 * 128];    // this CTA's 128 rows of two M = 256 groups, one 128-byte K-block
    alignas(1024) uint8_t b[2][128 * 128];    // this CTA's half (<= 128 rows) of B, two stages
    alignas(8) uint64_t bar[4];
    uint32_t tmem_base;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(576) probe(int N, int ld_warps, int n_batches, unsigned int *stop, unsigned long long *out) {
    extern __shared__ __align__(1024) uint8_t raw[];
    Smem &s = *reinterpret_cast<Smem *>(raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    unsigned int *my_stop = stop + (blockIdx.x >> 1);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&s.bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < (int)(sizeof(s.a) + sizeof(s.b)) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(s.a)[i] = 0x01010101u * (i & 3);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem_base;
    const int n_acc = 512 / N;
    unsigned long long mma_cycles = 0, n_mma = 0;
    if (rank == 0 && warp == 0 && lane == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const long long c0 = clock64();
        for (int bt = 0; bt < n_batches; ++bt) {
            if (bt >= 4) mbar_wait(&s.bar[bt & 3], ((bt >> 2) - 1) & 1);
            const uint32_t d = tmem + (uint32_t)((bt % n_acc) * N);
            for (int k = 0; k < 8; ++k) {
                const uint64_t da = make_desc_sw128(s.a[bt & 1]) + (uint64_t)((k & 3) * 2);
                const uint64_t db = make_desc_sw128(s.b[k >> 2]) + (uint64_t)((k & 3) * 2);
                umma2_ss(d, da, db, idesc, k ? 1u : 0u);
            }
            commit2(&s.bar[bt & 3], 1);     // leader's barrier only
        }
        for (int bt = n_batches - 4 > 0 ? n_batches - 4 : 0; bt < n_batches; ++bt) mbar_wait(&s.bar[bt & 3], (bt >> 2) & 1);
        mma_cycles = clock64() - c0;
        n_mma = (unsigned long long)n_batches * 8;
        *reinterpret_cast<volatile unsigned int *>(my_stop) = 1;
    }
    unsigned long long ld_cycles = 0, n_ld = 0;
    uint32_t sink = 0;
    if (warp >= 2 && warp - 2 < ld_warps) {
        const int w = warp - 2;
        const uint32_t t = tmem + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)((w >> 2) * 128);
        const long long c0 = clock64();
        int i = 0;
        while (true) {
            if ((i & 15) == 15 && *reinterpret_cast<volatile unsigned int *>(my_stop)) break;
            uint32_t v[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(t + (uint32_t)((i & 3) * 32)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) sink = max(sink, v[j]);
            ++i;
        }
        ld_cycles = clock64() - c0;
        n_ld = i;
    }
    if (sink == 0xDEADBEEFu) out[0] = sink;
    if (lane == 0) {
        unsigned long long *o = out + ((size_t)blockIdx.x * 18 + warp) * 4;
        o[0] = mma_cycles; o[1] = n_mma; o[2] = ld_cycles; o[3] = n_ld;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int blocks = p.multiProcessorCount & ~1;
    unsigned long long *d, *h = (unsigned long long *)malloc((size_t)blocks * 18 * 4 * 8);
    unsigned int *stop;
    CK(cudaMalloc(&d, (size_t)blocks * 18 * 4 * 8));
    CK(cudaMalloc(&stop, blocks * 4));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem) + 1024));
    printf("cta_group::2 SS, M = 256 per pair\n%4s %8s | %14s %22s %18s\n", "N", "ld warps", "cycles per MMA", "per 128x128x32 per SM", "ld B/clk/SM");
    for (int N = 128; N <= 256; N += 128)
        for (int lw = 0; lw <= 16; lw = lw ? lw * 2 : 4) {
            CK(cudaMemset(d, 0, (size_t)blocks * 18 * 4 * 8));
            CK(cudaMemset(stop, 0, blocks * 4));
            probe<<<blocks, 576, sizeof(Smem) + 1024>>>(N, lw, 4000, stop, d);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, d, (size_t)blocks * 18 * 4 * 8, cudaMemcpyDeviceToHost));
            double mma_c = 0, mma_n = 0, ld_rate = 0;
            for (int b = 0; b < blocks; ++b) {
                mma_c += (double)h[(b * 18 + 0) * 4 + 0]; mma_n += (double)h[(b * 18 + 0) * 4 + 1];
                for (int w = 2; w < 18; ++w) {
                    const double c = (double)h[(b * 18 + w) * 4 + 2], n = (double)h[(b * 18 + w) * 4 + 3];
                    if (c > 0) ld_rate += n * 4096.0 / c;
                }
            }
            const double per = mma_n ? mma_c / mma_n : 0;   // each MMA = 256 x N x 32 over two SMs = N/128 units per SM
            printf("%4d %8d | %14.1f %22.1f %18.1f\n", N, lw, per, per * 128.0 / N, ld_rate / blocks);
        }
    return 0;
}
