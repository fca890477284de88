// Probe: how fast do tcgen05.mma kind::i8 (K = 32) and tcgen05.ld run when they share an SM?
// One MMA-issuing thread streams MMAs into rotating accumulators (SS: A and B from shared memory;
// TS: A from TMEM) while 0 / 4 / 8 / 16 other warps loop tcgen05.ld.32x32b.x32 over accumulator
// columns.  Rates only: operand data is whatever shared memory holds.  Answers, for the batched
// scoring kernel (score_topk_gemm.cu): is the accumulator read-back or the operand fetch the
// resource the MMA and the epilogue fight over?  Development tool.
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
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t ta, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n}"
                 ::"r"(d), "r"(ta), "l"(db), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Smem {
    alignas(1024) uint8_t a[4][128 * 128];    // 4 "M-tiles" x one 128-byte K-block (4 K-steps of 32)
    alignas(1024) uint8_t b[2][256 * 128];    // two stages, up to N = 256 rows
    alignas(8) uint64_t bar[4];
    uint32_t tmem_base;
    volatile uint32_t stop;
};

// mode: 0 no MMA, 1 SS, 2 TS.   N: 128 or 256.
// interleave: 0 = the 8 K-steps of one accumulator back to back (a dependent chain), 1 = K-step outer,
// accumulator inner (consecutive MMAs never touch the same accumulator)
__global__ void __launch_bounds__(576) probe(int mode, int N, int ld_warps, int n_batches, int interleave, unsigned long long *out) {
    extern __shared__ __align__(1024) uint8_t raw[];
    Smem &s = *reinterpret_cast<Smem *>(raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&s.bar[i], 1);
        s.stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < (int)(sizeof(s.a) + sizeof(s.b)) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(s.a)[i] = 0x01010101u * (i & 3);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem_base;
    // accumulators: SS -> 512 / N of them; TS -> A occupies the last 256 columns (4 M-tiles x 64), D the first 256
    const int n_acc = mode == 2 ? 256 / N : 512 / N;
    unsigned long long mma_cycles = 0, n_mma = 0;
    if (warp == 0 && lane == 0 && mode != 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const long long c0 = clock64();
        if (interleave) {
            // a round = n_acc accumulators x 8 K-steps, K-step outer; one commit per round
            const int rounds = n_batches / n_acc;
            for (int r = 0; r < rounds; ++r) {
                if (r >= 2) mbar_wait(&s.bar[r & 1], ((r >> 1) - 1) & 1);       // at most 2 rounds in flight
                for (int k = 0; k < 8; ++k) {
                    const uint64_t db = make_desc_sw128(s.b[k >> 2]) + (uint64_t)((k & 3) * 2);
                    for (int m = 0; m < n_acc; ++m) {
                        const uint32_t d = tmem + (uint32_t)(m * N);
                        if (mode == 1) umma_ss(d, make_desc_sw128(s.a[m]) + (uint64_t)((k & 3) * 2), db, idesc, k ? 1u : 0u);
                        else umma_ts(d, tmem + 256u + (uint32_t)(m * 64 + k * 8), db, idesc, k ? 1u : 0u);
                    }
                }
                commit(&s.bar[r & 1]);
            }
            for (int r = rounds - 2 > 0 ? rounds - 2 : 0; r < rounds; ++r) mbar_wait(&s.bar[r & 1], (r >> 1) & 1);
            mma_cycles = clock64() - c0;
            n_mma = (unsigned long long)rounds * n_acc * 8;
            s.stop = 1;
        } else {
        for (int bt = 0; bt < n_batches; ++bt) {           // a batch = 8 MMAs (K = 256) into one accumulator
            if (bt >= 4) mbar_wait(&s.bar[bt & 3], ((bt >> 2) - 1) & 1);   // at most 4 batches in flight
            const uint32_t d = tmem + (uint32_t)((bt % n_acc) * N);
            const int m = bt & 3;
            for (int k = 0; k < 8; ++k) {
                const uint64_t db = make_desc_sw128(s.b[k >> 2]) + (uint64_t)((k & 3) * 2);
                if (mode == 1) umma_ss(d, make_desc_sw128(s.a[m]) + (uint64_t)((k & 3) * 2), db, idesc, k ? 1u : 0u);
                else umma_ts(d, tmem + 256u + (uint32_t)(m * 64 + k * 8), db, idesc, k ? 1u : 0u);
            }
            commit(&s.bar[bt & 3]);
        }
        for (int bt = n_batches - 4 > 0 ? n_batches - 4 : 0; bt < n_batches; ++bt) mbar_wait(&s.bar[bt & 3], (bt >> 2) & 1);
        mma_cycles = clock64() - c0;
        n_mma = (unsigned long long)n_batches * 8;
        s.stop = 1;
        }
    }
    unsigned long long ld_cycles = 0, n_ld = 0;
    uint32_t sink = 0;
    if (warp >= 2 && warp - 2 < ld_warps) {
        const int w = warp - 2;
        const uint32_t t = tmem + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)((w >> 2) * 128 % (mode == 2 ? 256 : 512));
        const long long c0 = clock64();
        int i = 0;
        const int fixed = mode == 0 ? n_batches * 8 : 1 << 30;
        while (i < fixed && !s.stop) {
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
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int blocks = p.multiProcessorCount;
    unsigned long long *d, *h = (unsigned long long *)malloc((size_t)blocks * 18 * 4 * 8);
    CK(cudaMalloc(&d, (size_t)blocks * 18 * 4 * 8));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem) + 1024));
    const char *names[3] = {"none", "SS", "TS"};
    printf("%-4s %4s %5s %8s | %14s %16s %18s\n", "mma", "N", "order", "ld warps", "cycles per MMA", "per 128x128x32", "ld B/clk/SM");
    for (int mode = 0; mode < 3; ++mode)
      for (int il = 0; il < 2; ++il)
        for (int N = 128; N <= 256; N += 128) {
            if (mode == 0 && il) continue;
            if (il && (mode == 2 ? 256 / N : 512 / N) < 2) continue;
            if (mode == 0 && N == 256) continue;
            for (int lw = 0; lw <= 16; lw = lw ? lw * 2 : 4) {
                if (mode == 0 && lw == 0) continue;
                CK(cudaMemset(d, 0, (size_t)blocks * 18 * 4 * 8));
                probe<<<blocks, 576, sizeof(Smem) + 1024>>>(mode, N, lw, 4000, il, d);
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
                const double per = mma_n ? mma_c / mma_n : 0;
                printf("%-4s %4d %5s %8d | %14.1f %16.1f %18.1f\n", names[mode], N, il ? "k,m" : "m,k", lw, per, per * 128.0 / N, ld_rate / blocks);
            }
        }
    return 0;
}
