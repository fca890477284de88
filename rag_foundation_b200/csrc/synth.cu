// Synthetic RF-1 corpora generated directly in HBM (oracle/SPEC.md "Synthetic corpora",
// SURVEY.md §8d).  Counter-based: row c of corpus `seed` depends only on (seed, c), so a shard can
// be produced on any GPU and re-derived on the CPU by the oracle without shipping 26 GB around.
// One warp per row: <= 127 tokens hashed in four lane-strided rounds into a per-warp shared-memory
// histogram, then written as one dim-byte int8 row (8 bytes per lane and 256-byte sub-row) with its sum of squares.
#include "rf_device.cuh"
#include "rf_internal.h"

namespace rf {

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t a, uint64_t b) {
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + a * 0xBF58476D1CE4E5B9ull + b * 0x94D049BB133111EBull +
                 0x2545F4914F6CDD1Dull;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

constexpr int kSynthWarps = 8;

template <int kM>
__global__ void __launch_bounds__(kSynthWarps * 32) synth_rows_kernel(
    uint64_t seed, uint64_t start_counter, uint64_t n_rows, const uint16_t *__restrict__ zipf_bucket,
    int8_t *__restrict__ F, int32_t *__restrict__ ff, uint32_t *__restrict__ seg, uint32_t first_seg,
    uint64_t rows_per_store) {
    constexpr int kD = kSubDim * kM;
    __shared__ uint32_t hist[kSynthWarps][kD];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint32_t *h = hist[warp];
    const uint64_t warps_total = static_cast<uint64_t>(gridDim.x) * kSynthWarps;
    for (uint64_t r = static_cast<uint64_t>(blockIdx.x) * kSynthWarps + warp; r < n_rows; r += warps_total) {
#pragma unroll
        for (int j = 0; j < kD / 32; ++j) h[lane + 32 * j] = 0;
        __syncwarp();
        const uint64_t c = start_counter + r;
        const int len = 64 + static_cast<int>(mix64(seed ^ 0xA5ull, c, 0) & 63);
        for (int j = lane; j < len; j += 32) {
            const uint32_t b = zipf_bucket[mix64(seed, c, static_cast<uint64_t>(j)) >> 48];
            atomicAdd(&h[b], 1u);
        }
        __syncwarp();
        int sq = 0;
#pragma unroll
        for (int sub = 0; sub < kM; ++sub) {
            uint32_t packed[2];
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                uint32_t v = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t t = min(h[sub * kSubDim + lane * 8 + w * 4 + b], 127u);
                    v |= t << (8 * b);
                    sq += static_cast<int>(t * t);
                }
                packed[w] = v;
            }
            *reinterpret_cast<uint2 *>(F + (r * kM + sub) * kSubBytes + lane * 8) = make_uint2(packed[0], packed[1]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(kFull, sq, o);
        if (lane == 0) {
            ff[r] = sq;
            seg[r] = first_seg + (rows_per_store ? static_cast<uint32_t>(r / rows_per_store) : 0u);
        }
        __syncwarp();
    }
}

}  // namespace

cudaError_t launch_synth_rows(uint64_t seed, uint64_t start_counter, uint64_t n_rows, const uint16_t *zipf_bucket_dev, uint32_t dim,
                              int8_t *F, int32_t *ff, uint32_t *seg, uint32_t first_seg, uint64_t rows_per_store,
                              cudaStream_t s) {
    if (n_rows == 0) return cudaSuccess;
    uint64_t blocks = (n_rows + kSynthWarps - 1) / kSynthWarps;
    if (blocks > 148ull * 16ull) blocks = 148ull * 16ull;
    auto go = [&](auto kern) {
        kern<<<static_cast<unsigned>(blocks), kSynthWarps * 32, 0, s>>>(seed, start_counter, n_rows, zipf_bucket_dev, F, ff, seg, first_seg,
                                                                      rows_per_store);
    };
    switch (dim) {
        case 256: go(synth_rows_kernel<1>); break;
        case 512: go(synth_rows_kernel<2>); break;
        case 1024: go(synth_rows_kernel<4>); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace rf
