// score_topk_scan: query-vs-chunk scoring fused with per-store top-k selection (sm_100a).
//
// Carries out the retrieval step of GeminiRag.ask_stream (reference
// backend/app/services/gemini_rag.py:517-551; the mock's canned citation is :704-718) as the
// RF-1 spec, steps 6-7 (oracle/SPEC.md).  HBM-bound integer work: 260 algorithmic bytes per chunk
// (256 B int8 features + 4 B store-segment word), one pass, scores never written back.
//
// Layout.  A warp scores a tile of 32 consecutive rows (8 KB) per step: 16 independent 128-bit
// streaming loads per lane, issued back to back (512 contiguous bytes per warp-wide load, i.e.
// perfectly coalesced), 4 dp4a per load against the lane's 16-byte slice of the query held in
// registers, then a 15-shuffle transposing butterfly that leaves lane l with the finished int32
// score of row 2*(l&15) + (l>>4).  The lane checks its row's store-segment word against the
// query scope (tenant mask + tombstones), packs (score, id) into one u64 key and offers it to the
// warp's running top-k, which it enters only if it beats the warp's threshold (rare).  Warps merge
// through shared memory, blocks through a per-query partial buffer, and the last block to finish
// (ticket counter) merges the partials and writes ids / scores / cosines: one launch per search.
#include "rf_device.cuh"
#include "rf_internal.h"

namespace rf {

namespace {

constexpr int kScanWarps = 8;
constexpr int kScanThreads = kScanWarps * 32;
static_assert(kTileRows == static_cast<int>(kScanTileRows), "host and device disagree on the tile height");

// Reduce 16 per-lane partial sums (one per load) across the 16 lanes of each half-warp, leaving
// lane l with the total of partial index (l & 15).  8 + 4 + 2 + 1 shuffles.
__device__ __forceinline__ int transpose_reduce16(int (&p)[16], int lane) {
    int v8[8];
    {
        const bool up = lane & 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int send = up ? p[j] : p[j + 8];
            const int keep = up ? p[j + 8] : p[j];
            v8[j] = keep + __shfl_xor_sync(kFull, send, 8);
        }
    }
    int v4[4];
    {
        const bool up = lane & 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int send = up ? v8[j] : v8[j + 4];
            const int keep = up ? v8[j + 4] : v8[j];
            v4[j] = keep + __shfl_xor_sync(kFull, send, 4);
        }
    }
    int v2[2];
    {
        const bool up = lane & 2;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int send = up ? v4[j] : v4[j + 2];
            const int keep = up ? v4[j + 2] : v4[j];
            v2[j] = keep + __shfl_xor_sync(kFull, send, 2);
        }
    }
    const bool up = lane & 1;
    const int send = up ? v2[0] : v2[1];
    const int keep = up ? v2[1] : v2[0];
    return keep + __shfl_xor_sync(kFull, send, 1);
}

// Warps -> warp 0 through shared memory.  On return warp 0's `top` holds the block's top-k.
__device__ __forceinline__ void block_merge(WarpTopK &top, uint64_t (*s_keys)[32], int k, int warp, int lane) {
    __syncthreads();
    s_keys[warp][lane] = lane < k ? top.mine : 0ull;
    __syncthreads();
    if (warp == 0) {
        for (int w = 1; w < kScanWarps; ++w) top.consume(s_keys[w][lane], k, lane);
    }
}

__global__ void __launch_bounds__(kScanThreads, 2) score_topk_scan_kernel(const ScanArgs a) {
    __shared__ uint64_t s_keys[kScanWarps][32];
    __shared__ uint32_t s_scope[RF_SCOPE_MAX];
    __shared__ uint32_t s_is_last;

    const int qi = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int k = static_cast<int>(a.k);
    const ScanPlan &plan = a.plans[a.shared_plan ? 0 : qi];
    const uint32_t n_scope = plan.n_scope;
    if (threadIdx.x < RF_SCOPE_MAX) s_scope[threadIdx.x] = threadIdx.x < n_scope ? plan.scope[threadIdx.x] : kTombstone;
    __syncthreads();

    // this lane's 16-byte slice of the query
    const int4 qv = *reinterpret_cast<const int4 *>(a.q + static_cast<size_t>(qi) * kDim + (lane & 15) * 16);

    const uint32_t total_tiles = plan.total_tiles;
    const uint32_t t_lo = static_cast<uint32_t>(static_cast<uint64_t>(total_tiles) * blockIdx.x / gridDim.x);
    const uint32_t t_hi = static_cast<uint32_t>(static_cast<uint64_t>(total_tiles) * (blockIdx.x + 1) / gridDim.x);
    const uint32_t *ext_lo = a.ext_lo + plan.ext_off;
    const uint32_t *ext_hi = a.ext_hi + plan.ext_off;
    const uint32_t *ext_tile0 = a.ext_tile0 + plan.ext_off + (a.shared_plan ? 0 : qi);
    const int my_row_in_tile = 2 * (lane & 15) + (lane >> 4);

    WarpTopK top;
    top.reset();
    uint32_t e = 0;
    if (t_lo < t_hi && plan.n_ext > 1) {  // first extent holding tile t_lo (binary search)
        uint32_t lo = 0, hi = plan.n_ext;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (ext_tile0[mid] <= t_lo) lo = mid; else hi = mid;
        }
        e = lo;
    }

    for (uint32_t t = t_lo + warp; t < t_hi; t += kScanWarps) {
        while (t >= ext_tile0[e + 1]) ++e;
        const uint32_t row0 = ext_lo[e] + (t - ext_tile0[e]) * kTileRows;
        const uint32_t row_end = ext_hi[e];
        const int4 *src = reinterpret_cast<const int4 *>(a.F + static_cast<size_t>(row0) * kRowBytes) + lane;

        int4 x[16];
        uint32_t seg;
        const uint32_t my_row = row0 + my_row_in_tile;
        if (row0 + kTileRows <= row_end) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = ld_stream_v4(src + i * 32);
            seg = __ldg(a.seg + my_row);
        } else {  // ragged last tile of an extent
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint32_t r = row0 + 2 * i + (lane >> 4);
                x[i] = r < row_end ? ld_stream_v4(src + i * 32) : make_int4(0, 0, 0, 0);
            }
            seg = my_row < row_end ? __ldg(a.seg + my_row) : kTombstone;
        }

        int p[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            int acc = __dp4a(x[i].x, qv.x, 0);
            acc = __dp4a(x[i].y, qv.y, acc);
            acc = __dp4a(x[i].z, qv.z, acc);
            p[i] = __dp4a(x[i].w, qv.w, acc);
        }
        const int score = transpose_reduce16(p, lane);

        bool ok = false;
        if (seg != kTombstone) {
            for (uint32_t j = 0; j < n_scope; ++j) ok |= (seg == s_scope[j]);
        }
        const uint64_t key = ok ? pack_key(score, a.id_base + my_row) : 0ull;
        top.consume(key, k, lane);
    }

    // ---- block top-k -> partial buffer
    block_merge(top, s_keys, k, warp, lane);
    uint64_t *part = a.partial + (static_cast<size_t>(qi) * gridDim.x) * k;
    if (warp == 0) {
        if (lane < k) __stcg(part + static_cast<size_t>(blockIdx.x) * k + lane, top.mine);
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            const uint32_t ticket = atomicAdd(a.tickets + qi, 1u);
            s_is_last = (ticket == gridDim.x - 1) ? 1u : 0u;
        }
    }
    __syncthreads();
    if (!s_is_last) return;

    // ---- last block of this query: merge all partials, write the answer
    __threadfence();
    top.reset();
    const uint32_t n_part = gridDim.x * k;
    for (uint32_t base = 0; base < n_part; base += kScanThreads) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t key = i < n_part ? __ldcg(part + i) : 0ull;
        top.consume(key, k, lane);
    }
    block_merge(top, s_keys, k, warp, lane);
    if (warp != 0) return;

    // ||q||^2 for the reported cosine (lanes 0..15 cover the 256 query bytes once)
    int qq = __dp4a(qv.x, qv.x, 0);
    qq = __dp4a(qv.y, qv.y, qq);
    qq = __dp4a(qv.z, qv.z, qq);
    qq = __dp4a(qv.w, qv.w, qq);
    if (lane >= 16) qq = 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(kFull, qq, o);

    const uint64_t key = lane < k ? top.mine : 0ull;
    const unsigned found = __ballot_sync(kFull, key != 0ull);
    if (lane < k) {
        const size_t o = static_cast<size_t>(qi) * k + lane;
        a.out_keys[o] = key;
        const uint32_t gid = key_gid(key);
        const int32_t sc = key_score(key);
        if (a.out_ids) a.out_ids[o] = key ? static_cast<uint64_t>(gid) : ~0ull;
        if (a.out_scores) a.out_scores[o] = key ? sc : 0;
        if (a.out_cos) {
            float c = 0.0f;
            if (key && a.ff) {
                const float nq = sqrtf(static_cast<float>(qq));
                const float nf = sqrtf(static_cast<float>(__ldg(a.ff + (gid - a.id_base))));
                const float den = __fmul_rn(nq, nf);
                c = den == 0.0f ? 0.0f : __fdiv_rn(static_cast<float>(sc), den);
            }
            a.out_cos[o] = c;
        }
    }
    if (lane == 0) {
        if (a.out_counts) a.out_counts[qi] = __popc(found);
        a.tickets[qi] = 0;  // ready for the next launch on this context
    }
}

// k-way merge of n_lists top-k lists per query (after the all-gather of the sharded path).
__global__ void __launch_bounds__(128) merge_topk_kernel(const uint64_t *__restrict__ keys, uint32_t n_lists,
                                                         uint32_t nq, uint32_t k, uint64_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint32_t qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= nq) return;
    WarpTopK top;
    top.reset();
    for (uint32_t l = 0; l < n_lists; ++l) {
        const uint64_t key = lane < static_cast<int>(k) ? keys[(static_cast<size_t>(l) * nq + qi) * k + lane] : 0ull;
        top.consume(key, static_cast<int>(k), lane);
    }
    if (lane < static_cast<int>(k)) out[static_cast<size_t>(qi) * k + lane] = top.mine;
}

}  // namespace

uint32_t scan_default_blocks_per_query(int sm_count) { return static_cast<uint32_t>(sm_count) * 2u; }

cudaError_t launch_score_topk_scan(const ScanArgs &a, uint32_t nq, uint32_t blocks_per_query, cudaStream_t s) {
    dim3 grid(blocks_per_query, nq, 1);
    score_topk_scan_kernel<<<grid, kScanThreads, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_merge_topk(const uint64_t *keys, uint32_t n_lists, uint32_t nq, uint32_t k, uint64_t *out_keys,
                              cudaStream_t s) {
    const uint32_t warps_per_block = 4;
    merge_topk_kernel<<<(nq + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, s>>>(keys, n_lists, nq, k,
                                                                                                 out_keys);
    return cudaGetLastError();
}

}  // namespace rf
