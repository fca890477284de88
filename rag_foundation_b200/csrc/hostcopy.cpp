// Host-side staging copy (pageable document -> pinned staging buffer) with non-temporal stores.
//
// The destination is read next by the GPU's DMA engine, never by this CPU: writing it with streaming stores
// skips the read-for-ownership of every destination line (a plain memcpy of a 2 MB chunk stays below glibc's
// non-temporal threshold and moves three bytes of memory traffic per byte copied instead of two).  Runtime
// dispatch on AVX2; anything else falls back to memcpy.
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>

__attribute__((target("avx2"))) static void stream_copy_avx2(uint8_t *dst, const uint8_t *src, size_t n) {
    // head: bring dst to 32-byte alignment
    const size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
    if (head) {
        const size_t h = head < n ? head : n;
        memcpy(dst, src, h);
        dst += h; src += h; n -= h;
    }
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 96), d);
    }
    _mm_sfence();
    if (i < n) memcpy(dst + i, src + i, n - i);
}
#endif

namespace rf {

void stage_copy(void *dst, const void *src, size_t n) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && n >= 4096) {
        stream_copy_avx2(static_cast<uint8_t *>(dst), static_cast<const uint8_t *>(src), n);
        return;
    }
#endif
    memcpy(dst, src, n);
}

}  // namespace rf
