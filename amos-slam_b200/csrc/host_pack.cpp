// host_pack.cpp -- host side of the batched MovingKeyPoints path: the dynamic-object masks only enter the pipeline as "mask != 0"
// (src/ORBextractor.cc:1697-1724 tests closing != 0, and erode / dilate of a 0 / non-0 image depend on nothing else), so the host-pointer
// batch call packs them to 1 bit per pixel BEFORE the PCIe transfer: the end-to-end figure of the 1080p configuration is bound by the
// host -> device link, and a packed mask is 1/8 of the bytes (4.1 -> 2.3 MB per frame with its image).
// Plain C++ (g++ through nvcc): AVX2 where the CPU has it (one compare + movemask per 32 pixels), 64-bit SWAR otherwise.
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {
// 8 mask bytes -> 8 bits (bit k = byte k non-zero)
inline uint32_t nz8(uint64_t x) {
    const uint64_t t = (((x & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full) | x) & 0x8080808080808080ull;   // 0x80 per non-zero byte
    return (uint32_t)(((t >> 7) * 0x0102040810204080ull) >> 56);
}
void pack_row_swar(const uint8_t* row, int cols, uint32_t* out, int wpr) {
    for (int w = 0; w < wpr; ++w) {
        const int x0 = 32 * w, n = cols - x0 < 32 ? cols - x0 : 32;
        uint32_t v = 0;
        int k = 0;
        for (; k + 8 <= n; k += 8) { uint64_t q; memcpy(&q, row + x0 + k, 8); v |= nz8(q) << k; }
        for (; k < n; ++k) v |= (uint32_t)(row[x0 + k] != 0) << k;
        out[w] = v;
    }
}
#if defined(__x86_64__)
__attribute__((target("avx2"))) void pack_row_avx2(const uint8_t* row, int cols, uint32_t* out, int wpr) {
    const __m256i z = _mm256_setzero_si256();
    const int full = cols / 32;
    for (int w = 0; w < full; ++w) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(row + 32 * w));
        out[w] = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, z));
    }
    if (full < wpr) pack_row_swar(row + 32 * full, cols - 32 * full, out + full, wpr - full);
}
#endif
}

// bits[y * wpr + w] bit k = mask[y][32 w + k] != 0, bits of columns >= cols are 0 (the layout of k_mask_pack, k_cull.cuh)
extern "C" void orbx_host_pack_mask(const uint8_t* mask, size_t step, int rows, int cols, uint32_t* bits) {
    const int wpr = (cols + 31) / 32;
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) { for (int y = 0; y < rows; ++y) pack_row_avx2(mask + (size_t)y * step, cols, bits + (size_t)y * wpr, wpr); return; }
#endif
    for (int y = 0; y < rows; ++y) pack_row_swar(mask + (size_t)y * step, cols, bits + (size_t)y * wpr, wpr);
}
