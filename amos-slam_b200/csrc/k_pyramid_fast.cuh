// k_pyramid_fast.cuh -- image pyramid (bilinear, fixed point) and per-cell FAST-9 detection kernels.
#pragma once
#include "orbx_common.cuh"
#include <cuda_pipeline.h>

// =================================================================================================
// K1  pyr_resize: level l from level l-1 (chained, /root/reference/src/ORBextractor.cc:1826-1886),
// cv::resize(INTER_LINEAR) 8UC1 fixed-point arithmetic:
//   H[d]  = src[y][sx]*w0 + src[y][min(sx+1,sw-1)]*w1          (11-bit weights, int)
//   out   = (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2
// Tables (offset + weight pairs per destination column / row) are built on the host with OpenCV's
// float arithmetic and are padded to the destination pitch, so the kernel needs no edge branches.
// One thread produces 4 horizontally adjacent pixels and stores them as one uchar4; a CTA of 32x8
// threads covers a 128x8 destination tile; grid.z walks the batch.  HBM-bound stencil: reads
// P(l-1) once (L1/L2 absorb the 2x2 footprint overlap), writes P(l) once.
// =================================================================================================
// Two forms.  k_pyr_resize_w (scale factors <= 2, i.e. every real ORB-SLAM configuration): the 4 source
// pixel pairs of a thread lie within 8 bytes of its first source pixel, so each source row is read as 3
// aligned 32-bit words, shifted into place with two funnel shifts, and the (left,right) byte pairs are
// picked with two PRMTs whose selectors come from the host table; the horizontal pass is then one
// 2-way integer dot product (IDP.2A: u16 weights x u8 pixels) per pixel.  k_pyr_resize is the
// byte-gather form for arbitrary scale factors.
struct ResizeTabs {
    const int* xofs; const short2* xw; const int* yofs; const short2* yw;
    const int2* xg;       // per group of 4 destination columns: {first source column, PRMT selectors (pair 0,1 | pair 2,3 << 16)}
    int wide;             // 1 = the word-load kernel applies to this level
};

__global__ void __launch_bounds__(256)
k_pyr_resize_w(const uint8_t* __restrict__ src, long long src_fstride, int spitch, int sw, int sh,
               uint8_t* __restrict__ dst, long long dst_fstride, int dpitch, int dw, int dh, ResizeTabs t) {
    const int gx = blockIdx.x * 32 + threadIdx.x, x = gx * 4;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const uint8_t* s = src + (long long)blockIdx.z * src_fstride;
    const int sy0 = __ldg(t.yofs + y);
    const int sy1 = min(sy0 + 1, sh - 1);
    const short2 bw = __ldg(t.yw + y);
    const int2 e = __ldg(t.xg + gx);
    const uint4 wq = __ldg(reinterpret_cast<const uint4*>(t.xw + x));          // 4 x (w0 | w1 << 16)
    const int wmax = (spitch >> 2) - 1;
    const int i0 = e.x >> 2, i1 = min(i0 + 1, wmax), i2 = min(i0 + 2, wmax);   // clamped words are never selected
    const int sh8 = (e.x & 3) * 8;
    const uint32_t s01 = (uint32_t)e.y & 0xFFFFu, s23 = (uint32_t)e.y >> 16;
    int h[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const uint32_t* row = reinterpret_cast<const uint32_t*>(s + (long long)(r ? sy1 : sy0) * spitch);
        const uint32_t a0 = __ldg(row + i0), a1 = __ldg(row + i1), a2 = __ldg(row + i2);
        const uint32_t W0 = __funnelshift_r(a0, a1, sh8), W1 = __funnelshift_r(a1, a2, sh8);
        const uint32_t X01 = __byte_perm(W0, W1, s01), X23 = __byte_perm(W0, W1, s23);     // (l0,r0,l1,r1), (l2,r2,l3,r3)
        h[r][0] = (int)__dp2a_lo(wq.x, X01, 0u); h[r][1] = (int)__dp2a_hi(wq.y, X01, 0u);
        h[r][2] = (int)__dp2a_lo(wq.z, X23, 0u); h[r][3] = (int)__dp2a_hi(wq.w, X23, 0u);
    }
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        v[k] = (uint32_t)(((((int)bw.x * (h[0][k] >> 4)) >> 16) + (((int)bw.y * (h[1][k] >> 4)) >> 16) + 2) >> 2);
    const uint32_t out = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
    // pitch is a multiple of 128 and tables are padded: the (up to 3) bytes past dw land in row padding
    *reinterpret_cast<uint32_t*>(dst + (long long)blockIdx.z * dst_fstride + (long long)y * dpitch + x) = out;
}

__global__ void __launch_bounds__(256)
k_pyr_resize(const uint8_t* __restrict__ src, long long src_fstride, int spitch, int sw, int sh,
             uint8_t* __restrict__ dst, long long dst_fstride, int dpitch, int dw, int dh, ResizeTabs t) {
    const int x = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const uint8_t* s = src + (long long)blockIdx.z * src_fstride;
    const int sy0 = __ldg(t.yofs + y);
    const int sy1 = min(sy0 + 1, sh - 1);
    const short2 b = __ldg(t.yw + y);
    const uint8_t* r0 = s + (long long)sy0 * spitch;
    const uint8_t* r1 = s + (long long)sy1 * spitch;
    const int4 xo = __ldg(reinterpret_cast<const int4*>(t.xofs + x));
    const int xs[4] = {xo.x, xo.y, xo.z, xo.w};
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int sx0 = xs[k], sx1 = min(sx0 + 1, sw - 1);
        const short2 a = __ldg(t.xw + x + k);
        const int h0 = (int)__ldg(r0 + sx0) * a.x + (int)__ldg(r0 + sx1) * a.y;
        const int h1 = (int)__ldg(r1 + sx0) * a.x + (int)__ldg(r1 + sx1) * a.y;
        const int v = ((((int)b.x * (h0 >> 4)) >> 16) + (((int)b.y * (h1 >> 4)) >> 16) + 2) >> 2;
        out |= (uint32_t)(v & 0xFF) << (8 * k);
    }
    // pitch is a multiple of 128 and tables are padded: the (up to 3) bytes past dw land in row padding
    *reinterpret_cast<uint32_t*>(dst + (long long)blockIdx.z * dst_fstride + (long long)y * dpitch + x) = out;
}

