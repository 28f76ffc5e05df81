// k_pyramid_fast.cuh -- image pyramid (bilinear resize, fixed point).  The FAST stage lives in k_fast.cuh.
#pragma once
#include "orbx_common.cuh"
#include <cuda_pipeline.h>
#include "tma.cuh"
#include <cooperative_groups.h>

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


// =================================================================================================
// K1t  pyr_resize, TMA-staged tile form (the one the batch and per-frame paths run; k_pyr_resize_w / k_pyr_resize above remain
// for scale factors whose source tile does not fit a TMA box).  One warp = one tile of 128 destination columns x RESIZE_ROWS
// destination rows; the source rectangle that tile reads (about 1.2x as wide and high at ORB's scale factor, widened to the
// left to a 16-byte boundary) is fetched by ONE bulk tensor copy and awaited on an mbarrier.  Everything a lane needs about its 4
// destination columns (first source column, PRMT selectors, horizontal weights) is loaded ONCE per tile instead of once per
// row, and the horizontal pass of a source row is computed once and reused by the (up to two) destination rows that read it --
// the previous per-thread form spent 155 instructions per 4 pixels, 110 of them on tables and addresses (profiles/r02b).
// =================================================================================================
#define RESIZE_ROWS 32
#define RESIZE_WARPS 4

__global__ void __launch_bounds__(RESIZE_WARPS * 32)
k_pyr_resize_t(const __grid_constant__ CUtensorMap map_l0, const CUtensorMap* __restrict__ gmap, int b0, int sh,
               uint8_t* __restrict__ dst, long long dst_fstride, int dpitch, int dw, int dh, ResizeTabs t, int bw, int bh) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Y0 = (blockIdx.y * RESIZE_WARPS + warp) * RESIZE_ROWS;
    if (Y0 >= dh) return;
    const int per_warp = (bw * bh + 16 + 127) & ~127;
    uint8_t* sm = smem_raw + (size_t)warp * per_warp;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + bw * bh);
    const int gx0 = blockIdx.x * 32;                                  // first group of 4 destination columns of the tile
    const int bx0 = __ldg(&t.xg[gx0].x) & ~15;                        // box origin: 16-byte boundary left of the tile's first source column
    const int by0 = __ldg(t.yofs + Y0);
    if (lane == 0) {
        mbar_init(bar, 1); mbar_fence_init();
        mbar_expect_tx(bar, (uint32_t)(bw * bh));
        tma_load_3d(sm, gmap ? gmap : &map_l0, bx0, by0, b0 + (int)blockIdx.z, bar);
    }
    // per-lane column constants, loaded while the copy is in flight
    const int gx = gx0 + lane, x = gx * 4;
    const int2 e = __ldg(t.xg + gx);
    const uint4 wq = __ldg(reinterpret_cast<const uint4*>(t.xw + x));          // 4 x (w0 | w1 << 16)
    const int rel = e.x - bx0;
    const int wi = rel >> 2, sh8 = (rel & 3) * 8;
    const uint32_t s01 = (uint32_t)e.y & 0xFFFFu, s23 = (uint32_t)e.y >> 16;
    const int wpr = bw >> 2;
    const bool active = x < dpitch;                                   // tables are padded to the destination pitch
    const bool store = x < dw;
    __syncwarp();
    mbar_wait(bar, 0);
    const uint32_t* col = reinterpret_cast<const uint32_t*>(sm) + (active ? wi : 0);
    // horizontal pass of one staged source row: 4 sums, already shifted (H >> 4)
    auto hpass = [&](int row, uint32_t (&G)[4]) {
        const uint32_t* r32 = col + row * wpr;
        const uint32_t a0 = r32[0], a1 = r32[1], a2 = r32[2];
        const uint32_t W0 = __funnelshift_r(a0, a1, sh8), W1 = __funnelshift_r(a1, a2, sh8);
        const uint32_t X01 = __byte_perm(W0, W1, s01), X23 = __byte_perm(W0, W1, s23);     // (l0,r0,l1,r1), (l2,r2,l3,r3)
        G[0] = __dp2a_lo(wq.x, X01, 0u) >> 4; G[1] = __dp2a_hi(wq.y, X01, 0u) >> 4;
        G[2] = __dp2a_lo(wq.z, X23, 0u) >> 4; G[3] = __dp2a_hi(wq.w, X23, 0u) >> 4;
    };
    uint32_t G0[4], G1[4] = {0u, 0u, 0u, 0u};
    int held = -1;                                                    // source row whose pass is in G1
    uint8_t* drow = dst + (long long)(b0 + (int)blockIdx.z) * dst_fstride + (long long)Y0 * dpitch + x;
    const int Y1 = min(Y0 + RESIZE_ROWS, dh);
    for (int y = Y0; y < Y1; ++y) {
        const int sy0 = __ldg(t.yofs + y), sy1 = min(sy0 + 1, sh - 1);
        const short2 bwt = __ldg(t.yw + y);
        if (sy0 == held) { G0[0] = G1[0]; G0[1] = G1[1]; G0[2] = G1[2]; G0[3] = G1[3]; }
        else hpass(sy0 - by0, G0);
        if (sy1 == sy0) { G1[0] = G0[0]; G1[1] = G0[1]; G1[2] = G0[2]; G1[3] = G0[3]; }
        else hpass(sy1 - by0, G1);
        held = sy1;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            v[k] = (uint32_t)(((((int)bwt.x * (int)G0[k]) >> 16) + (((int)bwt.y * (int)G1[k]) >> 16) + 2) >> 2);
        const uint32_t out = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
        if (store) *reinterpret_cast<uint32_t*>(drow) = out;         // the (up to 3) bytes past dw land in row padding
        drow += dpitch;
    }
}

// =================================================================================================
// K1c  pyr_chain: all levels >= 1 of ONE frame in one launch, by one thread-block cluster of 8 CTAs (8192 threads).  The pyramid is a
// chain (level l is resized from level l-1, /root/reference/src/ORBextractor.cc:1848), so a single frame costs 7 dependent launches
// of a few microseconds each in the per-level forms -- launch-latency bound (31 us of the 118 us a 640 x 480 frame spends on the
// device).  Here the dependency between levels is a cluster barrier (barrier.cluster arrive.release / wait.acquire, hardware-
// supported on sm_90+) instead of a kernel boundary: the 8 CTAs of a cluster are co-scheduled on one GPC, write level l to global
// memory (L2), synchronise, and read it back as the source of level l + 1.  Used for a handful of frames (one cluster per frame);
// batches keep k_pyr_resize_t.  Arithmetic = k_pyr_resize_w's (word-load form), source reads bypass L1 (__ldcg): the lines were
// written by other SMs during this launch.
// =================================================================================================
#define CHAIN_CTAS 8
#define CHAIN_THREADS 1024
struct ChainLevel { int sw, sh, spitch, dw, dh, dpitch; long long soff, doff; ResizeTabs t; };   // offsets inside one frame's pyramid block (level 1 reads the level-0 view)

__global__ void __cluster_dims__(CHAIN_CTAS, 1, 1) __launch_bounds__(CHAIN_THREADS)
k_pyr_chain(PyrView pv, int b0, const ChainLevel* __restrict__ lv, int L) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int b = b0 + (int)blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int gwarp = (int)blockIdx.x * (CHAIN_THREADS / 32) + (threadIdx.x >> 5), nwarps = CHAIN_CTAS * (CHAIN_THREADS / 32);
    uint8_t* frame = pv.pyr + (long long)b * pv.pyr_fstride;
    for (int l = 1; l < L; ++l) {
        const ChainLevel c = lv[l];
        const uint8_t* s = l == 1 ? pv.l0 + (long long)b * pv.l0_fstride : frame + c.soff;
        const int spitch = l == 1 ? pv.l0_pitch : c.spitch;
        uint8_t* d = frame + c.doff;
        const int ngx = (c.dw + 3) >> 2, wc = (ngx + 31) >> 5, ntask = wc * c.dh;
        const int wmax = (spitch >> 2) - 1;
        for (int task = gwarp; task < ntask; task += nwarps) {
            const int y = task / wc, gx = (task - y * wc) * 32 + lane;
            if (gx < ngx) {
                const int x = gx * 4;
                const int sy0 = __ldg(c.t.yofs + y), sy1 = min(sy0 + 1, c.sh - 1);
                const short2 bw = __ldg(c.t.yw + y);
                const int2 e = __ldg(c.t.xg + gx);
                const uint4 wq = __ldg(reinterpret_cast<const uint4*>(c.t.xw + x));
                const int i0 = e.x >> 2, i1 = min(i0 + 1, wmax), i2 = min(i0 + 2, wmax);   // clamped words are never selected
                const int sh8 = (e.x & 3) * 8;
                const uint32_t s01 = (uint32_t)e.y & 0xFFFFu, s23 = (uint32_t)e.y >> 16;
                int h[2][4];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const uint32_t* row = reinterpret_cast<const uint32_t*>(s + (long long)(r ? sy1 : sy0) * spitch);
                    const uint32_t a0 = __ldcg(row + i0), a1 = __ldcg(row + i1), a2 = __ldcg(row + i2);
                    const uint32_t W0 = __funnelshift_r(a0, a1, sh8), W1 = __funnelshift_r(a1, a2, sh8);
                    const uint32_t X01 = __byte_perm(W0, W1, s01), X23 = __byte_perm(W0, W1, s23);
                    h[r][0] = (int)__dp2a_lo(wq.x, X01, 0u); h[r][1] = (int)__dp2a_hi(wq.y, X01, 0u);
                    h[r][2] = (int)__dp2a_lo(wq.z, X23, 0u); h[r][3] = (int)__dp2a_hi(wq.w, X23, 0u);
                }
                uint32_t v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    v[k] = (uint32_t)(((((int)bw.x * (h[0][k] >> 4)) >> 16) + (((int)bw.y * (h[1][k] >> 4)) >> 16) + 2) >> 2);
                *reinterpret_cast<uint32_t*>(d + (long long)y * c.dpitch + x) = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
            }
        }
        __threadfence();
        cluster.sync();                                              // level l is complete and visible before any CTA of the cluster reads it
    }
}

// =================================================================================================
// K1p  pyr_tiles: the WHOLE pyramid of a frame in one launch without any synchronisation between CTAs.  The chain (level l from level
// l-1) makes a single frame wait for 7 dependent launches (31 us of its 109 us on the device).  Here every CTA owns one rectangle of
// every level (an equal split of each level over the CTA grid) and computes, level by level in shared memory, that rectangle plus the
// few extra columns / rows of the level that its own rectangles further up need as sources -- recomputing the seam instead of waiting
// for a neighbour.  A pixel is the same deterministic function of the level below whoever computes it, so the result is bit-identical
// to the per-level kernels.  Level 0 is read from global memory once, every level is written once; only __syncthreads() between levels.
// The regions are rectangles (the footprints are separable) precomputed by the host (build_plan): per level and CTA column / row the
// needed range [n0, n1) and the owned range [o0, o1), in groups of 4 columns / in rows.  Arithmetic = k_pyr_resize_w's.
// =================================================================================================
#define TILEPYR_THREADS 512
struct TilePyrLevel { int sw, sh, dw, dh, dpitch, spitch, sm_off, sm_pitch, tab_off, tab_rows, tab_groups; long long doff; ResizeTabs t; };   // tab_off: staged tables [yofs | yw | xg | xw]
struct TilePyrPlan { const TilePyrLevel* lv; const int* xr; const int* yr; int L, nx, ny; };   // xr: [L][nx][4] = need0, need1, own0, own1 (groups); yr: [L][ny][4] (rows)

__global__ void __launch_bounds__(TILEPYR_THREADS)
k_pyr_tiles(PyrView pv, int b0, TilePyrPlan P) {
    extern __shared__ __align__(16) uint8_t tp_sm[];
    const int b = b0 + (int)blockIdx.z;
    const uint8_t* l0 = pv.l0 + (long long)b * pv.l0_fstride;
    uint8_t* frame = pv.pyr + (long long)b * pv.pyr_fstride;
    // stage every level's slice of the resize tables in shared memory first, one warp per level: the loads of all levels are in flight
    // together, so the chain below pays the global-memory latency once instead of once per level
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int l = 1 + warp; l < P.L; l += TILEPYR_THREADS / 32) {
            const TilePyrLevel c = P.lv[l];
            const int* xr = P.xr + ((size_t)l * P.nx + blockIdx.x) * 4;
            const int* yr = P.yr + ((size_t)l * P.ny + blockIdx.y) * 4;
            int* ty = reinterpret_cast<int*>(tp_sm + c.tab_off); short2* tw = reinterpret_cast<short2*>(ty + c.tab_rows);
            int2* tg = reinterpret_cast<int2*>(tw + c.tab_rows); uint4* tq = reinterpret_cast<uint4*>(tg + c.tab_groups);
            for (int t = lane; t < yr[1] - yr[0]; t += 32) { ty[t] = __ldg(c.t.yofs + yr[0] + t); tw[t] = __ldg(c.t.yw + yr[0] + t); }
            for (int t = lane; t < xr[1] - xr[0]; t += 32) { tg[t] = __ldg(c.t.xg + xr[0] + t); tq[t] = __ldg(reinterpret_cast<const uint4*>(c.t.xw) + xr[0] + t); }
        }
        __syncthreads();
    }
    for (int l = 1; l < P.L; ++l) {
        const TilePyrLevel c = P.lv[l];
        const int* xr = P.xr + ((size_t)l * P.nx + blockIdx.x) * 4;
        const int* yr = P.yr + ((size_t)l * P.ny + blockIdx.y) * 4;
        const int gx0 = xr[0], gx1 = xr[1], ox0 = xr[2], ox1 = xr[3], y0 = yr[0], y1 = yr[1], oy0 = yr[2], oy1 = yr[3];
        const int ngr = gx1 - gx0, n = ngr * (y1 - y0);
        // source: level 0 in global memory (l = 1) or the region of level l-1 this CTA computed into shared memory
        const int* sxr = P.xr + ((size_t)(l - 1) * P.nx + blockIdx.x) * 4;
        const int* syr = P.yr + ((size_t)(l - 1) * P.ny + blockIdx.y) * 4;
        const TilePyrLevel cs = P.lv[l - 1];
        const uint8_t* sbase = l == 1 ? l0 : tp_sm + cs.sm_off;
        const int spitch = l == 1 ? pv.l0_pitch : cs.sm_pitch;
        const int sgx0 = l == 1 ? 0 : sxr[0], sy_org = l == 1 ? 0 : syr[0];
        const int wlast = l == 1 ? (pv.l0_pitch >> 2) - 1 : (sxr[1] - sxr[0]) - 1;     // last source word that exists (clamped words are never selected)
        uint8_t* dsm = tp_sm + c.sm_off;
        uint8_t* dgl = frame + c.doff;
        const int* ty = reinterpret_cast<const int*>(tp_sm + c.tab_off); const short2* tw = reinterpret_cast<const short2*>(ty + c.tab_rows);
        const int2* tg = reinterpret_cast<const int2*>(tw + c.tab_rows); const uint4* tq = reinterpret_cast<const uint4*>(tg + c.tab_groups);
        for (int t = threadIdx.x; t < n; t += TILEPYR_THREADS) {
            const int ry = t / ngr, rg = t - ry * ngr;
            const int y = y0 + ry, gx = gx0 + rg, x = gx * 4;
            const int sy0 = ty[ry], sy1 = min(sy0 + 1, c.sh - 1);
            const short2 bw = tw[ry];
            const int2 e = tg[rg];
            const uint4 wq = tq[rg];
            const int i0 = (e.x >> 2) - sgx0, i1 = min(i0 + 1, wlast), i2 = min(i0 + 2, wlast);
            const int sh8 = (e.x & 3) * 8;
            const uint32_t s01 = (uint32_t)e.y & 0xFFFFu, s23 = (uint32_t)e.y >> 16;
            int h[2][4];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const uint32_t* row = reinterpret_cast<const uint32_t*>(sbase + (long long)((r ? sy1 : sy0) - sy_org) * spitch);
                const uint32_t a0 = row[i0], a1 = row[i1], a2 = row[i2];
                const uint32_t W0 = __funnelshift_r(a0, a1, sh8), W1 = __funnelshift_r(a1, a2, sh8);
                const uint32_t X01 = __byte_perm(W0, W1, s01), X23 = __byte_perm(W0, W1, s23);
                h[r][0] = (int)__dp2a_lo(wq.x, X01, 0u); h[r][1] = (int)__dp2a_hi(wq.y, X01, 0u);
                h[r][2] = (int)__dp2a_lo(wq.z, X23, 0u); h[r][3] = (int)__dp2a_hi(wq.w, X23, 0u);
            }
            uint32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                v[k] = (uint32_t)(((((int)bw.x * (h[0][k] >> 4)) >> 16) + (((int)bw.y * (h[1][k] >> 4)) >> 16) + 2) >> 2);
            const uint32_t out = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
            *reinterpret_cast<uint32_t*>(dsm + ry * c.sm_pitch + 4 * rg) = out;
            if (gx >= ox0 && gx < ox1 && y >= oy0 && y < oy1 && x < c.dw)
                *reinterpret_cast<uint32_t*>(dgl + (long long)y * c.dpitch + x) = out;       // bytes past dw land in row padding
        }
        __syncthreads();
    }
}
