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

// =================================================================================================
// K2  fast_cells: per-cell FAST-9/16 with the iniThFAST / minThFAST retry and 3x3 non-max suppression
// (ORBextractor.cc:1089-1157 + cv::FAST semantics), in score-map form:
//   S(p) = max over the 16 arcs of 9 contiguous ring pixels of min(v - p_k)  or  min(p_k - v);
//   corner at threshold t <=> S > t;  response = S - 1;
//   a cell's keypoints at threshold t = strict 3x3 local maxima of S inside the cell's zone with S > t
//   (neighbours outside the zone count as 0), in raster order; if the cell yields none at iniTh, the
//   same set at minTh is used.
// One warp per cell, 4 cells per CTA; the cell's ROI (zone + 3-px ring) is staged in shared memory with
// aligned 32-bit loads.  The kernel is instruction-issue bound, so the work is arranged as a funnel of
// warp-compacted queues in which every stage runs with all lanes busy on survivors of the previous one:
//   A  4 pixels per lane per step (one 32-bit word): polarity-free compass pre-test with byte-SIMD
//      (VABSDIFF4 + SWAR): any 9-arc contains ring point 0 or 8 and ring point 4 or 12, so a corner needs
//      (|N-v| > t or |S-v| > t) and (|E-v| > t or |W-v| > t).  ~23 % of the pixels survive.
//   B  exact 16-point test on the survivors: bright / dark ring masks built with one funnel shift per
//      ring pixel, 9-run detection with shift-and.  ~7 % of the pixels are corners at minTh.
//   C  score S of the corners only (3-input min/max network), written to a zero-initialised u8 map.
//   D  strict 3x3 local maxima among the corners, count of those above iniTh, ordered emission.
// Queues are filled by prefix sums over lanes (raster order is preserved end to end); candidates go to
// the cell's fixed slot range (capacity = max possible local maxima): no atomics, deterministic layout;
// the octree stage gathers them in cell order, which reproduces vToDistributeKeys' order exactly.
// =================================================================================================
__device__ __forceinline__ bool has_run9(uint32_t m16) {
    // 9 contiguous set bits on a 16-bit circular mask
    uint32_t m = m16 | (m16 << 16);
    uint32_t r2 = m & (m >> 1);
    uint32_t r4 = r2 & (r2 >> 2);
    uint32_t r8 = r4 & (r4 >> 4);
    uint32_t r9 = r8 & (m >> 8);
    return (r9 & 0xFFFFu) != 0;
}

__device__ __forceinline__ int min3i(int a, int b, int c) { return min(min(a, b), c); }
__device__ __forceinline__ int max3i(int a, int b, int c) { return max(max(a, b), c); }

// ring order (dx,dy): (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
#define FAST_RING(p, ps, R) do { \
    const int _s2 = 2 * (ps), _s3 = 3 * (ps); \
    R[0] = (p)[_s3];        R[1] = (p)[_s3 + 1];   R[2] = (p)[_s2 + 2];    R[3] = (p)[(ps) + 3]; \
    R[4] = (p)[3];          R[5] = (p)[3 - (ps)];  R[6] = (p)[2 - _s2];    R[7] = (p)[1 - _s3]; \
    R[8] = (p)[-_s3];       R[9] = (p)[-_s3 - 1];  R[10] = (p)[-_s2 - 2];  R[11] = (p)[-(ps) - 3]; \
    R[12] = (p)[-3];        R[13] = (p)[(ps) - 3]; R[14] = (p)[_s2 - 2];   R[15] = (p)[_s3 - 1]; } while (0)

// byte-wise "non-zero" -> 0x80 flag per byte
__device__ __forceinline__ uint32_t swar_nz(uint32_t x) { return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; }

#define FAST_WARPS 4

// Each warp walks cells cell0, cell0 + W, cell0 + 2W, ... of its frame (W = warps per frame) and double-buffers the ROI:
// the LDGSTS copies of the next cell are in flight while the current cell is processed, so the warp never waits for
// its patch except on the first cell.
__device__ __forceinline__ void fast_issue_patch(const PyrView& pv, const LevelGeom* __restrict__ levels, const CellDesc& c, int b, int lane, uint8_t* dst) {
    const LevelGeom& g = levels[c.level];
    int pitch;
    const uint8_t* img = level_ptr(pv, g, c.level, b, pitch);
    const int xs = c.x0 & ~3;
    const int wpr = ((c.x0 + c.cw + 3) >> 2) - (xs >> 2);      // 32-bit words per patch row
    const int rpi = 32 / wpr, lr = lane / wpr, lc = lane - lr * wpr;
    const uint8_t* src = img + (long long)c.y0 * pitch + xs + 4 * lc;
    if (lr < rpi)
        for (int r = lr; r < c.ch; r += rpi) __pipeline_memcpy_async(dst + 4 * (r * wpr + lc), src + r * pitch, 4);
}

__global__ void __launch_bounds__(FAST_WARPS * 32)
k_fast_cells(PyrView pv, const LevelGeom* __restrict__ levels, const CellDesc* __restrict__ cells, int ncells,
             int slots_per_frame, int smem_per_warp, int patch_cap, int s_cap, int iniTh, int minTh,
             uint32_t* __restrict__ cand_slots,      // [B][slots_per_frame]  packed x:12|y:12|resp:8 (x,y relative to minBorder)
             uint16_t* __restrict__ cell_counts) {   // [B][ncells]
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int W = gridDim.x * FAST_WARPS;
    int cell = blockIdx.x * FAST_WARPS + warp;
    if (cell >= ncells) return;
    uint8_t* smw = smem_raw + (size_t)warp * smem_per_warp;    // [patch 0 | patch 1 | S | queue]
    uint8_t* S = smw + 2 * patch_cap;
    uint16_t* queue = reinterpret_cast<uint16_t*>(S + s_cap);  // zw*zh entries: (dark<<15) | y<<6 | x, raster order
    const uint32_t lt = (1u << lane) - 1u;
    // stage-A masks for both thresholds: |d| > 2^k - 1 with the largest 2^k - 1 <= T (exact for minTh = 7, 15 for iniTh = 20)
    uint32_t keep_ini, keep_min;
    {
        int tq = 0;
        while (2 * tq + 1 <= iniTh) tq = 2 * tq + 1;
        keep_ini = (uint32_t)(0xFF & ~tq) * 0x01010101u;
        tq = 0;
        while (2 * tq + 1 <= minTh) tq = 2 * tq + 1;
        keep_min = (uint32_t)(0xFF & ~tq) * 0x01010101u;
    }

    CellDesc c = cells[cell];
    fast_issue_patch(pv, levels, c, b, lane, smw);
    __pipeline_commit();
    for (int buf = 0; cell < ncells; cell += W, buf ^= 1) {
    const int next = cell + W;
    CellDesc cnext = c;
    if (next < ncells) { cnext = cells[next]; fast_issue_patch(pv, levels, cnext, b, lane, smw + (buf ^ 1) * patch_cap); }
    __pipeline_commit();

    uint8_t* sm = smw + buf * patch_cap;
    const int zw = c.cw - 6, zh = c.ch - 6;
    const int xs = c.x0 & ~3, shift = c.x0 & 3;
    const int wpr = ((c.x0 + c.cw + 3) >> 2) - (xs >> 2);      // 32-bit words per patch row
    const int ps = wpr * 4;                                    // patch row stride (bytes)
    uint32_t* patch32 = reinterpret_cast<uint32_t*>(sm);
    const int sst = zw + 2;                                    // S row stride; 1-px zero ring
    const int s_bytes = (sst * (zh + 2) + 3) & ~3;
    for (int w = lane; w < (s_bytes >> 2); w += 32) reinterpret_cast<uint32_t*>(S)[w] = 0u;
    __pipeline_wait_prior(1);                                  // this cell's patch has landed (the next one may still be in flight)
    __syncwarp();

    // The reference runs FAST at iniThFAST and only re-runs a cell at minThFAST when that came back empty
    // (src/ORBextractor.cc:1126-1139).  Same here: the whole funnel runs at T = iniTh, and again at minTh only for cells
    // without a surviving corner (the scores already in S are a subset of the second pass's and identical, so S is kept).
    const uint8_t* patch = sm + shift;
    uint32_t* out = cand_slots + (long long)b * slots_per_frame + c.slot;
    int n = 0;
    for (int pass = 0; pass < 2; ++pass) {
    const int T = pass ? minTh : iniTh;
    const uint32_t keep = pass ? keep_min : keep_ini;
    // ---- stage A: polarity-free compass pre-test, 4 pixels (one word) per lane per step ----
    int qn = 0;
    {
        const int pc0 = shift + 3;                             // patch column of zone x = 0
        const int wi0 = pc0 >> 2, wi1 = (pc0 + zw - 1) >> 2, nw = wi1 - wi0 + 1;
        const uint32_t mfirst = 0xFFFFFFFFu << (8 * (pc0 & 3));
        const uint32_t mlast = 0xFFFFFFFFu >> (8 * (3 - ((pc0 + zw - 1) & 3)));
        const int ntask = zh * nw;
        // task = (row, word); a lane handles tasks t0+lane and t0+32+lane per step (8 pixels), so one packed prefix sum
        // orders 256 pixels in raster order
        int ya = lane / nw, wa = lane - ya * nw;
        const int dy32 = 32 / nw, dw32 = 32 - dy32 * nw;
        int yb = ya + dy32, wb = wa + dw32;
        if (wb >= nw) { wb -= nw; ++yb; }
        const int dy64 = 64 / nw, dw64 = 64 - dy64 * nw;
        auto compass = [&](int y, int w) -> uint32_t {
            const uint32_t* row = patch32 + (y + 3) * wpr + wi0 + w;
            const uint32_t C = row[0], Cl = row[-1], Cr = row[1];
            const uint32_t Nn = row[-3 * wpr], Ss = row[3 * wpr];
            const uint32_t E = __byte_perm(C, Cr, 0x6543), Wn = __byte_perm(Cl, C, 0x4321);
            const uint32_t u = (__vabsdiffu4(C, Nn) | __vabsdiffu4(C, Ss)) & keep;
            const uint32_t v = (__vabsdiffu4(C, E) | __vabsdiffu4(C, Wn)) & keep;
            uint32_t f = swar_nz(u) & swar_nz(v);
            if (w == 0) f &= mfirst;
            if (w == nw - 1) f &= mlast;
            return f;
        };
        auto emit = [&](uint32_t f, int pos, int y, int w) {
            const int code = (y << 6) + (4 * (wi0 + w) - pc0);              // + not |: the first word may start left of the zone (negative x of byte 0)
            if (f & 0x00000080u) queue[pos++] = (uint16_t)code;
            if (f & 0x00008000u) queue[pos++] = (uint16_t)(code + 1);
            if (f & 0x00800000u) queue[pos++] = (uint16_t)(code + 2);
            if (f & 0x80000000u) queue[pos] = (uint16_t)(code + 3);
        };
        for (int t0 = 0; t0 < ntask; t0 += 64) {
            const uint32_t fa = (t0 + lane < ntask) ? compass(ya, wa) : 0u;
            const uint32_t fb = (t0 + 32 + lane < ntask) ? compass(yb, wb) : 0u;
            const int ca = __popc(fa), cb = __popc(fb);
            int incl = ca | (cb << 16);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
            const int tot = __shfl_sync(0xffffffffu, incl, 31);
            emit(fa, qn + (incl & 0xFFFF) - ca, ya, wa);
            emit(fb, qn + (tot & 0xFFFF) + (incl >> 16) - cb, yb, wb);
            qn += (tot & 0xFFFF) + (tot >> 16);
            ya += dy64; wa += dw64; if (wa >= nw) { wa -= nw; ++ya; }
            yb += dy64; wb += dw64; if (wb >= nw) { wb -= nw; ++yb; }
        }
    }
    __syncwarp();
    // ---- stage B: exact 16-point segment test on the survivors; corners compacted in place ----
    int cn = 0;
    for (int k0 = 0; k0 < qn; k0 += 32) {
        const int k = k0 + lane;
        int code = 0, corner = 0;
        if (k < qn) {
            code = queue[k];
            const int y = code >> 6, x = code & 63;
            const uint8_t* p = patch + (y + 3) * ps + (x + 3);
            const int hi = (int)p[0] + T, lo = (int)p[0] - T;
            int R[16];
            FAST_RING(p, ps, R);
            uint32_t mb = 0, md = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                mb = __funnelshift_l((uint32_t)(hi - R[i]), mb, 1);          // sign bit <=> p_k > v + t
                md = __funnelshift_l((uint32_t)(R[i] - lo), md, 1);          // sign bit <=> p_k < v - t
            }
            const bool dark = has_run9(md & 0xFFFFu);
            corner = dark || has_run9(mb & 0xFFFFu);
            code |= dark ? 0x8000 : 0;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, corner);
        __syncwarp();
        if (corner) queue[cn + __popc(m & lt)] = (uint16_t)code;
        cn += __popc(m);
    }
    __syncwarp();
    // ---- stage C: score of the corners ----
    for (int k = lane; k < cn; k += 32) {
        const int code = queue[k], y = (code >> 6) & 63, x = code & 63;
        const uint8_t* p = patch + (y + 3) * ps + (x + 3);
        const int v = p[0];
        int R[16];
        FAST_RING(p, ps, R);
        int d[16];
        if (code & 0x8000) {
#pragma unroll
            for (int i = 0; i < 16; ++i) d[i] = v - R[i];
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) d[i] = R[i] - v;
        }
        int m3[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) m3[i] = min3i(d[i], d[(i + 1) & 15], d[(i + 2) & 15]);
        int best = -256;
#pragma unroll
        for (int i = 0; i < 16; i += 2)
            best = max3i(best, min3i(m3[i], m3[(i + 3) & 15], m3[(i + 6) & 15]), min3i(m3[(i + 1) & 15], m3[(i + 4) & 15], m3[(i + 7) & 15]));
        S[(y + 1) * sst + x + 1] = (uint8_t)best;
    }
    __syncwarp();
    // ---- stage D: strict 3x3 local maxima among the corners (all score > T), emitted in raster order ----
    for (int k0 = 0; k0 < cn; k0 += 32) {
        const int k = k0 + lane;
        int f = 0, x = 0, y = 0;
        if (k < cn) {
            const int code = queue[k];
            y = (code >> 6) & 63; x = code & 63;
            const uint8_t* q = S + (y + 1) * sst + x + 1;
            const int s = q[0];
            if (s > q[-1] && s > q[1] && s > q[-sst - 1] && s > q[-sst] && s > q[-sst + 1] &&
                s > q[sst - 1] && s > q[sst] && s > q[sst + 1]) f = s;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, f != 0);
        if (f) out[n + __popc(m & lt)] = (uint32_t)(x + 3 + c.sx) | ((uint32_t)(y + 3 + c.sy) << 12) | ((uint32_t)(f - 1) << 24);
        n += __popc(m);
    }
    if (n > 0 || minTh == iniTh) break;
    __syncwarp();                                              // queue is rebuilt by the second pass
    }
    if (lane == 0) cell_counts[(long long)b * ncells + cell] = (uint16_t)n;
    c = cnext;
    __syncwarp();                                              // S / queue are reused by the next cell
    }
}
