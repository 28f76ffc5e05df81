// k_pyramid_fast.cuh -- image pyramid (bilinear, fixed point) and per-cell FAST-9 detection kernels.
#pragma once
#include "orbx_common.cuh"

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
struct ResizeTabs { const int* xofs; const short2* xw; const int* yofs; const short2* yw; };

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
// One warp per cell, 4 cells per CTA.  The cell's ROI (zone + 3-px ring) is staged in shared memory
// with aligned 32-bit loads; S is computed only where two adjacent compass points pass the minTh test
// (any 9-arc contains two adjacent compass points) and a 9-run exists (bit tricks on 16-bit masks).
// Candidates are written to the cell's fixed slot range (capacity = max possible local maxima), so
// there are no atomics and the layout is deterministic; the octree stage gathers them in cell order,
// which reproduces vToDistributeKeys' order exactly.
// =================================================================================================
__device__ __forceinline__ int fast_sliding_min9_max(const int (&a)[16]) {
    // max over k of min(a[k..k+8]) on the circular ring (log-step sliding minimum)
    int m2[16], m4[16], m8[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m2[k] = min(a[k], a[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) m4[k] = min(m2[k], m2[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) m8[k] = min(m4[k], m4[(k + 4) & 15]);
    int best = -256;
#pragma unroll
    for (int k = 0; k < 16; ++k) best = max(best, min(m8[k], a[(k + 8) & 15]));
    return best;
}

__device__ __forceinline__ bool has_run9(uint32_t m16) {
    // 9 contiguous set bits on a 16-bit circular mask
    uint32_t m = m16 | (m16 << 16);
    uint32_t r2 = m & (m >> 1);
    uint32_t r4 = r2 & (r2 >> 2);
    uint32_t r8 = r4 & (r4 >> 4);
    uint32_t r9 = r8 & (m >> 8);
    return (r9 & 0xFFFFu) != 0;
}

// compass pre-test at minTh: any 9-arc of the ring contains two ADJACENT compass points (k = 0,4,8,12), so a pixel
// can only be a corner if two adjacent compass points are both brighter than v+t or both darker than v-t
__device__ __forceinline__ bool fast_compass(const uint8_t* p, int ps, int minTh) {
    const int v = p[0];
    const int d0 = (int)p[3 * ps] - v, d4 = (int)p[3] - v, d8 = (int)p[-3 * ps] - v, d12 = (int)p[-3] - v;
    const int hi = max(max(min(d0, d4), min(d4, d8)), max(min(d8, d12), min(d12, d0)));     // best adjacent "bright" pair
    const int lo = min(min(max(d0, d4), max(d4, d8)), min(max(d8, d12), max(d12, d0)));     // best adjacent "dark" pair
    return hi > minTh || lo < -minTh;
}

// FAST score of the pixel at p (shared-memory patch, row stride ps), 0 if S <= minTh
__device__ __forceinline__ int fast_score(const uint8_t* p, int ps, int minTh) {
    const int v = p[0];
    int d[16];   // v - p_k ; ring order (dx,dy): (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
    d[0] = v - p[3 * ps];        d[1] = v - p[3 * ps + 1];   d[2] = v - p[2 * ps + 2];   d[3] = v - p[ps + 3];
    d[4] = v - p[3];             d[5] = v - p[-ps + 3];      d[6] = v - p[-2 * ps + 2];  d[7] = v - p[-3 * ps + 1];
    d[8] = v - p[-3 * ps];       d[9] = v - p[-3 * ps - 1];  d[10] = v - p[-2 * ps - 2]; d[11] = v - p[-ps - 3];
    d[12] = v - p[-3];           d[13] = v - p[ps - 3];      d[14] = v - p[2 * ps - 2];  d[15] = v - p[3 * ps - 1];
    uint32_t mb = 0, md = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { mb |= (uint32_t)(d[k] < -minTh) << k; md |= (uint32_t)(d[k] > minTh) << k; }
    if (has_run9(md)) return fast_sliding_min9_max(d);          // dark arc: min(v - p_k)
    if (has_run9(mb)) {
#pragma unroll
        for (int k = 0; k < 16; ++k) d[k] = -d[k];
        return fast_sliding_min9_max(d);                        // bright arc: min(p_k - v)
    }
    return 0;
}

#define FAST_WARPS 4

__global__ void __launch_bounds__(FAST_WARPS * 32)
k_fast_cells(PyrView pv, const LevelGeom* __restrict__ levels, const CellDesc* __restrict__ cells, int ncells,
             int slots_per_frame, int smem_per_warp, int iniTh, int minTh,
             uint32_t* __restrict__ cand_slots,      // [B][slots_per_frame]  packed x:12|y:12|resp:8 (x,y relative to minBorder)
             uint16_t* __restrict__ cell_counts) {   // [B][ncells]
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cell = blockIdx.x * FAST_WARPS + warp;
    const int b = blockIdx.y;
    if (cell >= ncells) return;
    const CellDesc c = cells[cell];
    const LevelGeom& g = levels[c.level];
    int pitch;
    const uint8_t* img = level_ptr(pv, g, c.level, b, pitch);

    uint8_t* sm = smem_raw + (size_t)warp * smem_per_warp;
    const int zw = c.cw - 6, zh = c.ch - 6;
    const int xs = c.x0 & ~3, shift = c.x0 & 3;
    const int wpr = ((c.x0 + c.cw + 3) >> 2) - (xs >> 2);      // 32-bit words per patch row (<= 18)
    const int ps = wpr * 4;                                    // patch row stride (bytes)
    uint32_t* patch32 = reinterpret_cast<uint32_t*>(sm);
    const int patch_bytes = ps * c.ch;
    const int sst = zw + 2;                                    // S row stride; 1-px zero ring
    const int s_bytes = (sst * (zh + 2) + 3) & ~3;
    uint8_t* S = sm + patch_bytes;
    uint16_t* queue = reinterpret_cast<uint16_t*>(S + s_bytes);   // zw*zh entries: candidate pixels in raster order

    // stage the ROI with aligned 32-bit loads: each iteration covers 32/wpr rows
    {
        const int rpi = 32 / wpr, lr = lane / wpr, lc = lane - lr * wpr;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(img + (long long)c.y0 * pitch + xs) + lc;
        if (lr < rpi)
            for (int r = lr; r < c.ch; r += rpi) patch32[r * wpr + lc] = __ldg(src + (long long)r * (pitch >> 2));
    }
    for (int w = lane; w < (s_bytes >> 2); w += 32) reinterpret_cast<uint32_t*>(S)[w] = 0u;
    __syncwarp();

    // pass A: compass pre-test on every zone pixel, passing pixels compacted (in raster order) into the queue
    const uint8_t* patch = sm + shift;
    const uint32_t lt = (1u << lane) - 1u;
    int qn = 0;
    for (int y = 0; y < zh; ++y) {
        const uint8_t* row = patch + (y + 3) * ps + 3;
        for (int x0 = 0; x0 < zw; x0 += 32) {
            const int x = x0 + lane;
            const bool pass = x < zw && fast_compass(row + x, ps, minTh);
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            if (pass) queue[qn + __popc(m & lt)] = (uint16_t)((y << 6) | x);
            qn += __popc(m);
        }
    }
    __syncwarp();
    // pass B: full 16-pixel ring test + score, all lanes busy on compacted candidates
    for (int k = lane; k < qn; k += 32) {
        const int code = queue[k], y = code >> 6, x = code & 63;
        const int s = fast_score(patch + (y + 3) * ps + (x + 3), ps, minTh);
        if (s) S[(y + 1) * sst + x + 1] = (uint8_t)s; else queue[k] = 0xFFFFu;
    }
    __syncwarp();
    // pass C: strict 3x3 local maxima among the corners; count those above iniTh
    int n_ini = 0;
    for (int k0 = 0; k0 < qn; k0 += 32) {
        const int k = k0 + lane;
        int f = 0;
        if (k < qn) {
            const int code = queue[k];
            if (code != 0xFFFF) {
                const int y = code >> 6, x = code & 63;
                const uint8_t* q = S + (y + 1) * sst + x + 1;
                const int s = q[0];
                if (s > q[-1] && s > q[1] && s > q[-sst - 1] && s > q[-sst] && s > q[-sst + 1] &&
                    s > q[sst - 1] && s > q[sst] && s > q[sst + 1]) f = s;
                else queue[k] = 0xFFFFu;
            }
        }
        n_ini += __popc(__ballot_sync(0xffffffffu, f > iniTh));
    }
    __syncwarp();
    const int th = n_ini > 0 ? iniTh : minTh;                  // retry at minTh iff the cell came back empty
    uint32_t* out = cand_slots + (long long)b * slots_per_frame + c.slot;
    int n = 0;
    for (int k0 = 0; k0 < qn; k0 += 32) {
        const int k = k0 + lane;
        int f = 0, x = 0, y = 0;
        if (k < qn) {
            const int code = queue[k];
            if (code != 0xFFFF) { y = code >> 6; x = code & 63; f = S[(y + 1) * sst + x + 1]; }
        }
        const bool keep = f > th;
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (keep) out[n + __popc(m & lt)] = (uint32_t)(x + 3 + c.sx) | ((uint32_t)(y + 3 + c.sy) << 12) | ((uint32_t)(f - 1) << 24);
        n += __popc(m);
    }
    if (lane == 0) cell_counts[(long long)b * ncells + cell] = (uint16_t)n;
}
