// k_describe.cuh -- 7x7 Gaussian blur, IC_Angle orientation and rotated-BRIEF descriptors.
#pragma once
#include "orbx_common.cuh"
#include "det_math.cuh"
#include "tma.cuh"
#include <cuda_pipeline.h>

// =================================================================================================
// K6  gauss7: cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on every pyramid level
// (/root/reference/src/ORBextractor.cc:1626-1634, 1791-1793), OpenCV's 8-bit fixed-point path
// (SURVEY.md A.2):  q = [18,34,48,56,48,34,18]/256;  h = sum q*px (16-bit range, no rounding);
// v = sum q*h (32-bit);  out = (v + 32768) >> 16.  Reflection is at the LEVEL edge (the reference
// blurs a clone of the ROI).
// One warp = one tile of 128 columns x BLUR_STRIP rows of one level of one frame (the tile list covers all levels, so a single
// launch blurs the whole pyramid of the whole batch).  The tile with its 3-px halo -- 160 bytes x (BLUR_STRIP + 6) rows, starting
// 16 bytes left of the tile because TMA's innermost coordinate must be 16-byte aligned -- is fetched by ONE bulk tensor copy
// (cp.async.bulk.tensor.3d over (x, y, frame)) and awaited on an mbarrier; elements outside the level arrive as zeros and the
// BORDER_REFLECT_101 rows / columns are then written into the staged tile (edge tiles only).  After that the row loop has no
// address arithmetic, no edge cases and no shuffles: per input row a lane reads three consecutive words of the staged row
// (conflict-free), forms its 4 horizontal sums with 8 integer dot products (IDP.4A on PRMT-aligned byte windows), keeps the last
// 6 packed row pairs in registers and emits one output word (previous form: 114 instructions per word and row, this one ~45;
// profiles/r02b vs r02c).  HBM traffic = read level + write level (+ the halo rows, which L2 serves).
// =================================================================================================
#define BLUR_STRIP 42                    // throughput form: 48 staged rows per tile
#define BLUR_STRIP_SMALL 12              // latency form (a handful of frames): 18 staged rows per tile, 3.5x as many warps
#define BLUR_TILE_W 128
#define BLUR_BOX_W 160                   // 16 bytes left of the tile + tile + 16 bytes right: 40 words per staged row
#define BLUR_WARPS 4
// staged rows = strip + 6: a multiple of 6 (the row loop is unrolled over the 6 ring slots)
__host__ __device__ static inline int blur_smem_per_warp(int strip) { return (BLUR_BOX_W * (strip + 6) + 16 + 127) & ~127; }   // tile + mbarrier, 128-byte aligned (TMA destination)
struct BlurTile { short level, xc, strip, pad; };

__device__ __noinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;     // loops only on levels narrower than the 3-px halo
    return p;
}

__global__ void __launch_bounds__(BLUR_WARPS * 32)
k_gauss7(const __grid_constant__ CUtensorMap map_l0, const CUtensorMap* __restrict__ maps, int b0,
         const LevelGeom* __restrict__ levels, const BlurTile* __restrict__ tiles, int ntiles, int strip,
         uint8_t* __restrict__ blur, long long blur_fstride) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x * BLUR_WARPS + warp;
    if (tile >= ntiles) return;
    const int srows = strip + 6;
    uint8_t* sm = smem_raw + (size_t)warp * blur_smem_per_warp(strip);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + BLUR_BOX_W * srows);
    const BlurTile t = tiles[tile];
    const int b = blockIdx.y;
    const LevelGeom& g = levels[t.level];
    const int gw = g.w, gh = g.h, gpitch = g.pitch;
    const int x0 = t.xc * BLUR_TILE_W, bx0 = x0 - 16;               // tile / box origin (level columns)
    const int y0 = t.strip * strip, y1 = min(y0 + strip, gh);
    const int nrows = y1 - y0 + 6;                                   // staged rows that are read: level rows y0 - 3 .. y1 + 2
    if (lane == 0) {
        mbar_init(bar, 1); mbar_fence_init();
        mbar_expect_tx(bar, (uint32_t)(BLUR_BOX_W * srows));
        tma_load_3d(sm, t.level == 0 ? &map_l0 : maps + t.level, bx0, y0 - 3, b0 + b, bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    uint32_t* sm32 = reinterpret_cast<uint32_t*>(sm);
    // ---- BORDER_REFLECT_101 (warp-uniform branches; interior tiles skip both) ----
    if (y0 - 3 < 0 || y1 + 2 >= gh) {                               // rows above / below the level := their mirror rows (whole staged rows)
        for (int i = 0; i < nrows; ++i) {
            const int r = y0 - 3 + i;
            if (r >= 0 && r < gh) continue;
            const int si = reflect101(r, gh) - (y0 - 3);
            for (int w = lane; w < BLUR_BOX_W / 4; w += 32) sm32[i * (BLUR_BOX_W / 4) + w] = sm32[si * (BLUR_BOX_W / 4) + w];
        }
        __syncwarp();
    }
    if (x0 == 0 || x0 + BLUR_TILE_W + 3 > gw) {                     // columns left / right of the level := their mirror columns
        for (int k = lane; k < nrows * 6; k += 32) {
            const int i = k / 6, q = k - i * 6;
            const int c = q < 3 ? q - 3 : gw + q - 3;               // -3, -2, -1, gw, gw + 1, gw + 2
            if (c >= bx0 && c < bx0 + BLUR_BOX_W && c >= x0 - 3 && c < x0 + BLUR_TILE_W + 3)
                sm[i * BLUR_BOX_W + (c - bx0)] = sm[i * BLUR_BOX_W + (reflect101(c, gw) - bx0)];
        }
        __syncwarp();
    }
    // ---- row loop ----
    const int x = x0 + 4 * lane;
    const bool store = x < gw;
    uint8_t* dst = blur + (long long)(b0 + b) * blur_fstride + g.off + x + (long long)(y0 - 6) * gpitch;   // running pointer: output row of the next input row
    const uint32_t Q0 = 18u | (34u << 8) | (48u << 16) | (56u << 24);     // taps -3..0
    const uint32_t Q1 = 48u | (34u << 8) | (18u << 16);                   // taps +1..+3 (4th byte unused)
    // Vertical pass on PAIRS of rows: the horizontal sums fit 16 bits (<= 65280), so row r is packed with row r-1 as it arrives
    // (P = h[r-1] | h[r] << 16) and an output row is three 2-way integer dot products (IDP.2A: 16-bit sums x 8-bit taps) plus one multiply:
    //   v(o) = (h[o-3], h[o-2]).(18, 34) + (h[o-1], h[o]).(48, 56) + (h[o+1], h[o+2]).(48, 34) + 18 h[o+3] + 32768
    // The ring holds the packed pairs of the last 6 input rows (static slots after unrolling).
    const uint32_t WA = 18u | (34u << 8), WB = 48u | (56u << 8), WC = 48u | (34u << 8);
    uint32_t P[6][4];
    uint32_t hprev[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int j = 0; j < 6; ++j) { P[j][0] = P[j][1] = P[j][2] = P[j][3] = 0; }
    const uint32_t* rowp = sm32 + 3 + lane;                               // words (x - 4, x, x + 4) of the staged row
    const int ngroups = (nrows + 5) / 6;
    for (int gi = 0; gi < ngroups; ++gi) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {                                     // ring slot j is static after unrolling: no register moves
            const uint32_t w0 = rowp[j * (BLUR_BOX_W / 4)], w1 = rowp[j * (BLUR_BOX_W / 4) + 1], w2 = rowp[j * (BLUR_BOX_W / 4) + 2];
            // h(x+k) = sum_{i=0..6} q[i] * px(x+k-3+i): two 4-tap integer dot products on byte-aligned windows
            uint32_t hh[4];
            hh[0] = __dp4a(__byte_perm(w0, w1, 0x4321), Q0, __dp4a(__byte_perm(w1, w2, 0x4321), Q1, 0u));
            hh[1] = __dp4a(__byte_perm(w0, w1, 0x5432), Q0, __dp4a(__byte_perm(w1, w2, 0x5432), Q1, 0u));
            hh[2] = __dp4a(__byte_perm(w0, w1, 0x6543), Q0, __dp4a(__byte_perm(w1, w2, 0x6543), Q1, 0u));
            hh[3] = __dp4a(w1, Q0, __dp4a(w2, Q1, 0u));
            uint32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                P[j][k] = __byte_perm(hprev[k], hh[k], 0x5410);           // rows (r-1, r)
                hprev[k] = hh[k];
                // rows (r-6, r-5) = slot j+1, (r-4, r-3) = slot j+3, (r-2, r-1) = slot j+5 (mod 6), row r alone
                v[k] = __dp2a_lo(P[(j + 1) % 6][k], WA, __dp2a_lo(P[(j + 3) % 6][k], WB, __dp2a_lo(P[(j + 5) % 6][k], WC, 18u * hh[k] + 32768u)));
            }
            // (v + 32768) >> 16 is byte 2 of each sum (v < 2^24)
            const uint32_t outw = __byte_perm(__byte_perm(v[0], v[1], 0x0062), __byte_perm(v[2], v[3], 0x0062), 0x5410);
            const int i = gi * 6 + j;                                     // staged row; completes output row y0 + i - 6
            if (store && i >= 6 && i < nrows)
                *reinterpret_cast<uint32_t*>(dst) = outw;                // bytes past the level width land in row padding
            dst += gpitch;
        }
        rowp += 6 * (BLUR_BOX_W / 4);
    }
}

// =================================================================================================
// IC_Angle (ORBextractor.cc:108-161): intensity-centroid moments over the radius-15 disc, one warp per
// keypoint, lane = column u in [-15,15], loop over the 31 rows (coalesced 31-byte row reads), integer
// moments reduced with shuffles, angle = fastAtan2((float)m01, (float)m10) in degrees.
// =================================================================================================
__device__ char4 g_pattern_t[8 * 32];    // rBRIEF pairs, transposed: [bit k][byte i] = (x0,y0,x1,y1) of pair 8*i+k

// dot product of 4 unsigned bytes (pixels) with 4 signed bytes (weights)
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// 27 lanes = 3 rows x 9 aligned 32-bit words cover one 31-pixel patch row each; 11 steps walk the 31 rows.
// Per word: circle mask by SWAR compare of |u| against umax[|v|], then two integer dot products (sum of u*I and sum of I).
__device__ __forceinline__ float ic_angle_warp(const uint8_t* center, int pitch, int lane) {
    const unsigned long long UMAX_NIBBLES = 0x3689ABCDDEEEFFFFull;           // umax[|v|], 4 bits each (ORBextractor.cc:579-608)
    const int rg = lane / 9, wi = lane - rg * 9;
    const uint8_t* row0 = center - ORBX_HALF_PATCH;
    const int a = (int)((uintptr_t)row0 & 3);                                 // same for every row: pitches are multiples of 4
    const int u0 = 4 * wi - ORBX_HALF_PATCH - a;                              // u of byte 0 of this lane's word
    const uint32_t wu = (uint32_t)((u0) & 0xFF) | ((uint32_t)((u0 + 1) & 0xFF) << 8) | ((uint32_t)((u0 + 2) & 0xFF) << 16) | ((uint32_t)((u0 + 3) & 0xFF) << 24);
    const uint32_t absu = __vabsdiffu4(wu ^ 0x80808080u, 0x80808080u);       // |u| per byte (bias trick: u + 128 is unsigned)
    const uint8_t* p = row0 - a + 4 * wi + (long long)(rg - ORBX_HALF_PATCH) * pitch;
    int m10 = 0, m01 = 0;
    if (lane < 27) {
#pragma unroll
        for (int it = 0; it < 11; ++it) {
            const int v = it * 3 + rg - ORBX_HALF_PATCH;
            if (v <= ORBX_HALF_PATCH) {
                const uint32_t um = (uint32_t)(UMAX_NIBBLES >> (4 * abs(v))) & 15u;
                const uint32_t inside = ((um * 0x01010101u + 0x80808080u) - absu) & 0x80808080u;     // 0x80 where |u| <= umax
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p)) & ((inside >> 7) * 255u);
                m10 = dp4a_us(w, wu, m10);                                      // unsigned pixels x signed u
                m01 += v * (int)__dp4a(w, 0x01010101u, 0u);
            }
            p += 3 * (long long)pitch;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
    return fast_atan2_deg((float)m01, (float)m10);
}

// computeOrbDescriptor (ORBextractor.cc:173-227): lane i produces descriptor byte i.
// The rotated pattern stays within +-18 px of the keypoint, so the warp first stages the 37-row x 40-byte (10 aligned
// words) blurred window into shared memory with coalesced loads and then gathers its 512 samples from there: the
// random byte gathers cost shared-memory bank cycles instead of one L1 wavefront per lane.
#define BRIEF_R 18
#define BRIEF_ROWS (2 * BRIEF_R + 1)
#define BRIEF_PS 80                  // bytes per staged row: 64 (4 x 16-byte chunks cover 37 px at any alignment) + 16 padding to spread banks
// asynchronous staging (LDGSTS.128): issued before the orientation is computed so that the copy overlaps IC_Angle
__device__ __forceinline__ void brief_stage(const uint8_t* center, int pitch, int lane, uint8_t* sm) {
    const uint8_t* row0 = center - BRIEF_R;
    const int al = (int)((uintptr_t)row0 & 15);
    const uint8_t* base = row0 - al - BRIEF_R * pitch + 16 * (lane & 3);
    uint8_t* dst = sm + 16 * (lane & 3);
#pragma unroll
    for (int it = 0; it < (BRIEF_ROWS * 4 + 31) / 32; ++it) {
        const int r = it * 8 + (lane >> 2);
        if (r < BRIEF_ROWS) __pipeline_memcpy_async(dst + r * BRIEF_PS, base + r * pitch, 16);
    }
    __pipeline_commit();
}
__device__ __forceinline__ uint32_t brief_byte(const uint8_t* center, float a, float b, int lane, const uint8_t* sm) {
    const int al = (int)((uintptr_t)(center - BRIEF_R) & 15);
    __pipeline_wait_prior(0);
    __syncwarp();
    const uint8_t* c = sm + BRIEF_R * BRIEF_PS + BRIEF_R + al;
    uint32_t val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const char4 pt = g_pattern_t[k * 32 + lane];                      // coalesced (a lane-indexed __constant__ read would serialise 32-way)
        const float x0 = (float)pt.x, y0 = (float)pt.y, x1 = (float)pt.z, y1 = (float)pt.w;
        // center[cvRound(x*b + y*a)*step + cvRound(x*a - y*b)], unfused, round-half-even
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int q0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int q1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = c[r0 * BRIEF_PS + q0];
        const int t1 = c[r1 * BRIEF_PS + q1];
        val |= (uint32_t)(t0 < t1) << k;
    }
    __syncwarp();
    return val;
}

struct KpOut { float x, y, size, angle, response; int octave, class_id; };

// =================================================================================================
// K5/K7  orient_describe: one warp per keypoint slot of the level-keypoint array produced by the octree.
//   DESCRIBE = true : operator()(image, mask, keypoints, descriptors)  ORBextractor.cc:1544-1668 --
//                     angle on the raw level, 256-bit rBRIEF on the blurred level, pt *= scale for
//                     level > 0, level-major concatenation.
//   DESCRIBE = false: the keypoints-only overload (:1672-1686): level coordinates, angle, no descriptor.
// =================================================================================================
template <bool DESCRIBE>
__global__ void __launch_bounds__(128)
k_orient_describe(PyrView pv, const LevelGeom* __restrict__ levels, int nlevels, int kp_per_frame,
                  const uint32_t* __restrict__ kp_level, const int* __restrict__ kp_count,
                  const uint8_t* __restrict__ blur, long long blur_fstride,
                  KpOut* __restrict__ kp_out, uint8_t* __restrict__ desc_out, int cap, int* __restrict__ counts_out,
                  int* __restrict__ level_counts_out) {
    __shared__ __align__(16) uint8_t brief_sm[DESCRIBE ? 4 : 1][DESCRIBE ? BRIEF_ROWS * BRIEF_PS : 4];
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int b = blockIdx.y;
    if (slot >= kp_per_frame) return;
    // level of this slot and output position of its first keypoint: lane l looks at level l, one ballot + two warp reductions
    const int* cnt = kp_count + b * nlevels;
    const int off_l = lane < nlevels ? levels[lane].kp_off : 0x7FFFFFFF;
    const int cnt_l = lane < nlevels ? cnt[lane] : 0;
    const int level = __popc(__ballot_sync(0xffffffffu, slot >= off_l)) - 1;
    const int base = __reduce_add_sync(0xffffffffu, lane < level ? cnt_l : 0);
    const int total = __reduce_add_sync(0xffffffffu, cnt_l);
    const LevelGeom& g = levels[level];
    const int k = slot - g.kp_off;
    if (slot == 0 && lane == 0) {
        if (counts_out) counts_out[b] = total;                  // the true count: a value above cap tells the caller that keypoints were dropped
        if (level_counts_out) for (int l = 0; l < nlevels; ++l) level_counts_out[b * nlevels + l] = cnt[l];
    }
    if (k >= __shfl_sync(0xffffffffu, cnt_l, level)) return;
    const int oi = base + k;
    if (oi >= cap) return;                                      // caller capacity: counts_out[b] > cap reports it (ORBX_E_CAPACITY from the host calls)
    const uint32_t p = kp_level[(long long)b * kp_per_frame + slot];
    const int x = (int)(p & 0xFFF) + g.minBX, y = (int)((p >> 12) & 0xFFF) + g.minBY;   // :1184-1185
    int pitch;
    const uint8_t* img = level_ptr(pv, g, level, b, pitch);
    const uint8_t* bl = DESCRIBE ? blur + (long long)b * blur_fstride + g.off + (long long)y * g.pitch + x : nullptr;
    if (DESCRIBE) brief_stage(bl, g.pitch, lane, brief_sm[threadIdx.x >> 5]);
    const float angle = ic_angle_warp(img + (long long)y * pitch + x, pitch, lane);
    if (DESCRIBE) {
        const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);      // :164
        float sn, cs;
        det_sincos(__fmul_rn(angle, factorPI), &sn, &cs);
        const uint32_t byte = brief_byte(bl, cs, sn, lane, brief_sm[threadIdx.x >> 5]);
        desc_out[((long long)b * cap + oi) * 32 + lane] = (uint8_t)byte;
    }
    if (lane == 0) {
        KpOut o;
        o.x = (float)x; o.y = (float)y;
        if (DESCRIBE && level != 0) { o.x = __fmul_rn(o.x, g.scale); o.y = __fmul_rn(o.y, g.scale); }   // :1651-1660
        o.size = g.kp_size; o.angle = angle; o.response = (float)(p >> 24); o.octave = level; o.class_id = -1;
        kp_out[(long long)b * cap + oi] = o;
    }
}

// ProcessDesp (ORBextractor.cc:1747-1820): descriptors for caller-supplied per-level keypoints (level
// coordinates, angles as given), on the resident pyramid of frame 0.  One warp per keypoint.
__global__ void __launch_bounds__(128)
k_describe_given(const LevelGeom* __restrict__ levels, int nlevels, const KpOut* __restrict__ kp_in, int n,
                 const uint8_t* __restrict__ blur, KpOut* __restrict__ kp_out, uint8_t* __restrict__ desc_out) {
    const int lane = threadIdx.x & 31;
    __shared__ __align__(16) uint8_t brief_sm[4][BRIEF_ROWS * BRIEF_PS];
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= n) return;
    KpOut kp = kp_in[i];
    const int level = kp.octave;
    const LevelGeom& g = levels[level];
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float sn, cs;
    det_sincos(__fmul_rn(kp.angle, factorPI), &sn, &cs);
    const int x = __float2int_rn(kp.x), y = __float2int_rn(kp.y);
    const uint8_t* bl = blur + g.off + (long long)y * g.pitch + x;
    brief_stage(bl, g.pitch, lane, brief_sm[threadIdx.x >> 5]);
    const uint32_t byte = brief_byte(bl, cs, sn, lane, brief_sm[threadIdx.x >> 5]);
    desc_out[(long long)i * 32 + lane] = (uint8_t)byte;
    if (lane == 0) {
        if (level != 0) { kp.x = __fmul_rn(kp.x, g.scale); kp.y = __fmul_rn(kp.y, g.scale); }
        kp_out[i] = kp;
    }
}

// single-frame result block [count | overflow flag | pad | keypoint slots | descriptor slots] for one download (orbx_extract)
__global__ void __launch_bounds__(256)
k_gather_result(const int* __restrict__ count, const int* __restrict__ overflow, const uint32_t* __restrict__ kp, int kp_words,
                const uint32_t* __restrict__ desc, int desc_words, uint32_t* __restrict__ out) {
    const int i = blockIdx.x * 256 * 4 + threadIdx.x;
    if (i == 0) { out[0] = (uint32_t)count[0]; out[1] = (uint32_t)overflow[0]; out[2] = 0; out[3] = 0; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = i + k * 256;
        if (j < kp_words) out[4 + j] = kp[j];
        else if (j < kp_words + desc_words) out[4 + j] = desc[j - kp_words];
    }
}
