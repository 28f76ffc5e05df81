// k_describe.cuh -- 7x7 Gaussian blur, IC_Angle orientation and rotated-BRIEF descriptors.
#pragma once
#include "orbx_common.cuh"
#include "det_math.cuh"

// =================================================================================================
// K6  gauss7: cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on every pyramid level
// (/root/reference/src/ORBextractor.cc:1626-1634, 1791-1793), OpenCV's 8-bit fixed-point path
// (SURVEY.md A.2):  q = [18,34,48,56,48,34,18]/256;  h = sum q*px (16-bit range, no rounding);
// v = sum q*h (32-bit);  out = (v + 32768) >> 16.  Reflection is at the LEVEL edge (the reference
// blurs a clone of the ROI).
// One warp = one tile of 128 columns x BLUR_STRIP rows of one level of one frame (the tile list covers all
// levels, so a single launch blurs the whole pyramid of the whole batch).  Each lane owns 4 adjacent
// columns = one aligned 32-bit word per row: it loads ONLY its own word, takes the two neighbouring words
// from the adjacent lanes by shuffle, forms the 4 horizontal sums, keeps the last 7 rows of them in
// registers and emits one output word per row.  No shared memory; HBM traffic = read level + write level.
// Lanes whose 10-byte window crosses the image edge take a byte-wise reflect path (2 lanes per row).
// =================================================================================================
#define BLUR_STRIP 64
struct BlurTile { short level, xc, strip, pad; };

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;     // loops only on levels narrower than the 3-px halo
    return p;
}

__global__ void __launch_bounds__(128)
k_gauss7(PyrView pv, const LevelGeom* __restrict__ levels, const BlurTile* __restrict__ tiles, int ntiles,
         uint8_t* __restrict__ blur, long long blur_fstride) {
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (tile >= ntiles) return;
    const BlurTile t = tiles[tile];
    const int b = blockIdx.y;
    const LevelGeom& g = levels[t.level];
    int pitch;
    const uint8_t* img = level_ptr(pv, g, t.level, b, pitch);
    const int x = t.xc * 128 + lane * 4;
    const int y0 = t.strip * BLUR_STRIP, y1 = min(y0 + BLUR_STRIP, g.h);
    const bool valid = x < g.w;
    const bool inside_r = x + 6 <= g.w - 1;
    const bool fast = valid && x >= 4 && inside_r;                   // window x-3 .. x+6 entirely inside the level
    const bool left = valid && x == 0 && inside_r && g.w >= 4;       // left edge: pixels -3..-1 mirror bytes 3..1 of the own word
    int xo[10];                                                      // slow lanes (right edge, tiny levels): reflected columns, row-invariant
#pragma unroll
    for (int j = 0; j < 10; ++j) xo[j] = (valid && !fast && !left) ? reflect101(x - 3 + j, g.w) : 0;
    uint8_t* dst = blur + (long long)b * blur_fstride + g.off + x;
    int hb[7][4];
#pragma unroll
    for (int j = 0; j < 7; ++j) { hb[j][0] = hb[j][1] = hb[j][2] = hb[j][3] = 0; }
#pragma unroll 7
    for (int r = y0 - 3; r < y1 + 3; ++r) {
        const uint8_t* row = img + (long long)reflect101(r, g.h) * pitch;
        const uint32_t w1 = valid ? __ldg(reinterpret_cast<const uint32_t*>(row + x)) : 0u;
        uint32_t w0 = __shfl_up_sync(0xffffffffu, w1, 1), w2 = __shfl_down_sync(0xffffffffu, w1, 1);
        if (fast && lane == 0) w0 = __ldg(reinterpret_cast<const uint32_t*>(row + x - 4));
        if ((fast || left) && lane == 31) w2 = __ldg(reinterpret_cast<const uint32_t*>(row + x + 4));
        if (left) w0 = __byte_perm(w1, 0, 0x1230);                    // bytes (.,p3,p2,p1): reflect-101 of columns -3..-1
        int B[10];                                                    // pixels x-3 .. x+6
        if (fast || left) {
            B[0] = (w0 >> 8) & 0xFF; B[1] = (w0 >> 16) & 0xFF; B[2] = w0 >> 24;
            B[3] = w1 & 0xFF; B[4] = (w1 >> 8) & 0xFF; B[5] = (w1 >> 16) & 0xFF; B[6] = w1 >> 24;
            B[7] = w2 & 0xFF; B[8] = (w2 >> 8) & 0xFF; B[9] = (w2 >> 16) & 0xFF;
        } else if (valid) {
#pragma unroll
            for (int j = 0; j < 10; ++j) B[j] = row[xo[j]];
        } else {
#pragma unroll
            for (int j = 0; j < 10; ++j) B[j] = 0;
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) { hb[j][0] = hb[j + 1][0]; hb[j][1] = hb[j + 1][1]; hb[j][2] = hb[j + 1][2]; hb[j][3] = hb[j + 1][3]; }
#pragma unroll
        for (int k = 0; k < 4; ++k) hb[6][k] = 18 * (B[k] + B[k + 6]) + 34 * (B[k + 1] + B[k + 5]) + 48 * (B[k + 2] + B[k + 4]) + 56 * B[k + 3];
        const int o = r - 3;                                          // output row completed by this input row
        if (o >= y0 && valid) {
            uint32_t outw = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t v = 18u * (uint32_t)(hb[0][k] + hb[6][k]) + 34u * (uint32_t)(hb[1][k] + hb[5][k]) + 48u * (uint32_t)(hb[2][k] + hb[4][k]) + 56u * (uint32_t)hb[3][k];
                outw |= ((v + 32768u) >> 16) << (8 * k);
            }
            *reinterpret_cast<uint32_t*>(dst + (long long)o * g.pitch) = outw;   // bytes past the level width land in row padding
        }
    }
}

// =================================================================================================
// IC_Angle (ORBextractor.cc:108-161): intensity-centroid moments over the radius-15 disc, one warp per
// keypoint, lane = column u in [-15,15], loop over the 31 rows (coalesced 31-byte row reads), integer
// moments reduced with shuffles, angle = fastAtan2((float)m01, (float)m10) in degrees.
// =================================================================================================
__constant__ int c_umax[16];             // [15,15,15,15,14,14,14,13,13,12,11,10,9,8,6,3]
__constant__ char4 c_pattern_t[8 * 32];  // rBRIEF pairs, transposed: [bit k][byte i] = (x0,y0,x1,y1) of pair 8*i+k

__device__ __forceinline__ float ic_angle_warp(const uint8_t* center, int pitch, int lane) {
    const int u = lane - ORBX_HALF_PATCH;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int au = abs(u);
#pragma unroll 1
        for (int v = -ORBX_HALF_PATCH; v <= ORBX_HALF_PATCH; ++v) {
            if (au <= c_umax[abs(v)]) {
                const int val = __ldg(center + (long long)v * pitch + u);
                m10 += u * val; m01 += v * val;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
    return fast_atan2_deg((float)m01, (float)m10);
}

// computeOrbDescriptor (ORBextractor.cc:173-227): lane i produces descriptor byte i
__device__ __forceinline__ uint32_t brief_byte(const uint8_t* center, int pitch, float a, float b, int lane) {
    uint32_t val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const char4 pt = c_pattern_t[k * 32 + lane];
        const float x0 = (float)pt.x, y0 = (float)pt.y, x1 = (float)pt.z, y1 = (float)pt.w;
        // center[cvRound(x*b + y*a)*step + cvRound(x*a - y*b)], unfused, round-half-even
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int q0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int q1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = __ldg(center + (long long)r0 * pitch + q0);
        const int t1 = __ldg(center + (long long)r1 * pitch + q1);
        val |= (uint32_t)(t0 < t1) << k;
    }
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
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int b = blockIdx.y;
    if (slot >= kp_per_frame) return;
    int level = 0;
    while (level + 1 < nlevels && slot >= levels[level + 1].kp_off) ++level;
    const LevelGeom& g = levels[level];
    const int k = slot - g.kp_off;
    const int* cnt = kp_count + b * nlevels;
    int base = 0, total = 0;
    for (int l = 0; l < nlevels; ++l) { const int c = cnt[l]; if (l < level) base += c; total += c; }
    if (slot == 0 && lane == 0) {
        if (counts_out) counts_out[b] = min(total, cap);
        if (level_counts_out) for (int l = 0; l < nlevels; ++l) level_counts_out[b * nlevels + l] = cnt[l];
    }
    if (k >= cnt[level]) return;
    const int oi = base + k;
    if (oi >= cap) return;                                      // caller capacity (status reported by the host)
    const uint32_t p = kp_level[(long long)b * kp_per_frame + slot];
    const int x = (int)(p & 0xFFF) + g.minBX, y = (int)((p >> 12) & 0xFFF) + g.minBY;   // :1184-1185
    int pitch;
    const uint8_t* img = level_ptr(pv, g, level, b, pitch);
    const float angle = ic_angle_warp(img + (long long)y * pitch + x, pitch, lane);
    if (DESCRIBE) {
        const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);      // :164
        float sn, cs;
        det_sincos(__fmul_rn(angle, factorPI), &sn, &cs);
        const uint8_t* bl = blur + (long long)b * blur_fstride + g.off + (long long)y * g.pitch + x;
        const uint32_t byte = brief_byte(bl, g.pitch, cs, sn, lane);
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
    const uint32_t byte = brief_byte(bl, g.pitch, cs, sn, lane);
    desc_out[(long long)i * 32 + lane] = (uint8_t)byte;
    if (lane == 0) {
        if (level != 0) { kp.x = __fmul_rn(kp.x, g.scale); kp.y = __fmul_rn(kp.y, g.scale); }
        kp_out[i] = kp;
    }
}
