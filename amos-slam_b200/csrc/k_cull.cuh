// k_cull.cuh -- Amos-SLAM dynamic-mask keypoint culling (ORBextractor::MovingKeyPoints,
// /root/reference/src/ORBextractor.cc:1688-1745) and the REFLECT_101 border export of mvImagePyramid.
#pragma once
#include "orbx_common.cuh"

// copyMakeBorder(BORDER_REFLECT_101) of one level into a tight (w+2b) x (h+2b) buffer  (:1859-1882)
__global__ void k_border101(const uint8_t* __restrict__ src, int pitch, int w, int h, int border, uint8_t* __restrict__ dst, int W, int H) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    int sx = x - border, sy = y - border;
    // gfedcb|abcdefgh|gfedcba ; iterate for borders wider than the image
    while (sx < 0 || sx >= w) { if (sx < 0) sx = -sx; else sx = 2 * w - 2 - sx; if (w == 1) { sx = 0; break; } }
    while (sy < 0 || sy >= h) { if (sy < 0) sy = -sy; else sy = 2 * h - 2 - sy; if (h == 1) { sy = 0; break; } }
    dst[(size_t)y * W + x] = src[(size_t)sy * pitch + sx];
}

// Frames whose rows are not 16-byte aligned (a 1241-px KITTI row) are re-pitched on the device before TMA reads them: ONE launch for the whole batch
// (a cudaMemcpy2DAsync per frame cost ~4 us each: 0.25 ms per 64 KITTI frames, more than their FAST stage's share of a launch).  One thread per 4 destination
// bytes; the source words are read aligned and funnel-shifted into place.
__global__ void __launch_bounds__(256)
k_repitch(const uint8_t* __restrict__ src, long long src_fstride, long long src_step, uint8_t* __restrict__ dst, long long dst_fstride, int dst_pitch, int cols, int rows) {
    const int x4 = (blockIdx.x * 256 + threadIdx.x) * 4, y = blockIdx.y, b = blockIdx.z;
    if (x4 >= cols) return;
    const uint8_t* s = src + (long long)b * src_fstride + (long long)y * src_step + x4;
    uint8_t* d = dst + (long long)b * dst_fstride + (long long)y * dst_pitch + x4;
    if (x4 >= 4 && x4 + 8 <= cols) {                                         // both aligned words lie inside this row (never in front of / behind the caller's buffer)
        const unsigned a = (unsigned)((uintptr_t)s & 3);
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(s - a);
        const uint32_t w0 = __ldg(sw);
        const uint32_t v = a ? __funnelshift_r(w0, __ldg(sw + 1), 8 * a) : w0;
        *reinterpret_cast<uint32_t*>(d) = v;
    } else {
        for (int k = 0; k < 4 && x4 + k < cols; ++k) d[k] = s[k];
    }
}

// half-widths of the 31x31 MORPH_ELLIPSE rows (cv::getStructuringElement), filled by the host
__constant__ int c_ell_dx[31];

// -------------------------------------------------------------------------------------------------
// closing = erode(dilate(mask, ellipse31), ellipse31)  (:1697-1704) is only ever tested for "!= 0" (:1734-1735), and only at keypoint positions.
// With cv's default border (outside pixels ignored by both passes) that predicate depends only on the binary
// image (mask != 0):  closing(p) != 0  <=>  every q in p+E has some r in q+E with mask(r) != 0.  So the masks
// are bit-packed (1 bit per pixel, 32 pixels per word, LSB = leftmost), the dilation D = dilate(M) is one binary pass over the plane
// (k_bin_dilate31: per ellipse row the window of half-width dx is OR-ed over the (left, mid, right) words with a doubling OR), and the
// erosion is evaluated per keypoint (closing_nonzero_at).  This replaces two 729-tap grey-scale max / min passes per pixel.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t swar_nz_byte(uint32_t x) { return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; }
__device__ __forceinline__ uint32_t nz_nibble(uint32_t w) { return (((swar_nz_byte(w) >> 7) * 0x00204081u) >> 21) & 0xFu; }   // 4 bytes -> 4 bits

// mask: [B] frames of rows x pitch bytes (pitch multiple of 32, rows of padding past `cols` may hold garbage)
__global__ void k_mask_pack(const uint8_t* __restrict__ mask, long long fstride, int pitch, int rows, int cols,
                            uint32_t* __restrict__ bits, int wpr) {
    const int wx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (wx >= wpr) return;
    const uint4* src = reinterpret_cast<const uint4*>(mask + (long long)b * fstride + (long long)y * pitch + 32 * wx);
    const uint4 a = __ldg(src), c = __ldg(src + 1);
    uint32_t w = nz_nibble(a.x) | (nz_nibble(a.y) << 4) | (nz_nibble(a.z) << 8) | (nz_nibble(a.w) << 12) |
                 (nz_nibble(c.x) << 16) | (nz_nibble(c.y) << 20) | (nz_nibble(c.z) << 24) | (nz_nibble(c.w) << 28);
    const int valid = cols - 32 * wx;
    if (valid < 32) w &= (1u << valid) - 1u;
    bits[((long long)b * rows + y) * wpr + wx] = w;
}

// half-width of the ellipse row at distance k from the centre row: round(sqrt(225 - k^2)) -- the values upload_ellipse() computes with OpenCV's formula and
// checks against this table, which lets the dilation kernel unroll over the rows with every window width a compile-time constant
__host__ __device__ constexpr int ell_dx(int k) {
    return k <= 3 ? 15 : k <= 6 ? 14 : k <= 8 ? 13 : k == 9 ? 12 : k == 10 ? 11 : k == 11 ? 10 : k == 12 ? 9 : k == 13 ? 7 : k == 14 ? 5 : 0;
}

__device__ __forceinline__ unsigned long long or_window(unsigned long long x, int n) {   // OR of x >> k, k = 0..n-1, 1 <= n <= 16
    unsigned long long t = x; int p = 1;
    if (n >= 2) { t |= t >> 1; p = 2; }
    if (n >= 4) { t |= t >> 2; p = 4; }
    if (n >= 8) { t |= t >> 4; p = 8; }
    if (n >= 16) { t |= t >> 8; p = 16; }
    return t | (t >> (n - p));
}

// out = dilate(in) with the 31x31 ellipse.  One CTA per DIL_TW x DIL_TR tile of output words: the (DIL_TR + 30) x (DIL_TW + 2) input words it reads are staged
// in shared memory once (the one-thread-per-word form read every input word 93 times through L2 and was bound by that: 147 us per 32 frames of 1080p).
// Rows y-k and y+k share the half-width dx[15-k], and consecutive k often do too (15: k=0..3, 14: k=4..6, 13: k=7,8, ...): dilation distributes over OR,
// so the rows of one half-width are OR-ed first and dilated horizontally once (10 window passes per output word instead of 31).
#define DIL_TW 32
#define DIL_TR 32
__global__ void __launch_bounds__(256)
k_bin_dilate31(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int wpr, int rows, int cols) {
    __shared__ uint32_t sm[DIL_TR + 30][DIL_TW + 2];
    const int wx0 = blockIdx.x * DIL_TW, y0 = blockIdx.y * DIL_TR, b = blockIdx.z;
    const uint32_t* plane = in + (long long)b * rows * wpr;
    for (int i = threadIdx.x; i < (DIL_TR + 30) * (DIL_TW + 2); i += 256) {
        const int r = i / (DIL_TW + 2), c = i - r * (DIL_TW + 2);
        const int yy = y0 - 15 + r, wx = wx0 - 1 + c;
        sm[r][c] = (yy >= 0 && yy < rows && wx >= 0 && wx < wpr) ? __ldg(plane + (long long)yy * wpr + wx) : 0u;    // (bits past `cols` are 0 in the packed mask)
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, wx = wx0 + tx;
    if (wx >= wpr) return;
    const uint32_t lastmask = (cols & 31) ? ((1u << (cols & 31)) - 1u) : 0xFFFFFFFFu;
    for (int ty = threadIdx.x >> 5; ty < DIL_TR; ty += 8) {
        const int y = y0 + ty;
        if (y >= rows) break;
        uint32_t acc = 0, gL = 0, gM = 0, gR = 0;
#pragma unroll
        for (int k = 15; k >= 0; --k) {                                  // fully unrolled: ell_dx(k) and the group boundaries are compile-time constants
            const uint32_t* r0 = &sm[ty + 15 - k][tx];
            gL |= r0[0]; gM |= r0[1]; gR |= r0[2];
            if (k) { const uint32_t* r1 = &sm[ty + 15 + k][tx]; gL |= r1[0]; gM |= r1[1]; gR |= r1[2]; }
            if (k == 0 || ell_dx(k - 1) != ell_dx(k)) {                  // last row pair of this half-width: dilate the collected rows horizontally once
                const int gd = ell_dx(k);
                const unsigned long long A = ((unsigned long long)gM << 32) | gL, Bw = ((unsigned long long)gR << 32) | gM;
                acc |= (uint32_t)(or_window(A >> (32 - gd), gd + 1) | or_window(Bw, gd + 1));
                gL = gM = gR = 0;
            }
        }
        if (wx == wpr - 1) acc &= lastmask;
        out[((long long)b * rows + y) * wpr + wx] = acc;
    }
}

// closing(p) != 0 for ONE pixel, from D = dilate(mask != 0): the erosion is only ever consulted at keypoint positions (:1734-1735), so instead of eroding the
// whole plane, every keypoint tests its own 31 ellipse rows: closing(p) != 0  <=>  D(q) = 1 for every q of p + E inside the image (cv's default border:
// outside pixels are ignored by the erosion).
__device__ __forceinline__ bool closing_nonzero_at(const uint32_t* __restrict__ D, int wpr, int rows, int cols, int px, int py) {
    // all 31 rows' words are fetched before any is tested (62 independent loads in flight: the plane is L2-resident, a row-by-row loop paid its latency 31 times)
    uint32_t lo[31], hi[31];
#pragma unroll
    for (int k = -15; k <= 15; ++k) {
        const int yy = min(max(py + k, 0), rows - 1);
        const int dx = c_ell_dx[15 + k];
        const int w0 = max(px - dx, 0) >> 5;
        const uint32_t* row = D + (long long)yy * wpr;
        lo[k + 15] = __ldg(row + w0); hi[k + 15] = (w0 + 1 < wpr) ? __ldg(row + w0 + 1) : 0u;
    }
    bool all = true;
#pragma unroll
    for (int k = -15; k <= 15; ++k) {
        const int yy = py + k;
        const int dx = c_ell_dx[15 + k];
        const int x0 = max(px - dx, 0), x1 = min(px + dx, cols - 1);     // at most 31 pixels: two words
        const unsigned long long v = (((unsigned long long)hi[k + 15] << 32) | lo[k + 15]) >> (x0 & 31);
        const unsigned long long need = (1ull << (x1 - x0 + 1)) - 1ull;
        all = all && (yy < 0 || yy >= rows || (v & need) == need);
    }
    return all;
}

// Batched MovingKeyPoints on the per-level keypoint slots of the quadtree stage (:1718-1741): one warp per
// (level, frame) drops the keypoints whose position p = (int)(pt * scale) lies on a set bit of the closed mask, keeps the
// order of the survivors (stable erase) and rewrites the level's count.
// Optional super-pixel term (:1722-1736): labels = imLS as 16-bit super-pixel ids (1-based; the reference stores them as CV_64F), flagged[id - 1] =
// (rm_vector[centers[id - 1].id] == 1) per frame, prepared by the caller from the two small tables.  An id outside [1, n_labels] (undefined
// behaviour in the reference) counts as "not flagged".
struct LabelView { const uint16_t* labels; long long fstride; int pitch; const uint8_t* flagged; int n_labels; };

__global__ void __launch_bounds__(32)
k_cull_levelkp(const LevelGeom* __restrict__ levels, int nlevels, int kp_per_frame, uint32_t* __restrict__ kp_level, int* __restrict__ kp_count,
               const uint32_t* __restrict__ dilated, int wpr, int rows, int cols, LabelView lv, int* __restrict__ culled_count) {   // dilated = dilate(mask != 0), k_bin_dilate31
    const int level = blockIdx.x, b = blockIdx.y, lane = threadIdx.x;
    const LevelGeom& g = levels[level];
    uint32_t* kp = kp_level + (long long)b * kp_per_frame + g.kp_off;
    const int n = kp_count[b * nlevels + level];
    const uint32_t* plane = dilated + (long long)b * rows * wpr;
    const float scale = level ? g.scale : 1.0f;                          // :1712-1715
    int kept = 0;
    for (int k0 = 0; k0 < n; k0 += 32) {
        const int k = k0 + lane;
        uint32_t p = 0; bool keep = false;
        if (k < n) {
            p = kp[k];
            const float x = (float)((int)(p & 0xFFF) + g.minBX), y = (float)((int)((p >> 12) & 0xFFF) + g.minBY);
            const int px = (int)__fmul_rn(x, scale), py = (int)__fmul_rn(y, scale);
            keep = true;
            if (px >= 0 && py >= 0 && px < cols && py < rows) {
                keep = !closing_nonzero_at(plane, wpr, rows, cols, px, py);
                if (keep && lv.labels) {
                    const int id = lv.labels[(long long)b * lv.fstride + (long long)py * lv.pitch + px];
                    if (id >= 1 && id <= lv.n_labels && lv.flagged[(long long)b * lv.n_labels + id - 1]) keep = false;
                }
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) kp[kept + __popc(m & ((1u << lane) - 1u))] = p;        // kept + rank <= k: in-place compaction is safe after the sync
        kept += __popc(m);
    }
    if (lane == 0) {
        kp_count[b * nlevels + level] = kept;
        if (culled_count && n > kept) atomicAdd(culled_count + b, n - kept);
    }
}

struct KpIn { float x, y, size, angle, response; int octave, class_id; };

// per-keypoint lookup (:1718-1741): cull if closing(p) != 0 or rm_vector[centers[label(p)-1].id] == 1,
// p = (int)(pt * scale).  flags[i] = 1 => culled.  Out-of-range label / id indices (undefined behaviour
// in the reference) are treated as "not flagged".
__global__ void k_cull_flags(const KpIn* __restrict__ kp, const float* __restrict__ kp_scale, int n,
                             const uint32_t* __restrict__ dilated, int wpr, const uint8_t* __restrict__ rm_flag,
                             int rows, int cols, uint8_t* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s = kp_scale[i];
    const int px = (int)__fmul_rn(kp[i].x, s), py = (int)__fmul_rn(kp[i].y, s);
    int flag = rm_flag[i];                                   // rm_vector[centers[label(p) - 1].id] == 1, looked up by the caller-side marshalling
    if (px >= 0 && py >= 0 && px < cols && py < rows && closing_nonzero_at(dilated, wpr, rows, cols, px, py)) flag = 1;
    flags[i] = (uint8_t)flag;
}
