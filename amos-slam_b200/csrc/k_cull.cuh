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

// half-widths of the 31x31 MORPH_ELLIPSE rows (cv::getStructuringElement), filled by the host
__constant__ int c_ell_dx[31];

// grey-scale dilate / erode with the 31x31 ellipse; pixels outside the image are ignored
// (cv::morphologyDefaultBorderValue).  CTA = 32x16 outputs, input tile + 15-px halo in shared memory.
#define MORPH_TW 32
#define MORPH_TH 16
template <bool DILATE>
__global__ void __launch_bounds__(MORPH_TW * MORPH_TH)
k_morph_ellipse31(const uint8_t* __restrict__ src, int spitch, uint8_t* __restrict__ dst, int dpitch, int w, int h) {
    __shared__ uint8_t tile[MORPH_TH + 30][MORPH_TW + 32];
    const int x0 = blockIdx.x * MORPH_TW, y0 = blockIdx.y * MORPH_TH;
    const int tid = threadIdx.y * MORPH_TW + threadIdx.x;
    const uint8_t neutral = DILATE ? 0 : 255;
    for (int i = tid; i < (MORPH_TH + 30) * (MORPH_TW + 30); i += MORPH_TW * MORPH_TH) {
        const int r = i / (MORPH_TW + 30), c = i - r * (MORPH_TW + 30);
        const int yy = y0 + r - 15, xx = x0 + c - 15;
        tile[r][c] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? src[(size_t)yy * spitch + xx] : neutral;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= w || y >= h) return;
    int acc = neutral;
#pragma unroll 1
    for (int dy = 0; dy < 31; ++dy) {
        const int dx = c_ell_dx[dy];
        const uint8_t* row = &tile[threadIdx.y + dy][threadIdx.x + 15 - dx];
        for (int k = 0; k <= 2 * dx; ++k) acc = DILATE ? max(acc, (int)row[k]) : min(acc, (int)row[k]);
    }
    dst[(size_t)y * dpitch + x] = (uint8_t)acc;
}

struct KpIn { float x, y, size, angle, response; int octave, class_id; };

// per-keypoint lookup (:1718-1741): cull if closing(p) != 0 or rm_vector[centers[label(p)-1].id] == 1,
// p = (int)(pt * scale).  flags[i] = 1 => culled.  Out-of-range label / id indices (undefined behaviour
// in the reference) are treated as "not flagged".
__global__ void k_cull_flags(const KpIn* __restrict__ kp, const float* __restrict__ kp_scale, int n,
                             const uint8_t* __restrict__ closed, int cpitch, const double* __restrict__ label, int lpitch_elems,
                             int rows, int cols, const int* __restrict__ centers_id, int ncenters,
                             const int* __restrict__ rm_vector, int nrm, uint8_t* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s = kp_scale[i];
    const int px = (int)__fmul_rn(kp[i].x, s), py = (int)__fmul_rn(kp[i].y, s);
    int flag = 0;
    if (px >= 0 && py >= 0 && px < cols && py < rows) {
        const double sp = label[(size_t)py * lpitch_elems + px];
        const double idxd = sp - 1.0;
        if (idxd >= 0.0 && idxd < (double)ncenters) {
            const int id = centers_id[(size_t)idxd];
            if (id >= 0 && id < nrm && rm_vector[id] == 1) flag = 1;
        }
        if (closed[(size_t)py * cpitch + px] != 0) flag = 1;
    }
    flags[i] = (uint8_t)flag;
}
