// k_frame.cuh -- device-resident Frame: the steps between extractor and matchers.
//   Frame::UndistortKeyPoints      /root/reference/src/Frame.cc:1052-1117   (cv::undistortPoints(mat, mat, mK, mDistCoef, Mat(), mK))
//   Frame::ComputeImageBounds      /root/reference/src/Frame.cc:1120-1176
//   Frame::ComputeStereoFromRGBD   /root/reference/src/Frame.cc:1576-1614
// cv::undistortPoints is OpenCV's (un-vendored) iterative inverse of the Brown model.  Its arithmetic, pinned against
// cv2 4.13 golden vectors (tests/golden/frame_cv2.npz): everything in double, no contraction; x = (u - cx) * (1/fx);
// 5 fixed-point iterations  icdist = (1 + ((k6 r2 + k5) r2 + k4) r2) / (1 + ((k3 r2 + k2) r2 + k1) r2),
// dX = 2 p1 x y + p2 (r2 + 2 x x) + s1 r2 + s2 r2 r2,  x = (x0 - dX) icdist  (rational / thin-prism terms are zero for
// the 5-coefficient model the reference configures); icdist < 0 restarts from the normalised point and stops;
// re-projection with P = mK: xx = fx x + 0 y + cx, ww = 1 / (0 x + 0 y + 1); result rounded to float.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "k_match.cuh"

struct CamDev {                    // doubles converted from the caller's floats exactly as cv::Mat::convertTo(CV_64F) does
    double fx, fy, cx, cy, ifx, ify, k1, k2, p1, p2, k3;
    int distorted;                 // mDistCoef.at<float>(0) != 0.0   (:1058)
    float bf;
};

__device__ __forceinline__ void undistort_point(const CamDev& c, float uf, float vf, float& ox, float& oy) {
    const double u = (double)uf, v = (double)vf;
    double x = __dmul_rn(__dsub_rn(u, c.cx), c.ifx), y = __dmul_rn(__dsub_rn(v, c.cy), c.ify);
    const double x0 = x, y0 = y;
#pragma unroll 1
    for (int j = 0; j < 5; ++j) {
        const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(0.0, r2), 0.0), r2), 0.0), r2));
        const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(c.k3, r2), c.k2), r2), c.k1), r2));
        const double icdist = __ddiv_rn(num, den);
        if (icdist < 0.0) { x = x0; y = y0; break; }
        // deltaX = 2*k[2]*x*y + k[3]*(r2 + 2*x*x) + k[8]*r2 + k[9]*r2*r2   (left-to-right, k[8..11] = 0)
        const double r4z = __dmul_rn(__dmul_rn(0.0, r2), r2), r2z = __dmul_rn(0.0, r2);
        const double dX = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, c.p1), x), y),
                                                         __dmul_rn(c.p2, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x)))), r2z), r4z);
        const double dY = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c.p1, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))),
                                                         __dmul_rn(__dmul_rn(__dmul_rn(2.0, c.p2), x), y)), r2z), r4z);
        x = __dmul_rn(__dsub_rn(x0, dX), icdist);
        y = __dmul_rn(__dsub_rn(y0, dY), icdist);
    }
    const double xx = __dadd_rn(__dadd_rn(__dmul_rn(c.fx, x), __dmul_rn(0.0, y)), c.cx);
    const double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(c.fy, y)), c.cy);
    const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(0.0, y)), 1.0));
    ox = (float)__dmul_rn(xx, ww); oy = (float)__dmul_rn(yy, ww);
}

// mvKeysUn, mvuRight, mvDepth for N keypoints.  depth_mode 0: no depth (both -1); 1: depth image (pitch in floats);
// 2: depth[i] gathered by the caller.
__global__ void __launch_bounds__(128)
k_frame_undistort_stereo(const KpM* __restrict__ keys, int n, CamDev cam, const float* __restrict__ depth, int depth_mode, int depth_pitch,
                         int rows, int cols, KpM* __restrict__ keys_un, float* __restrict__ u_right, float* __restrict__ depth_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    KpM kp = keys[i];
    const float u = kp.x, v = kp.y;
    if (cam.distorted) undistort_point(cam, u, v, kp.x, kp.y);
    keys_un[i] = kp;
    float ur = -1.f, dd = -1.f;
    if (depth_mode) {
        float d;
        if (depth_mode == 1) {
            const int r = (int)v, c = (int)u;                              // imDepth.at<float>(v, u): float -> int truncation (:1595)
            d = (r >= 0 && r < rows && c >= 0 && c < cols) ? depth[(size_t)r * depth_pitch + c] : 0.f;   // keypoints always lie inside the image
        } else d = depth[i];
        if (d > 0.f) { dd = d; ur = __fsub_rn(kp.x, __fdiv_rn(cam.bf, d)); }                               // :1603-1608
    }
    u_right[i] = ur; depth_out[i] = dd;
}

// ComputeStereoFromRGBD alone (mvKeysUn already computed): depth looked up at the RAW keypoint, uRight from the undistorted x
__global__ void __launch_bounds__(128)
k_frame_rgbd(const KpM* __restrict__ keys, const KpM* __restrict__ keys_un, int n, float bf, const float* __restrict__ depth, int depth_mode, int depth_pitch,
             int rows, int cols, float* __restrict__ u_right, float* __restrict__ depth_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float d;
    if (depth_mode == 1) {
        const int r = (int)keys[i].y, c = (int)keys[i].x;
        d = (r >= 0 && r < rows && c >= 0 && c < cols) ? depth[(size_t)r * depth_pitch + c] : 0.f;
    } else d = depth[i];
    float ur = -1.f, dd = -1.f;
    if (d > 0.f) { dd = d; ur = __fsub_rn(keys_un[i].x, __fdiv_rn(bf, d)); }
    u_right[i] = ur; depth_out[i] = dd;
}

// ComputeImageBounds with distortion: the four image corners through undistortPoints (:1136-1163)
__global__ void k_frame_bounds(CamDev cam, int rows, int cols, float* __restrict__ out /* minX, maxX, minY, maxY */) {
    if (threadIdx.x != 0) return;
    float x[4], y[4];
    undistort_point(cam, 0.f, 0.f, x[0], y[0]);
    undistort_point(cam, (float)cols, 0.f, x[1], y[1]);
    undistort_point(cam, 0.f, (float)rows, x[2], y[2]);
    undistort_point(cam, (float)cols, (float)rows, x[3], y[3]);
    out[0] = fminf(x[0], x[2]); out[1] = fmaxf(x[1], x[3]); out[2] = fminf(y[0], y[1]); out[3] = fmaxf(y[2], y[3]);
}
