// orbx_internal.h -- private interface between the two translation units of liborbx_b200.so (not exported).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "orbx_common.cuh"

struct orbx_extractor;
// resident pyramid of frame 0 of an extractor handle (mvImagePyramid), device pointers
struct OrbxPyramidInfo {
    int device, nlevels; cudaStream_t stream;
    const uint8_t* ptr[ORBX_MAX_LEVELS]; int pitch[ORBX_MAX_LEVELS], w[ORBX_MAX_LEVELS], h[ORBX_MAX_LEVELS];
    const float* scale; const float* inv_scale;     // host arrays owned by the handle (mvScaleFactor / mvInvScaleFactor)
};
int orbx_internal_pyramid(orbx_extractor* h, OrbxPyramidInfo* out);

// keypoints / descriptors of frame 0 of the handle's last orbx_extract or orbx_describe call, still on the device
struct OrbxLastResult {
    int device, n, nlevels; cudaStream_t stream;
    const void* keys;            // n x 28-byte cv::KeyPoint records
    const uint8_t* desc;         // n x 32
    const float* scale;          // host: mvScaleFactor[nlevels]
};
int orbx_internal_last_result(orbx_extractor* h, OrbxLastResult* out);

// result and pyramids of the handle's last BATCHED extract call (frames 0 .. B-1), still on the device: frame b's level l starts at
// ptr[l] + b * fstride[l]; its keypoints at keys + b * cap (28-byte records), descriptors at desc + b * cap * 32, count at counts[b]
struct OrbxBatchInfo {
    int device, nlevels, B, cap; cudaStream_t stream;
    const uint8_t* ptr[ORBX_MAX_LEVELS]; long long fstride[ORBX_MAX_LEVELS]; int pitch[ORBX_MAX_LEVELS], w[ORBX_MAX_LEVELS], h[ORBX_MAX_LEVELS];
    const void* keys; const uint8_t* desc; const int* counts;
    const float* scale; const float* inv_scale;
};
int orbx_internal_last_batch(orbx_extractor* h, OrbxBatchInfo* out);
