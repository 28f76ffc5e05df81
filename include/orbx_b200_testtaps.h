/* orbx_b200_testtaps.h -- stage taps of liborbx_b200.so for the parity tests (tests/ only).
 *
 * NOT part of the drop-in ABI (that is include/orbx_b200.h): these calls expose intermediate results of the extractor
 * so that every stage can be compared with the oracle on its own.  Host pointers, synchronous.
 */
#ifndef ORBX_B200_TESTTAPS_H
#define ORBX_B200_TESTTAPS_H

#include "orbx_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Operate on frame `b` of the last call.  Candidates are vToDistributeKeys of src/ORBextractor.cc:1073-1157 in reference order. */
int orbx_debug_level_candidates(orbx_extractor* h, int b, int level, orbx_keypoint* out, int cap, int* n_out);
int orbx_debug_blurred_level(orbx_extractor* h, int b, int level, uint8_t* dst, size_t dst_step);
int orbx_debug_pyramid_level(orbx_extractor* h, int b, int level, uint8_t* dst, size_t dst_step);
/* DistributeOctTree (src/ORBextractor.cc:706-1049) on caller-provided candidates (x,y integer-valued
 * floats relative to minX/minY, response) -- runs the same device kernels as the pipeline. */
int orbx_debug_distribute(orbx_extractor* h, const orbx_keypoint* cand, int ncand, int minX, int maxX,
                          int minY, int maxY, int N, orbx_keypoint* out, int cap, int* n_out);

/* The rotation of computeOrbDescriptor (src/ORBextractor.cc:178-181: cos / sin of the float angle = glibc sincosf) as the device
 * evaluates it, for the n floats whose bit patterns are lo_bits, lo_bits + 1, ... */
int orbx_debug_sincos(orbx_extractor* h, unsigned lo_bits, int n, float* sin_out, float* cos_out);

/* ORBX_CANARY=1: every device buffer of the extractor handles carries 256 guard bytes on both sides; returns how many were overwritten. */
int orbx_debug_canary_check(int* n_blocks);

#ifdef __cplusplus
}
#endif
#endif /* ORBX_B200_TESTTAPS_H */
