// orbx_common.cuh -- shared definitions of the B200-native ORB front-end (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define ORBX_MAX_LEVELS 32
#define ORBX_EDGE 19            // EDGE_THRESHOLD   /root/reference/src/ORBextractor.cc:93
#define ORBX_HALF_PATCH 15      // HALF_PATCH_SIZE  :92
#define ORBX_MAXD 13            // quadtree digits per key (levels up to 4095 px across need 12)
#define ORBX_ROOT_SHIFT (2 * ORBX_MAXD)
#define ORBX_MAX_DIM 4095       // packed candidate = x:12 | y:12 | response:8

// per-level geometry, built on the host by Plan (orbx_extractor.cu) with the reference's arithmetic
struct LevelGeom {
    int w, h, pitch;             // level size (cvRound(cols*invScale)), bytes per row in pyramid storage
    long long off;               // byte offset of the level inside one frame's pyramid block (level 0: unused)
    // FAST cell grid  (ORBextractor.cc:1067-1086)
    int minBX, minBY, maxBX, maxBY;
    int cell_begin, cell_count;  // range in the cell table (valid cells only, row-major)
    // quadtree (ORBextractor.cc:706-755)
    int N;                       // mnFeaturesPerLevel[level]
    int nIni; float hX;
    int cand_off, cand_cap;      // per-frame offsets into ordered / sorted candidate arrays
    int slot_off;                // per-frame offset of this level's first cell slot
    int kp_off, kp_cap;          // per-frame per-level keypoint slots
    float scale;                 // mvScaleFactor[level]
    float kp_size;               // (float)(int)(31 * scale)
    int fast_bw, fast_bh;        // TMA box of this level's FAST cells (k_fast.cuh): bytes per patch row (multiple of 16), rows
    // quadtree path code of a candidate = code_x[x] | code_y[y] (k_octree.cuh: the x and y decisions of the 13 splits are independent, the root follows from x);
    // device tables built by the host with the same float arithmetic, null = compute
    const uint32_t* code_x; const uint32_t* code_y; int code_nx, code_ny;
};

// one FAST detection cell (ORBextractor.cc:1089-1157): ROI = [x0,x0+cw) x [y0,y0+ch) in level coords
struct CellDesc {
    short x0, y0, cw, ch;
    short sx, sy;                // j*wCell, i*hCell  (shift applied at :1150-1151)
    short level, cap;            // cap = ceil(zw/2)*ceil(zh/2): max number of strict 3x3 local maxima
    int slot;                    // offset of the cell's candidate slots inside the frame's slot array
    int geo;                     // k_fast.cuh stage A: zone word columns | rows per strip << 8 | strips << 16
    int rcp;                     // 65536 / (zone word columns) + 1:  lane / columns = (lane * rcp) >> 16 for lane < 32
    int pad;
};

// where the pixels of (frame b, level l) live
struct PyrView {
    const uint8_t* l0; long long l0_fstride; int l0_pitch;   // level 0 (internal copy or the caller's device frames)
    uint8_t* pyr; long long pyr_fstride;                     // levels >= 1
};
__device__ __forceinline__ const uint8_t* level_ptr(const PyrView& v, const LevelGeom& g, int level, int b, int& pitch) {
    if (level == 0) { pitch = v.l0_pitch; return v.l0 + (long long)b * v.l0_fstride; }
    pitch = g.pitch; return v.pyr + (long long)b * v.pyr_fstride + g.off;
}

#define ORBX_OVF_SORT 1
#define ORBX_OVF_TREE 2
#define ORBX_OVF_DEPTH 4
