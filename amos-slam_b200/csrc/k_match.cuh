// k_match.cuh -- 256-bit Hamming matching kernels (ORBmatcher / Frame grid / ComputeStereoMatches).
//
// Shape of every windowed matcher on the device:
//   1. k_grid_build      Frame::AssignFeaturesToGrid + PosInGrid (/root/reference/src/Frame.cc:431-461,1007-1030):
//                        bucket keypoints into the 64x48 grid; CSR layout sorted by (cell, index) so that a
//                        column of cells is one contiguous, insertion-ordered range.
//   2. k_window_search   Frame::GetFeaturesInArea (:894-1003) + ORBmatcher::DescriptorDistance
//                        (/root/reference/src/ORBmatcher.cc:1913-1933): one warp per query; candidate indices in
//                        the reference's exact order (cells column-major, insertion order inside), 8x __popc on
//                        the xor of two uint4 halves per candidate; COUNT pass, scan, FILL pass (no caps).
//   3. k_resolve_*       the order-dependent part of each matcher (running exclusions, steals, re-assignment,
//                        30-bin rotation histogram + ComputeThreeMaxima) replayed by ONE warp in the reference's
//                        loop order over the precomputed (index, distance) lists: lanes share a query's
//                        candidates, warp-shuffle reductions give best / second best with first-wins tie-breaks.
// No tensor cores: nothing here is a dense float contraction (popcount + integer compare only).
#pragma once
#include "orbx_common.cuh"

#define GRID_COLS 64
#define GRID_ROWS 48
#define GRID_CELLS (GRID_COLS * GRID_ROWS)
#define M_TH_HIGH 100
#define M_TH_LOW 50
#define M_HISTO 30

struct KpM { float x, y, size, angle, response; int octave, class_id; };

struct FrameDev {                 // device mirror of orbx_frame_view + its grid
    int n; const KpM* keys; const uint8_t* desc; const float* u_right;
    float min_x, min_y, max_x, max_y, gw_inv, gh_inv;
    const float* scale; int nlevels;
    const int* cell_start;        // GRID_CELLS + 1
    const int* entries;           // keypoint indices sorted by (cell, index)
};

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint4* __restrict__ b) {
    const uint4 b0 = __ldg(b), b1 = __ldg(b + 1);
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ORBmatcher::DescriptorDistance on n independent pairs
__global__ void k_descriptor_distance(const uint4* __restrict__ a, const uint4* __restrict__ b, int n, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = hamming256(__ldg(a + 2 * i), __ldg(a + 2 * i + 1), b + 2 * i);
}

// ---- grid build: keys = cell<<20 | index, bitonic sort (one CTA; in shared memory when the frame fits), then cell_start by binary search ----
#define GRID_SMEM_KEYS 8192
__device__ __forceinline__ void grid_build_body(const KpM* __restrict__ keys, int n, float min_x, float min_y, float gw_inv, float gh_inv,
             uint32_t* __restrict__ skeys_g, int* __restrict__ entries, int* __restrict__ cell_start) {
    __shared__ uint32_t skeys_s[GRID_SMEM_KEYS];
    uint32_t* skeys = n <= GRID_SMEM_KEYS ? skeys_s : skeys_g;
    const int tid = threadIdx.x;
    if (n <= GRID_SMEM_KEYS / 2) {
        // Counting sort by cell (the usual case: a frame's ~1000-2000 keypoints): cell sizes by shared-memory atomics, exclusive scan = cell_start, an unordered fill
        // through per-cell cursors, then every entry ranks itself among the handful of entries of its cell by keypoint index (the reference's push_back order).
        // The bitonic sort below (~60 block barriers for 1000 keys) stays for larger sets.
        __shared__ int cnt[GRID_CELLS + 2];
        __shared__ int wsum[32];
        uint32_t* cellof = skeys_s; uint32_t* tmp = skeys_s + GRID_SMEM_KEYS / 2;
        const int lane = tid & 31, warp = tid >> 5;
        for (int c = tid; c < GRID_CELLS + 2; c += 1024) cnt[c] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += 1024) {
            const int px = (int)roundf(__fmul_rn(__fsub_rn(keys[i].x, min_x), gw_inv));
            const int py = (int)roundf(__fmul_rn(__fsub_rn(keys[i].y, min_y), gh_inv));
            const int cell = (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) ? GRID_CELLS : px * GRID_ROWS + py;   // outside the grid: behind every cell
            cellof[i] = (uint32_t)cell;
            atomicAdd(&cnt[cell], 1);
        }
        __syncthreads();
        // exclusive scan of cnt[0 .. GRID_CELLS]: every thread owns a run of cells
        constexpr int PER = (GRID_CELLS + 1 + 1023) / 1024;
        const int c0 = min(tid * PER, GRID_CELLS + 1), c1 = min(c0 + PER, GRID_CELLS + 1);
        int mine = 0;
        for (int c = c0; c < c1; ++c) mine += cnt[c];
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int run = incl - mine;
        for (int w = 0; w < warp; ++w) run += wsum[w];
        for (int c = c0; c < c1; ++c) { const int v = cnt[c]; cnt[c] = run; cell_start[c] = run; run += v; }   // cnt becomes the fill cursor
        __syncthreads();
        for (int i = tid; i < n; i += 1024) tmp[atomicAdd(&cnt[cellof[i]], 1)] = (uint32_t)i;
        __syncthreads();
        for (int p = tid; p < n; p += 1024) {
            const uint32_t i = tmp[p];
            const int c = (int)cellof[i];
            const int e1 = cnt[c];                                             // end of the cell's run (cursor after the fill)
            int s0 = p; while (s0 > 0 && cellof[tmp[s0 - 1]] == (uint32_t)c) --s0;   // its start (runs are a few entries long)
            int r = 0;
            for (int q = s0; q < e1; ++q) r += tmp[q] < i;
            entries[s0 + r] = (int)i;
        }
        return;
    }
    for (int i = tid; i < n; i += 1024) {
        // posX = round((kp.pt.x - mnMinX) * mfGridElementWidthInv)   (roundf: half away from zero)
        const int px = (int)roundf(__fmul_rn(__fsub_rn(keys[i].x, min_x), gw_inv));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(keys[i].y, min_y), gh_inv));
        const uint32_t cell = (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) ? 0xFFFu : (uint32_t)(px * GRID_ROWS + py);
        skeys[i] = (cell << 20) | (uint32_t)i;
    }
    __syncthreads();
    for (int k = 2; (k >> 1) < n; k <<= 1) {
        for (int i = tid; i < n; i += 1024) { const int j = i ^ (k - 1); if (j > i && j < n) { uint32_t a = skeys[i], c = skeys[j]; if (a > c) { skeys[i] = c; skeys[j] = a; } } }
        __syncthreads();
        for (int s = k >> 2; s > 0; s >>= 1) {
            for (int i = tid; i < n; i += 1024) { const int j = i ^ s; if (j > i && j < n) { uint32_t a = skeys[i], c = skeys[j]; if (a > c) { skeys[i] = c; skeys[j] = a; } } }
            __syncthreads();
        }
    }
    for (int i = tid; i < n; i += 1024) entries[i] = (int)(skeys[i] & 0xFFFFFu);
    for (int c = tid; c <= GRID_CELLS; c += 1024) {
        const uint32_t T = (uint32_t)c << 20;
        int lo = 0, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (skeys[mid] < T) lo = mid + 1; else hi = mid; }
        cell_start[c] = lo;
    }
}
__global__ void __launch_bounds__(1024)
k_grid_build(const KpM* __restrict__ keys, int n, float min_x, float min_y, float gw_inv, float gh_inv,
             uint32_t* __restrict__ skeys_g, int* __restrict__ entries, int* __restrict__ cell_start) {
    grid_build_body(keys, n, min_x, min_y, gw_inv, gh_inv, skeys_g, entries, cell_start);
}

// ---- queries ----
#define MODE_INIT 0
#define MODE_PROJ_FRAME 1
#define MODE_PROJ_POINTS 2
#define MODE_AREA 3               // plain Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel) windows
struct QueryParams {
    int mode, nq;
    // INIT: F1 keys / desc, vbPrevMatched, windowSize
    const KpM* q_keys; const uint8_t* q_desc; const float* q_xy; float window;
    // PROJ_FRAME: proj_invz, last_octave, valid, th, forward/backward, mbf        (q_xy = proj_uv, q_desc = mp_desc)
    const float* q_invz; const int* q_octave; const uint8_t* q_valid; float th; int forward, backward; float mbf; int no_ur;   // no_ur: the KeyFrame form has no uRight test
    // PROJ_POINTS: track_ur, track_level, track_view_cos, th                      (q_xy = track_uv, q_desc = mp_desc)
    const float* q_ur; const float* q_viewcos;
    // Fuse: chi2 != 0 adds the reprojection gate of ORBmatcher::Fuse (:1097-1137) to PROJ_FRAME: q_ur = the point's right-image coordinate,
    // q_invsigma2 = mvInvLevelSigma2 by level
    int chi2; const float* q_invsigma2;
    // AREA: q_xy = (x, y), q_r, q_minlevel, q_maxlevel
    const float* q_r; const int* q_minlevel; const int* q_maxlevel;
};

// per-query window of GetFeaturesInArea and the static candidate filters; returns false if the query is skipped
__device__ __forceinline__ bool query_window(const QueryParams& P, const FrameDev& F, int q, float& x, float& y, float& r,
                                             int& minLevel, int& maxLevel, float& ur, bool& use_ur) {
    use_ur = false; ur = 0.f;
    if (P.mode == MODE_INIT) {
        if (P.q_keys[q].octave > 0) return false;                                   // ORBmatcher.cc:537
        x = P.q_xy[2 * q]; y = P.q_xy[2 * q + 1]; r = P.window; minLevel = 0; maxLevel = 0;
        return true;
    }
    if (P.mode == MODE_PROJ_FRAME) {
        if (!P.q_valid[q]) return false;
        const float u = P.q_xy[2 * q], v = P.q_xy[2 * q + 1], invz = P.q_invz[q];
        if (invz < 0.f) return false;                                               // :1612
        if (u < F.min_x || u > F.max_x) return false;                               // :1620
        if (v < F.min_y || v > F.max_y) return false;
        const int oct = P.q_octave[q];
        x = u; y = v; r = __fmul_rn(P.th, F.scale[oct]);                            // :1629
        if (P.forward == 2) { minLevel = oct - 1; maxLevel = oct; }                 // KeyFrame x map points: kpLevel in [nPredictedLevel - 1, nPredictedLevel] (:462-463)
        else if (P.forward) { minLevel = oct; maxLevel = -1; }                      // :1637-1642
        else if (P.backward) { minLevel = 0; maxLevel = oct; }
        else { minLevel = oct - 1; maxLevel = oct + 1; }
        ur = __fsub_rn(u, __fmul_rn(P.mbf, invz)); use_ur = !P.no_ur;               // :1665
        if (P.chi2) ur = P.q_ur[q];
        return true;
    }
    if (P.mode == MODE_AREA) {
        x = P.q_xy[2 * q]; y = P.q_xy[2 * q + 1]; r = P.q_r[q]; minLevel = P.q_minlevel[q]; maxLevel = P.q_maxlevel[q];
        return true;
    }
    // MODE_PROJ_POINTS  (:89-103)
    const int lvl = P.q_octave[q];
    float rr = ((double)P.q_viewcos[q] > 0.998) ? 2.5f : 4.0f;                      // RadiusByViewingCos :178-185
    if (P.th != 1.0f) rr = __fmul_rn(rr, P.th);
    x = P.q_xy[2 * q]; y = P.q_xy[2 * q + 1]; r = __fmul_rn(rr, F.scale[lvl]);
    minLevel = lvl - 1; maxLevel = lvl;
    ur = P.q_ur[q]; use_ur = true;
    return true;
}

// one warp per query.  FILL = false: counts[q] only.  FILL = true: cand[offsets[q] + k] = dist<<20 | index.
template <bool FILL>
__device__ __forceinline__ void window_search_body(const QueryParams& P, const FrameDev& F, int* __restrict__ counts, const int* __restrict__ offsets, uint32_t* __restrict__ cand, int cand_cap,
                uint2* __restrict__ pre_best /* FILL: per query (best, second) key = dist<<20 | position in list, ignoring running exclusions */, int q, int lane) {
    if (q >= P.nq) return;
    if (FILL && offsets[P.nq] > cand_cap) return;                   // lists do not fit: the host grows the arena and repeats the call
    float x, y, r, ur; int minLevel, maxLevel; bool use_ur;
    if (!query_window(P, F, q, x, y, r, minLevel, maxLevel, ur, use_ur)) { if (!FILL && lane == 0) counts[q] = 0; return; }
    // cell range (Frame.cc:913-939), float arithmetic in the reference's order
    const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, F.min_x), r), F.gw_inv)));
    const int cx1 = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, F.min_x), r), F.gw_inv)));
    const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, F.min_y), r), F.gh_inv)));
    const int cy1 = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, F.min_y), r), F.gh_inv)));
    int n = 0;
    uint32_t pk1 = 0xFFFFFFFFu, pk2 = 0xFFFFFFFFu;
    const uint32_t dlimit = P.mode == MODE_INIT ? 0xFFFu : 256u;    // the projection matchers start from bestDist = 256 (strict <)
    if (!(cx0 >= GRID_COLS || cx1 < 0 || cy0 >= GRID_ROWS || cy1 < 0)) {
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        uint4 d0 = make_uint4(0, 0, 0, 0), d1 = d0;
        uint32_t* out = nullptr;
        if (FILL) {
            if (P.mode != MODE_AREA) {
                const uint4* qd = reinterpret_cast<const uint4*>(P.q_desc) + 2 * q;
                d0 = __ldg(qd); d1 = __ldg(qd + 1);
            }
            out = cand + offsets[q];
        }
        // The window covers cx1 - cx0 + 1 grid columns; the features of a column's cells cy0 .. cy1 are one contiguous run of F.entries (column-major grid).  Walking the
        // columns one by one left most lanes idle (a 100-px window holds ~6 features per column: 21 sparse steps per query); instead every lane takes one column's
        // run, a prefix sum concatenates the runs, and the warp walks the concatenation 32 entries at a time -- the same order, a quarter of the steps.
        for (int cg = cx0; cg <= cx1; cg += 32) {
            const int ixl = cg + lane;
            int run0 = 0, len = 0;
            if (ixl <= cx1) { run0 = F.cell_start[ixl * GRID_ROWS + cy0]; len = F.cell_start[ixl * GRID_ROWS + cy1 + 1] - run0; }
            int incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            for (int t0 = 0; t0 < total; t0 += 32) {
                const int pp = t0 + lane;
                int col = 0;                                                     // number of columns whose runs end at or before position pp
#pragma unroll
                for (int st = 16; st > 0; st >>= 1) { const int v = __shfl_sync(0xffffffffu, incl, col + st - 1); if (v <= pp) col += st; }
                const int cbase = __shfl_sync(0xffffffffu, run0, col), cprev = __shfl_sync(0xffffffffu, incl - len, col);
                const int ee = cbase + (pp - cprev);
                bool keep = false; int j = -1;
                if (pp < total) {
                    j = F.entries[ee];
                    const KpM kp = F.keys[j];
                    keep = true;
                    if (bCheckLevels) {
                        if (kp.octave < minLevel) keep = false;
                        if (maxLevel >= 0 && kp.octave > maxLevel) keep = false;
                    }
                    if (keep && !(fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r)) keep = false;    // :986-991
                    if (keep && P.chi2) {                                                                              // Fuse: chi-square gate on the reprojection error
                        const float ex = __fsub_rn(x, kp.x), ey = __fsub_rn(y, kp.y);
                        const float urj = F.u_right ? F.u_right[j] : -1.f;
                        const float inv = P.q_invsigma2[kp.octave];
                        if (urj >= 0.f) {
                            const float er = __fsub_rn(ur, urj);
                            const float e2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(er, er));
                            if ((double)__fmul_rn(e2, inv) > 7.8) keep = false;                                          // :1111-1113
                        } else {
                            const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                            if ((double)__fmul_rn(e2, inv) > 5.99) keep = false;                                         // :1123-1125
                        }
                    }
                    if (keep && use_ur && F.u_right) {
                        const float urj = F.u_right[j];
                        if (urj > 0.f && fabsf(__fsub_rn(ur, urj)) > r) keep = false;                                // :1662-1669 / :129-139
                    }
                }
                const uint32_t m = __ballot_sync(0xffffffffu, keep);
                if (FILL && keep) {
                    const int d = P.mode == MODE_AREA ? 0 : hamming256(d0, d1, reinterpret_cast<const uint4*>(F.desc) + 2 * j);
                    const int pos = n + __popc(m & ((1u << lane) - 1u));
                    out[pos] = ((uint32_t)d << 20) | (uint32_t)j;
                    if ((uint32_t)d < dlimit) {
                        const uint32_t key = ((uint32_t)d << 20) | (uint32_t)pos;
                        if (key < pk1) { pk2 = pk1; pk1 = key; } else if (key < pk2) pk2 = key;
                    }
                }
                n += __popc(m);
            }
        }
    }
    if (!FILL && lane == 0) counts[q] = n;
    if (FILL) {
        const uint32_t g1 = __reduce_min_sync(0xffffffffu, pk1);
        const uint32_t g2 = __reduce_min_sync(0xffffffffu, (pk1 == g1 && g1 != 0xFFFFFFFFu) ? pk2 : pk1);
        if (lane == 0) pre_best[q] = make_uint2(g1, g2);
    }
}
template <bool FILL>
__global__ void __launch_bounds__(128)
k_window_search(QueryParams P, FrameDev F, int* __restrict__ counts, const int* __restrict__ offsets, uint32_t* __restrict__ cand, int cand_cap, uint2* __restrict__ pre_best) {
    window_search_body<FILL>(P, F, counts, offsets, cand, cand_cap, pre_best, blockIdx.x * 4 + (threadIdx.x >> 5), threadIdx.x & 31);
}

// exclusive scan of n counts by one CTA; total written to offsets[n]
__device__ __forceinline__ void scan_counts_body(const int* __restrict__ counts, int n, int* __restrict__ offsets) {
    __shared__ int wsum[32];
    __shared__ int carry_sm;
    const int tid = threadIdx.x;
    if (tid == 0) carry_sm = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const int v = i < n ? counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += t; }
        if ((tid & 31) == 31) wsum[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) { int w = wsum[tid], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, wi, o); if (tid >= o) wi += t; }
            wsum[tid] = wi - w; if (tid == 31) wsum[31] = wi - w; }
        __syncthreads();
        const int carry = carry_sm;
        if (i < n) offsets[i] = carry + wsum[tid >> 5] + incl - v;
        __syncthreads();
        if (tid == 1023) carry_sm = carry + wsum[31] + incl;
        __syncthreads();
    }
    if (tid == 0) offsets[n] = carry_sm;
}
__global__ void __launch_bounds__(1024) k_scan_counts(const int* __restrict__ counts, int n, int* __restrict__ offsets) { scan_counts_body(counts, n, offsets); }

// independent best match of every query (no running state): the FILL pass already reduced each list to its best key
__global__ void k_best_extract(int nq, const int* __restrict__ offsets, const uint32_t* __restrict__ cand, const uint2* __restrict__ pre_best, int cand_cap, int max_dist,
                               int* __restrict__ best_idx) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int out = -1;
    if (offsets[nq] <= cand_cap && offsets[q + 1] > offsets[q]) {                // skipped queries have an empty list and no pre_best entry
        const uint32_t k = pre_best[q].x;
        if (k != 0xFFFFFFFFu && (int)(k >> 20) <= max_dist) out = (int)(cand[offsets[q] + (int)(k & 0xFFFFFu)] & 0xFFFFFu);
    }
    best_idx[q] = out;
}
// SearchBySim3's mutual-consistency check (ORBmatcher.cc:1535-1550)
__global__ void k_mutual_check(int n1, const int* __restrict__ match1, const int* __restrict__ match2, int* __restrict__ match12, int* __restrict__ nfound) {
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    int ok = 0, idx2 = -1;
    if (i1 < n1) { idx2 = match1[i1]; ok = idx2 >= 0 && match2[idx2] == i1; match12[i1] = ok ? idx2 : -1; }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(nfound, __popc(m));
}

// two smallest candidate keys (dist<<20 | position-in-list) of a query under a per-candidate predicate
struct Best2 { uint32_t k1, k2; };
template <typename Pred>
__device__ __forceinline__ Best2 warp_best2(const uint32_t* __restrict__ cand, int cnt, int lane, Pred ok) {
    uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
    for (int p = lane; p < cnt; p += 32) {
        const uint32_t c = cand[p];
        if (!ok(c)) continue;
        const uint32_t key = (c & 0xFFF00000u) | (uint32_t)p;       // distance, then order of appearance
        if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
    }
    // two hardware warp reductions (REDUX.MIN): keys are distinct per lane except for the 0xFFFFFFFF sentinel, so the
    // runner-up is the minimum over lanes of (the winner lane's second key, every other lane's first key)
    const uint32_t g1 = __reduce_min_sync(0xffffffffu, k1);
    const uint32_t g2 = __reduce_min_sync(0xffffffffu, (k1 == g1 && g1 != 0xFFFFFFFFu) ? k2 : k1);
    k1 = g1; k2 = g2;
    Best2 b; b.k1 = k1; b.k2 = k2; return b;
}

__device__ __forceinline__ int rot_bin(float a1, float a2) {        // ORBmatcher.cc:597-602
    const float factor = (float)M_HISTO / 360.0f;
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, factor));
    if (bin == M_HISTO) bin = 0;
    return bin;
}
// ORBmatcher::ComputeThreeMaxima   ORBmatcher.cc:1866-1908
__device__ __forceinline__ void three_maxima(const int* hist, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0; ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < M_HISTO; i++) {
        const int s = hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
}

// -------------------------------------------------------------------------------------------------
// Sequential "resolve" passes.  SearchForInitialization and both SearchByProjection variants are order-dependent
// (running exclusions, steals, occupied flags), so one warp replays the reference's loop; what makes that fast is
// that nothing on the serial chain touches global memory: the whole block (RESOLVE_THREADS) first stages the
// candidate lists, a compacted list of the queries that have candidates, and the per-kernel state / lookup arrays
// in shared memory.  If the lists do not fit the shared-memory budget the kernels fall back to the global arrays.
// -------------------------------------------------------------------------------------------------
#define RESOLVE_THREADS 256
struct StagedLists {
    bool staged;            // lists / act / aoff are valid shared-memory copies
    const uint32_t* lists;  // concatenated candidate lists
    const uint32_t* act;    // query index of the k-th query with a non-empty list
    const uint32_t* aoff;   // start of its list; aoff[niter] = total
    const uint2* pre;       // staged: (best, second) of the k-th active query; else: indexed by query
    int niter;              // staged: number of active queries; else: n
    uint32_t* free_words;   // first shared word after the staged lists
};
// words needed besides the lists themselves: act[n] + aoff[n+1] + pre[2n]
__device__ __forceinline__ StagedLists stage_lists(uint32_t* sm, int sm_words, int extra_words, int n, const int* __restrict__ counts,
                                                   const int* __restrict__ offsets, const uint32_t* __restrict__ cand, const uint2* __restrict__ pre_best) {
    __shared__ int sl_wsum[RESOLVE_THREADS / 32];
    __shared__ int sl_nact;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = offsets[n];
    StagedLists r;
    r.staged = 4 * n + 2 + total + extra_words <= sm_words;
    uint32_t* act = sm; uint32_t* aoff = act + n; uint2* pre = reinterpret_cast<uint2*>(aoff + n + 2);     /* act[n] + aoff[n+1] is an odd number of words: one pad word keeps uint2 aligned */ uint32_t* lists = reinterpret_cast<uint32_t*>(pre + n);
    r.lists = r.staged ? lists : cand; r.act = act; r.aoff = aoff; r.pre = r.staged ? pre : pre_best; r.niter = n; r.free_words = r.staged ? lists + total : sm;
    if (!r.staged) return r;
    for (int i = tid; i < total; i += RESOLVE_THREADS) lists[i] = cand[i];
    int nact = 0;
    for (int base = 0; base < n; base += RESOLVE_THREADS) {               // ordered compaction of the queries with candidates
        const int i = base + tid;
        const int has = (i < n && counts[i] > 0) ? 1 : 0;
        const uint32_t m = __ballot_sync(0xffffffffu, has);
        if (lane == 0) sl_wsum[warp] = __popc(m);
        __syncthreads();
        int before = 0, chunk = 0;
#pragma unroll
        for (int w = 0; w < RESOLVE_THREADS / 32; ++w) { const int c = sl_wsum[w]; if (w < warp) before += c; chunk += c; }
        if (has) { const int k = nact + before + __popc(m & ((1u << lane) - 1u)); act[k] = (uint32_t)i; aoff[k] = (uint32_t)offsets[i]; pre[k] = pre_best[i]; }
        nact += chunk;
        __syncthreads();
    }
    if (tid == 0) { aoff[nact] = (uint32_t)total; sl_nact = nact; }
    __syncthreads();
    r.niter = sl_nact;
    return r;
}
// iteration `it` of the replay loop -> (query index, list pointer, list length); cnt == 0 means "skip"
#define RESOLVE_QUERY(SL, it, qi, c, cnt, pb) \
    int qi, cnt; const uint32_t* c; uint2 pb; \
    if ((SL).staged) { qi = (int)(SL).act[it]; const int _lo = (int)(SL).aoff[it]; cnt = (int)(SL).aoff[(it) + 1] - _lo; c = (SL).lists + _lo; pb = (SL).pre[it]; } \
    else { qi = (it); cnt = counts[qi]; c = (SL).lists + (cnt ? offsets[qi] : 0); pb = cnt ? (SL).pre[qi] : make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu); }

// -------------------------------------------------------------------------------------------------
// Order-preserving replay, 32 queries per step.  Every lane takes one query (in loop order) and evaluates it against the state as
// it stands: if its precomputed best pair is still selectable the acceptance test is O(1); otherwise the lane rescans its own
// (short) list under the running exclusions.  A lane may be committed together with its predecessors of the step unless one of
// them claims a feature that the lane's decision rests on (its best / second-best feature) -- the running state only ever
// removes candidates, so nothing else can change the outcome.  So: commit the longest conflict-free prefix in lane order and
// re-evaluate from the first stale lane.  The result equals the reference's one-by-one loop.
//   Ops:  bool clean(uint32_t cand_entry)                      candidate still selectable under the running state
//         uint32_t dlimit                                      candidates at this distance or above never win (256 / none)
//         bool decide(int qi, const uint32_t* c, Best2 b, int& j)   acceptance test (reads only static data besides b)
//         bool commit(int qi, int j, Best2 b, int rank)        state update by one lane; returns true if it displaced a match
//         void post(int naccept, int ndisplaced)               uniform counters
// -------------------------------------------------------------------------------------------------
#define NOJ 0xFFFFFFFFu
template <bool USE_SECOND, class Ops>
__device__ __forceinline__ void replay_queries(const StagedLists& SL, const int* __restrict__ counts, const int* __restrict__ offsets, int lane, Ops& ops) {
    const uint32_t lt = (1u << lane) - 1u;
    for (int base = 0; base < SL.niter; base += 32) {
        const int it = base + lane;
        const bool act = it < SL.niter;
        int qi = 0, cnt = 0, lo = 0; uint2 pb = make_uint2(NOJ, NOJ);
        if (act) {
            if (SL.staged) { qi = (int)SL.act[it]; lo = (int)SL.aoff[it]; cnt = (int)SL.aoff[it + 1] - lo; pb = SL.pre[it]; }
            else { qi = it; cnt = counts[qi]; if (cnt) { lo = offsets[qi]; pb = SL.pre[qi]; } }
        }
        const uint32_t* c = SL.lists + lo;
        const int nb = min(32, SL.niter - base);
        Best2 b; b.k1 = pb.x; b.k2 = pb.y;                     // best pair under the running state (starts as the unconstrained pair)
        int start = 0;
        while (start < nb) {
            bool accept = false; int j = -1; uint32_t r1 = NOJ, r2 = NOJ;
            if (act && lane >= start && cnt > 0) {
                bool ok = true;
                if (b.k1 != NOJ) ok = ops.clean(c[b.k1 & 0xFFFFFu]);
                if (ok && USE_SECOND && b.k2 != NOJ) ok = ops.clean(c[b.k2 & 0xFFFFFu]);
                if (!ok) {                                         // one of the pair has been taken: this lane rescans its own list
                    uint32_t k1 = NOJ, k2 = NOJ;
                    for (int p = 0; p < cnt; ++p) {
                        const uint32_t v = c[p];
                        if ((v >> 20) >= ops.dlimit || !ops.clean(v)) continue;
                        const uint32_t key = (v & 0xFFF00000u) | (uint32_t)p;
                        if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
                    }
                    b.k1 = k1; b.k2 = k2;
                }
                if (b.k1 != NOJ) r1 = c[b.k1 & 0xFFFFFu] & 0xFFFFFu;
                if (USE_SECOND && b.k2 != NOJ) r2 = c[b.k2 & 0xFFFFFu] & 0xFFFFFu;
                accept = ops.decide(qi, c, b, j);
            }
            // does a predecessor of this step (lanes start .. lane-1) claim a feature this lane's decision rests on?
            const uint32_t myclaim = accept ? (uint32_t)j : NOJ;
            bool stale = false;
#pragma unroll
            for (int K = 0; K < 31; ++K) {
                const uint32_t v = __shfl_sync(0xffffffffu, myclaim, K);
                stale |= (K >= start) & (K < lane) & (v != NOJ) & ((v == r1) | (v == r2));
            }
            const uint32_t stop = __ballot_sync(0xffffffffu, lane >= start && lane < nb && stale);
            const int L = stop ? __ffs(stop) - 1 : nb;             // first lane that cannot be committed with its predecessors (never `start` itself)
            const bool mine = accept && lane >= start && lane < L;
            const uint32_t am = __ballot_sync(0xffffffffu, mine);
            bool displaced = false;
            if (mine) displaced = ops.commit(qi, j, b, __popc(am & lt));
            const uint32_t dm = __ballot_sync(0xffffffffu, displaced);
            ops.post(__popc(am), __popc(dm));
            __syncwarp();
            start = L;
        }
    }
}

// ---- SearchForInitialization, sequential part (ORBmatcher.cc:532-640) ----
struct InitOps {
    int* md; int* s21; int* m12; int* bin_of; int* hist; const float* ang1; const float* ang2; const KpM* k1s; const KpM* k2s;
    float nnratio; int checkOri; bool staged; int nmatches;
    __device__ __forceinline__ bool clean(uint32_t v) const { return !(md[v & 0xFFFFFu] <= (int)(v >> 20)); }                   // :561
    static constexpr uint32_t dlimit = 0xFFFu;                                                                                   // no distance cut in the list scan
    __device__ __forceinline__ bool decide(int, const uint32_t* c, Best2 b, int& j) const {
        if (b.k1 == NOJ) return false;
        const int bestDist = (int)(b.k1 >> 20);
        const int bestDist2 = b.k2 == NOJ ? INT_MAX : (int)(b.k2 >> 20);
        if (!(bestDist <= M_TH_LOW && (float)bestDist < __fmul_rn((float)bestDist2, nnratio))) return false;                  // :577-579
        j = (int)(c[b.k1 & 0xFFFFFu] & 0xFFFFFu);
        return true;
    }
    __device__ __forceinline__ bool commit(int i1, int j, Best2 b, int) {
        const int old = s21[j];
        if (old >= 0) m12[old] = -1;                                                                                           // :583-587
        m12[i1] = j; s21[j] = i1; md[j] = (int)(b.k1 >> 20);
        if (checkOri) { const int bin = staged ? rot_bin(ang1[i1], ang2[j]) : rot_bin(k1s[i1].angle, k2s[j].angle); bin_of[i1] = bin; atomicAdd(&hist[bin], 1); }
        return old >= 0;
    }
    __device__ __forceinline__ void post(int na, int nd) { nmatches += na - nd; }
};

__device__ __forceinline__ void resolve_init_body(int n1, int n2, const KpM* __restrict__ k1s, const KpM* __restrict__ k2s, const int* __restrict__ counts, const int* __restrict__ offsets,
               const uint32_t* __restrict__ cand, const uint2* __restrict__ pre_best, int cand_cap, float nnratio, int checkOri, int smem_words,
               int* __restrict__ matchedDist /*n2*/, int* __restrict__ m21 /*n2*/, int* __restrict__ m12 /*n1*/, int* __restrict__ bin_of /*n1*/,
               float* __restrict__ prev_xy, int* __restrict__ nmatches_out) {
    extern __shared__ __align__(16) uint32_t rs_sm[];
    __shared__ int hist[M_HISTO];
    const int tid = threadIdx.x, lane = tid & 31;
    if (offsets[n1] > cand_cap) return;                            // candidate lists were not written (see k_window_search)
    const StagedLists SL = stage_lists(rs_sm, smem_words, 3 * n2 + n1, n1, counts, offsets, cand, pre_best);
    const bool staged = SL.staged;
    // extra shared arrays: md[n2] | s21[n2] | ang2[n2] | ang1[n1]
    int* md = staged ? reinterpret_cast<int*>(SL.free_words) : matchedDist;
    int* s21 = staged ? md + n2 : m21;
    float* ang2 = reinterpret_cast<float*>(SL.free_words) + 2 * n2;
    float* ang1 = ang2 + n2;
    for (int i = tid; i < n2; i += RESOLVE_THREADS) { md[i] = INT_MAX; s21[i] = -1; if (staged) ang2[i] = k2s[i].angle; }
    for (int i = tid; i < n1; i += RESOLVE_THREADS) { m12[i] = -1; bin_of[i] = -1; if (staged) ang1[i] = k1s[i].angle; }
    if (tid < M_HISTO) hist[tid] = 0;
    __syncthreads();
    if (tid >= 32) return;
    InitOps ops{md, s21, m12, bin_of, hist, ang1, ang2, k1s, k2s, nnratio, checkOri, staged, 0};
    replay_queries<true>(SL, counts, offsets, lane, ops);
    int nmatches = ops.nmatches;
    __syncwarp();
    if (checkOri) {
        int ind1, ind2, ind3;
        three_maxima(hist, ind1, ind2, ind3);
        int removed = 0;
        for (int i = lane; i < n1; i += 32) {
            const int bin = bin_of[i];
            if (bin >= 0 && bin != ind1 && bin != ind2 && bin != ind3 && m12[i] >= 0) { m12[i] = -1; ++removed; }   // :620-633
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        nmatches -= removed;
    }
    __syncwarp();
    for (int i = lane; i < n1; i += 32) { const int j = m12[i]; if (j >= 0) { prev_xy[2 * i] = k2s[j].x; prev_xy[2 * i + 1] = k2s[j].y; } }   // :638-640
    if (lane == 0) *nmatches_out = nmatches;
}
__global__ void __launch_bounds__(RESOLVE_THREADS)
k_resolve_init(int n1, int n2, const KpM* __restrict__ k1s, const KpM* __restrict__ k2s, const int* __restrict__ counts, const int* __restrict__ offsets,
               const uint32_t* __restrict__ cand, const uint2* __restrict__ pre_best, int cand_cap, float nnratio, int checkOri, int smem_words,
               int* __restrict__ matchedDist, int* __restrict__ m21, int* __restrict__ m12, int* __restrict__ bin_of, float* __restrict__ prev_xy, int* __restrict__ nmatches_out) {
    resolve_init_body(n1, n2, k1s, k2s, counts, offsets, cand, pre_best, cand_cap, nnratio, checkOri, smem_words, matchedDist, m21, m12, bin_of, prev_xy, nmatches_out);
}

// ---- SearchForInitialization for P independent frame pairs per launch (the shard unit of config C2): the same four steps,
// one grid-build CTA / scan CTA / resolve CTA per pair and one warp per (pair, query) in between.  Every pair owns a fixed slice of the
// candidate arena; a pair whose lists do not fit reports its total and the host repeats the call with a larger slice. ----
struct InitPair {
    FrameDev F2; uint32_t* sort_keys;            // sort_keys == nullptr: F2's grid is already built (device-resident frame)
    const KpM* k1; const uint8_t* d1; float* prev; int n1;
    int* counts; int* offsets; uint2* pre; uint32_t* cand;
    int* md; int* m21; int* m12; int* binof; int* nmatches; int* total;
};
__global__ void __launch_bounds__(1024) k_grid_build_pairs(const InitPair* __restrict__ pairs) {
    const InitPair& p = pairs[blockIdx.x];
    if (!p.sort_keys) return;
    grid_build_body(p.F2.keys, p.F2.n, p.F2.min_x, p.F2.min_y, p.F2.gw_inv, p.F2.gh_inv, p.sort_keys, const_cast<int*>(p.F2.entries), const_cast<int*>(p.F2.cell_start));
}
template <bool FILL>
__global__ void __launch_bounds__(128) k_window_search_pairs(const InitPair* __restrict__ pairs, float window, int cand_cap) {
    const InitPair& p = pairs[blockIdx.y];
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (q >= p.n1) return;
    QueryParams P;
    P.mode = MODE_INIT; P.nq = p.n1; P.q_keys = p.k1; P.q_desc = p.d1; P.q_xy = p.prev; P.window = window;
    P.q_invz = nullptr; P.q_octave = nullptr; P.q_valid = nullptr; P.th = 0.f; P.forward = P.backward = 0; P.mbf = 0.f; P.no_ur = 0; P.q_ur = nullptr; P.q_viewcos = nullptr;
    P.chi2 = 0; P.q_invsigma2 = nullptr; P.q_r = nullptr; P.q_minlevel = nullptr; P.q_maxlevel = nullptr;
    window_search_body<FILL>(P, p.F2, p.counts, p.offsets, p.cand, cand_cap, p.pre, q, threadIdx.x & 31);
}
__global__ void __launch_bounds__(1024) k_scan_counts_pairs(const InitPair* __restrict__ pairs) {
    const InitPair& p = pairs[blockIdx.x];
    scan_counts_body(p.counts, p.n1, p.offsets);
    if (threadIdx.x == 0) *p.total = p.offsets[p.n1];          // written by this thread inside scan_counts_body
}
__global__ void __launch_bounds__(RESOLVE_THREADS) k_resolve_init_pairs(const InitPair* __restrict__ pairs, int cand_cap, float nnratio, int checkOri, int smem_words) {
    const InitPair& p = pairs[blockIdx.x];
    resolve_init_body(p.n1, p.F2.n, p.k1, p.F2.keys, p.counts, p.offsets, p.cand, p.pre, cand_cap, nnratio, checkOri, smem_words, p.md, p.m21, p.m12, p.binof, p.prev, p.nmatches);
}

// ---- SearchByProjection(Frame, Frame), sequential part (ORBmatcher.cc:1595-1725) ----
struct ProjFrameOps {
    int* occ; int* cur_match; int* pushes; int* hist; const int* obs; const uint8_t* mp_observed; const float* angl; const float* angc;
    const float* last_angle; const KpM* cur_keys; int checkOri; bool staged; int nmatches, npush; int max_dist;
    __device__ __forceinline__ bool clean(uint32_t v) const { return !occ[v & 0xFFFFFu]; }                                        // :1658-1660
    static constexpr uint32_t dlimit = 256u;                                                                                     // bestDist starts at 256, strict <
    __device__ __forceinline__ bool decide(int, const uint32_t* c, Best2 b, int& j) const {
        if (b.k1 == NOJ || (int)(b.k1 >> 20) > max_dist) return false;                                                           // :1683 (TH_HIGH) / :1820 (ORBdist)
        j = (int)(c[b.k1 & 0xFFFFFu] & 0xFFFFFu);
        return true;
    }
    __device__ __forceinline__ bool commit(int i, int j, Best2, int rank) {
        cur_match[j] = i;
        occ[j] = staged ? obs[i] : (mp_observed ? (mp_observed[i] != 0) : 0);
        if (checkOri) {
            const int bin = staged ? rot_bin(angl[i], angc[j]) : rot_bin(last_angle[i], cur_keys[j].angle);
            atomicAdd(&hist[bin], 1); pushes[2 * (npush + rank)] = bin; pushes[2 * (npush + rank) + 1] = j;
        }
        return false;
    }
    __device__ __forceinline__ void post(int na, int) { nmatches += na; npush += na; }
};

__global__ void __launch_bounds__(RESOLVE_THREADS)
k_resolve_proj_frame(int n_last, int n_cur, const KpM* __restrict__ cur_keys, const float* __restrict__ last_angle, const uint8_t* __restrict__ mp_observed,
                     const uint8_t* __restrict__ cur_occupied, const int* __restrict__ counts, const int* __restrict__ offsets, const uint32_t* __restrict__ cand,
                     const uint2* __restrict__ pre_best, int cand_cap, int checkOri, int max_dist, int smem_words, int* __restrict__ occupied_g /*n_cur*/, int* __restrict__ cur_match /*n_cur*/,
                     int* __restrict__ pushes /*2*n_last*/, int* __restrict__ nmatches_out) {
    extern __shared__ __align__(16) uint32_t rs_sm[];
    __shared__ int hist[M_HISTO];
    const int tid = threadIdx.x, lane = tid & 31;
    if (offsets[n_last] > cand_cap) return;
    // extra shared arrays: occ[n_cur] (ints) | angc[n_cur] | angl[n_last] | obs[n_last] (ints)
    const StagedLists SL = stage_lists(rs_sm, smem_words, 2 * n_cur + 2 * n_last, n_last, counts, offsets, cand, pre_best);
    const bool staged = SL.staged;
    int* occ = staged ? reinterpret_cast<int*>(SL.free_words) : occupied_g;
    float* angc = reinterpret_cast<float*>(SL.free_words) + n_cur;
    float* angl = angc + n_cur;
    int* obs = reinterpret_cast<int*>(angl + n_last);
    for (int j = tid; j < n_cur; j += RESOLVE_THREADS) {
        cur_match[j] = -1;
        occ[j] = cur_occupied ? (cur_occupied[j] != 0) : 0;
        if (staged) angc[j] = cur_keys[j].angle;
    }
    if (staged) for (int i = tid; i < n_last; i += RESOLVE_THREADS) { angl[i] = last_angle[i]; obs[i] = mp_observed ? (mp_observed[i] != 0) : 0; }
    if (tid < M_HISTO) hist[tid] = 0;
    __syncthreads();
    if (tid >= 32) return;
    ProjFrameOps ops{occ, cur_match, pushes, hist, obs, mp_observed, angl, angc, last_angle, cur_keys, checkOri, staged, 0, 0, max_dist};
    replay_queries<false>(SL, counts, offsets, lane, ops);
    int nmatches = ops.nmatches;
    const int npush = ops.npush;
    __syncwarp();
    if (checkOri) {
        int ind1, ind2, ind3;
        three_maxima(hist, ind1, ind2, ind3);
        int removed = 0;
        for (int e = lane; e < npush; e += 32) {
            const int bin = pushes[2 * e];
            if (bin != ind1 && bin != ind2 && bin != ind3) { cur_match[pushes[2 * e + 1]] = -2; ++removed; }   // :1714-1724 (every pushed entry counts); -2 = assigned, then reset to NULL
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        nmatches -= removed;
    }
    if (lane == 0) *nmatches_out = nmatches;
}

// ---- SearchByProjection(Frame, vpMapPoints), sequential part (ORBmatcher.cc:77-172) ----
struct ProjPointsOps {
    int* occ; int* f_match; const int* oct; const KpM* f_keys; const int* obs; const uint8_t* mp_observed; float nnratio; bool staged; int nmatches;
    __device__ __forceinline__ bool clean(uint32_t v) const { return !occ[v & 0xFFFFFu]; }                                        // :124-126
    static constexpr uint32_t dlimit = 256u;
    __device__ __forceinline__ bool decide(int, const uint32_t* c, Best2 b, int& j) const {
        if (b.k1 == NOJ) return false;
        const int bestDist = (int)(b.k1 >> 20);
        if (bestDist > M_TH_HIGH) return false;                                                                                  // :163
        const int bestIdx = (int)(c[b.k1 & 0xFFFFFu] & 0xFFFFFu);
        const int bestLevel = staged ? oct[bestIdx] : f_keys[bestIdx].octave;
        int bestDist2 = 256, bestLevel2 = -1;
        if (b.k2 != NOJ) { bestDist2 = (int)(b.k2 >> 20); const int j2 = (int)(c[b.k2 & 0xFFFFFu] & 0xFFFFFu); bestLevel2 = staged ? oct[j2] : f_keys[j2].octave; }
        if (bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nnratio, (float)bestDist2)) return false;                     // :166
        j = bestIdx;
        return true;
    }
    __device__ __forceinline__ bool commit(int i, int j, Best2, int) {
        f_match[j] = i;
        occ[j] = staged ? obs[i] : (mp_observed ? (mp_observed[i] != 0) : 0);
        return false;
    }
    __device__ __forceinline__ void post(int na, int) { nmatches += na; }
};

__global__ void __launch_bounds__(RESOLVE_THREADS)
k_resolve_proj_points(int n_points, int n_f, const KpM* __restrict__ f_keys, const uint8_t* __restrict__ mp_observed, const uint8_t* __restrict__ f_occupied,
                      const int* __restrict__ counts, const int* __restrict__ offsets, const uint32_t* __restrict__ cand, const uint2* __restrict__ pre_best,
                      int cand_cap, float nnratio, int smem_words, int* __restrict__ occupied_g, int* __restrict__ f_match, int* __restrict__ nmatches_out) {
    extern __shared__ __align__(16) uint32_t rs_sm[];
    const int tid = threadIdx.x, lane = tid & 31;
    if (offsets[n_points] > cand_cap) return;
    // extra shared arrays: occ[n_f] | oct[n_f] | obs[n_points]
    const StagedLists SL = stage_lists(rs_sm, smem_words, 2 * n_f + n_points, n_points, counts, offsets, cand, pre_best);
    const bool staged = SL.staged;
    int* occ = staged ? reinterpret_cast<int*>(SL.free_words) : occupied_g;
    int* oct = reinterpret_cast<int*>(SL.free_words) + n_f;
    int* obs = oct + n_f;
    for (int j = tid; j < n_f; j += RESOLVE_THREADS) {
        f_match[j] = -1;
        occ[j] = f_occupied ? (f_occupied[j] != 0) : 0;
        if (staged) oct[j] = f_keys[j].octave;
    }
    if (staged) for (int i = tid; i < n_points; i += RESOLVE_THREADS) obs[i] = mp_observed ? (mp_observed[i] != 0) : 0;
    __syncthreads();
    if (tid >= 32) return;
    ProjPointsOps ops{occ, f_match, oct, f_keys, obs, mp_observed, nnratio, staged, 0};
    replay_queries<true>(SL, counts, offsets, lane, ops);
    if (lane == 0) *nmatches_out = ops.nmatches;
}

// =================================================================================================
// Frame::ComputeStereoMatches   /root/reference/src/Frame.cc:1179-1573
// =================================================================================================
struct PyrLevelDev { const uint8_t* ptr; int pitch, w, h; long long fstride; };    // fstride: distance between the frames of a batch (0 for a single frame)
// batched form: stereo pair b = blockIdx.y reads its keypoints at k + b * stride, descriptors at d + b * stride * 32, counts from nl_arr / nr_arr
struct StereoBatch { const int* nl_arr; const int* nr_arr; int stride; };
struct StereoPyr { PyrLevelDev lv[ORBX_MAX_LEVELS]; };

// vRowIndices (:1228-1251): for every image row the right keypoints whose band [floor(y - r), ceil(y + r)], r = 2 * scale[octave], covers it.  One CTA per stereo
// pair: band sizes counted per row in shared memory, exclusive scan, fill.  The order inside a row is arbitrary (the search below takes the minimum of
// distance << 20 | index, which is the reference's "first of the smallest" whatever the order).  Without it every left keypoint walked all right keypoints
// (2000 band tests for ~50 candidates: 0.71 ms per 64 KITTI pairs, the largest stage of C4).
#define STEREO_MAX_ROWS 4096
__global__ void __launch_bounds__(1024)
k_stereo_rows(const KpM* __restrict__ kr, int nr, int nRows, const float* __restrict__ scale, int* __restrict__ row_off, int* __restrict__ entries, int ent_cap, StereoBatch SB) {
    __shared__ int cnt[STEREO_MAX_ROWS + 1];
    __shared__ int wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fb = blockIdx.x;
    if (SB.nl_arr) { nr = min(SB.nr_arr[fb], SB.stride); kr += (size_t)fb * SB.stride; }
    row_off += (size_t)fb * (nRows + 1); entries += (size_t)fb * ent_cap;
    for (int r = tid; r <= nRows; r += 1024) cnt[r] = 0;
    __syncthreads();
    for (int iR = tid; iR < nr; iR += 1024) {
        const KpM k = kr[iR];
        const float r = __fmul_rn(2.0f, scale[k.octave]);                      // :1239
        const int maxr = min((int)ceilf(__fadd_rn(k.y, r)), nRows - 1), minr = max((int)floorf(__fsub_rn(k.y, r)), 0);
        for (int y = minr; y <= maxr; ++y) atomicAdd(&cnt[y], 1);
    }
    __syncthreads();
    // exclusive scan of cnt[0 .. nRows): every thread owns a run of rows
    const int per = (nRows + 1023) / 1024, r0 = min(tid * per, nRows), r1 = min(r0 + per, nRows);
    int mine = 0;
    for (int r = r0; r < r1; ++r) mine += cnt[r];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += wsum[w];
    int run = base + incl - mine;
    for (int r = r0; r < r1; ++r) { const int c = cnt[r]; cnt[r] = run; row_off[r] = run; run += c; }
    if (tid == 1023) row_off[nRows] = run;
    __syncthreads();
    for (int iR = tid; iR < nr; iR += 1024) {
        const KpM k = kr[iR];
        const float r = __fmul_rn(2.0f, scale[k.octave]);
        const int maxr = min((int)ceilf(__fadd_rn(k.y, r)), nRows - 1), minr = max((int)floorf(__fsub_rn(k.y, r)), 0);
        for (int y = minr; y <= maxr; ++y) { const int p = atomicAdd(&cnt[y], 1); if (p < ent_cap) entries[p] = iR; }
    }
}

// one warp per left keypoint: row-band Hamming search over the right keypoints of its row (k_stereo_rows), then the 11-shift 11x11 SAD
// refinement on the two pyramids, parabola sub-pixel fit, disparity -> depth.
__global__ void __launch_bounds__(128)
k_stereo_match(const KpM* __restrict__ kl, const uint8_t* __restrict__ dl, int nl, const KpM* __restrict__ kr, const uint8_t* __restrict__ dr, int nr,
               StereoPyr PL, StereoPyr PR, const float* __restrict__ scale, const float* __restrict__ inv_scale, float mb, float mbf,
               float* __restrict__ u_right, float* __restrict__ depth, int* __restrict__ sad_dist /* INT_MAX = not stored */, StereoBatch SB,
               const int* __restrict__ row_off, const int* __restrict__ entries, int ent_cap) {
    const int lane = threadIdx.x & 31;
    const int iL = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int fb = blockIdx.y;
    if (SB.nl_arr) {
        nl = min(SB.nl_arr[fb], SB.stride);
        const size_t o = (size_t)fb * SB.stride;
        kl += o; kr += o; dl += o * 32; dr += o * 32; u_right += o; depth += o; sad_dist += o;
    }
    if (iL >= nl) { if (SB.nl_arr && iL < SB.stride && lane == 0) { u_right[iL] = -1.0f; depth[iL] = -1.0f; sad_dist[iL] = INT_MAX; } return; }
    if (lane == 0) { u_right[iL] = -1.0f; depth[iL] = -1.0f; sad_dist[iL] = INT_MAX; }
    const KpM kpL = kl[iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const int row = (int)vL;                                                   // vRowIndices[vL]  :1298
    const int nRows = PL.lv[0].h;
    if (row < 0 || row >= nRows) return;
    const float minZ = mb, minD = 0.f;
    const float maxD = __fdiv_rn(mbf, minZ);                                   // :1272 (mb == 0 => +inf)
    const float minU = __fsub_rn(uL, maxD), maxU = __fsub_rn(uL, minD);
    if (maxU < 0.f) return;
    const uint4* qd = reinterpret_cast<const uint4*>(dl) + 2 * iL;
    const uint4 d0 = __ldg(qd), d1 = __ldg(qd + 1);
    uint32_t best = 0xFFFFFFFFu;                                               // dist<<20 | iR ; first (smallest iR) wins ties
    (void)nr;                                                                  // (the row lists already bound the right keypoints)
    row_off += (size_t)fb * (nRows + 1); entries += (size_t)fb * ent_cap;
    const int e1 = min(row_off[row + 1], ent_cap);
    for (int e = row_off[row] + lane; e < e1; e += 32) {                       // vRowIndices[vL]  :1298
        const int iR = entries[e];
        const KpM kpR = kr[iR];
        if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;      // :1343
        if (!(kpR.x >= minU && kpR.x <= maxU)) continue;                       // :1353
        const int d = hamming256(d0, d1, reinterpret_cast<const uint4*>(dr) + 2 * iR);
        if (d < M_TH_HIGH) best = min(best, ((uint32_t)d << 20) | (uint32_t)iR);     // bestDist starts at TH_HIGH  :1322
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (best == 0xFFFFFFFFu) return;
    const int bestDist = (int)(best >> 20), bestIdxR = (int)(best & 0xFFFFFu);
    if (!(bestDist < (M_TH_HIGH + M_TH_LOW) / 2)) return;                      // thOrbDist  :1194,1377
    const float uR0 = kr[bestIdxR].x;
    const float sf = inv_scale[levelL];
    const float scaleduL = roundf(__fmul_rn(kpL.x, sf)), scaledvL = roundf(__fmul_rn(kpL.y, sf)), scaleduR0 = roundf(__fmul_rn(uR0, sf));
    const int w = 5, L = 5;
    PyrLevelDev IL = PL.lv[levelL], IR = PR.lv[levelL];
    IL.ptr += fb * IL.fstride; IR.ptr += fb * IR.fstride;
    const float iniu = __fsub_rn(__fadd_rn(scaleduR0, (float)L), (float)w), endu = __fadd_rn(__fadd_rn(__fadd_rn(scaleduR0, (float)L), (float)w), 1.f);
    if (iniu < 0.f || endu >= (float)IR.w) return;                              // :1425
    const int cy = (int)scaledvL, cxl = (int)scaleduL, cxr = (int)scaleduR0;
    // the reference's rowRange/colRange would throw on windows leaving the level; cannot happen for keypoints that
    // are >= 19 px inside their level, guarded here so foreign input can never read out of bounds
    if (cy - w < 0 || cy + w >= IL.h || cy + w >= IR.h || cxl - w < 0 || cxl + w >= IL.w || cxr - L - w < 0 || cxr + L + w >= IR.w) return;
    const int cl = IL.ptr[(size_t)cy * IL.pitch + cxl];
    float vDists[11];
    int bestS = INT_MAX, bestinc = 0;
    // the left patch does not move with the shift: its (up to) four pixels of this lane are read once, and so are the row offsets of the right patch
    int aL[4]; const uint8_t* pR[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int p = lane + 32 * k;
        const int dy = p / 11 - w, dx = p % 11 - w;
        aL[k] = p < 121 ? (int)IL.ptr[(size_t)(cy + dy) * IL.pitch + cxl + dx] - cl : 0;
        pR[k] = IR.ptr + (size_t)(cy + (p < 121 ? dy : 0)) * IR.pitch + cxr + (p < 121 ? dx : 0);
    }
    const uint8_t* pC = IR.ptr + (size_t)cy * IR.pitch + cxr;
#pragma unroll 1
    for (int inc = -L; inc <= L; ++inc) {
        const int cr = pC[inc];
        int s = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (lane + 32 * k < 121) s += abs(aL[k] - ((int)pR[k][inc] - cr));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float dist = (float)s;                                           // cv::norm(IL, IR, NORM_L1): exact integers
        if (dist < (float)bestS) { bestS = (int)dist; bestinc = inc; }         // :1449 (int bestDist compared as float)
        vDists[L + inc] = dist;
    }
    if (bestinc == -L || bestinc == L) return;                                  // :1468
    float dist1 = 0.f, dist2 = 0.f, dist3 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) { if (k == L + bestinc - 1) dist1 = vDists[k]; if (k == L + bestinc) dist2 = vDists[k]; if (k == L + bestinc + 1) dist3 = vDists[k]; }
    // deltaR = (dist1-dist3)/(2.0f*(dist1+dist3-2.0f*dist2))   :1494
    const float deltaR = __fdiv_rn(__fsub_rn(dist1, dist3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2))));
    if (deltaR < -1.f || deltaR > 1.f) return;
    float bestuR = __fmul_rn(scale[levelL], __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));      // :1509
    float disparity = __fsub_rn(uL, bestuR);
    if (disparity >= minD && disparity < maxD) {
        if (disparity <= 0.f) { disparity = 0.01f; bestuR = __fsub_rn(uL, 0.01f); }
        if (lane == 0) { depth[iL] = __fdiv_rn(mbf, disparity); u_right[iL] = bestuR; sad_dist[iL] = bestS; }
    }
}

// median cut (:1548-1569): thDist = 1.5f*1.4f*median(dist); entries with dist >= thDist are dropped.  One CTA.
#define STEREO_SMEM_VALS 8192
__global__ void __launch_bounds__(1024)
k_stereo_median_cut(int nl, const int* __restrict__ sad_dist, float* __restrict__ u_right, float* __restrict__ depth, StereoBatch SB) {
    if (SB.nl_arr) { nl = min(SB.nl_arr[blockIdx.x], SB.stride); const size_t o = (size_t)blockIdx.x * SB.stride; sad_dist += o; u_right += o; depth += o; }
    // the SAD values of the surviving matches are compacted into shared memory (order is irrelevant for a rank), so the rank
    // counting below reads broadcast shared words instead of re-walking the global array once per match
    __shared__ __align__(16) int vals[STEREO_SMEM_VALS];
    __shared__ int total_sm, median_sm;
    const int tid = threadIdx.x;
    if (tid == 0) { total_sm = 0; median_sm = -1; }
    __syncthreads();
    for (int i0 = 0; i0 < nl; i0 += 1024) {                                    // (one shared-memory atomic per warp, not per match)
        const int i = i0 + tid;
        const int d = i < nl ? sad_dist[i] : INT_MAX;
        const uint32_t mk = __ballot_sync(0xffffffffu, d != INT_MAX);
        int base = 0;
        if (mk && (tid & 31) == __ffs(mk) - 1) base = atomicAdd(&total_sm, __popc(mk));
        base = __shfl_sync(0xffffffffu, base, mk ? __ffs(mk) - 1 : 0);
        if (d != INT_MAX) { const int p = base + __popc(mk & ((1u << (tid & 31)) - 1u)); if (p < STEREO_SMEM_VALS) vals[p] = d; }
    }
    __syncthreads();
    const int total = total_sm;
    if (total == 0) return;
    const int k = total / 2;                                                   // vDistIdx[size/2].first of the sorted list
    if (total <= STEREO_SMEM_VALS) {
        // k-th smallest by a two-level radix select on the 16-bit SAD values (121 pixels x |difference| <= 510): histogram of the high byte, the bucket holding rank k,
        // histogram of the low byte inside it -- linear in the number of matches (the rank counting it replaces compared every pair of matches)
        __shared__ int hist[256];
        __shared__ int sel_bucket, sel_rank;
        int want = k, key_hi = 0;
        for (int level = 0; level < 2; ++level) {
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
            for (int i = tid; i < total; i += 1024) {
                const int d = min(vals[i], 0xFFFF);
                if (level == 0) atomicAdd(&hist[d >> 8], 1);
                else if ((d >> 8) == key_hi) atomicAdd(&hist[d & 255], 1);
            }
            __syncthreads();
            if (tid < 32) {                                                    // lane l owns buckets 8 l .. 8 l + 7
                int c[8], mine = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) { c[q] = hist[8 * tid + q]; mine += c[q]; }
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += t; }
                int before = incl - mine;
                if (before <= want && want < incl) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) { if (want >= before && want < before + c[q]) { sel_bucket = 8 * tid + q; sel_rank = want - before; } before += c[q]; }
                }
            }
            __syncthreads();
            if (level == 0) { key_hi = sel_bucket; want = sel_rank; }
            else if (tid == 0) median_sm = (key_hi << 8) | sel_bucket;
            __syncthreads();
        }
    } else {
        for (int i = tid; i < nl; i += 1024) {
            const int d = sad_dist[i];
            if (d == INT_MAX) continue;
            int less = 0, leq = 0;
            for (int j = 0; j < nl; ++j) { const int e = sad_dist[j]; if (e != INT_MAX) { less += e < d; leq += e <= d; } }
            if (less <= k && k < leq) median_sm = d;
        }
    }
    __syncthreads();
    const float median = (float)median_sm;
    const float thDist = __fmul_rn(1.5f * 1.4f, median);
    for (int i = tid; i < nl; i += 1024) {
        const int d = sad_dist[i];
        if (d != INT_MAX && !((float)d < thDist)) { u_right[i] = -1.f; depth[i] = -1.f; }
    }
}

// =================================================================================================
// Brute-force all-pairs best / second-best (throughput form of the inner kernel).
// The first version (one query per warp, 8 POPC per distance, one 32-byte shared-memory read per distance) ran into two walls at once
// (profiles/r01h_bruteforce_raw.csv): the XU pipe, which executes POPC, at 118 % of its sustained rate and the LSU wavefronts at 93 %
// (2-way bank conflicts on the 32-byte descriptor stride), with the ALU pipe at 16 %.  This version moves work to where there is room:
//   * every warp holds BF_QPW queries in registers, so a train descriptor read from shared memory serves BF_QPW distances;
//   * the tile is stored as two 16-byte planes, so a warp's reads are conflict-free;
//   * the 8 XOR words of a distance go through four carry-save adders (8 LOP3 on the ALU pipe) and need 4 POPC instead of 8:
//     d = popc(s2) + popc(x7) + 2 popc(twos) + 4 popc(fours)  (a truncated Harley-Seal tree, chosen so that ALU and XU pipes are
//     loaded about equally), the weighted sum on the FMA pipe (IMAD).
// Results are identical: key = distance << 20 | train index, smallest two keys per query.
// =================================================================================================
#define BF_TILE 256
#define BF_QPW 4
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t& sum, uint32_t& carry) {
    sum = a ^ b ^ c; carry = (a & b) | (a & c) | (b & c);                       // one LOP3 each
}
__device__ __forceinline__ int hamming256_hs(const uint32_t* __restrict__ q, const uint4 b0, const uint4 b1) {
    const uint32_t x0 = q[0] ^ b0.x, x1 = q[1] ^ b0.y, x2 = q[2] ^ b0.z, x3 = q[3] ^ b0.w, x4 = q[4] ^ b1.x, x5 = q[5] ^ b1.y, x6 = q[6] ^ b1.z, x7 = q[7] ^ b1.w;
    uint32_t s0, c0, s1, c1, s2, c2, t0, e0;
    csa(x0, x1, x2, s0, c0); csa(x3, x4, x5, s1, c1); csa(s0, s1, x6, s2, c2);    // ones: s2 and x7 remain
    csa(c0, c1, c2, t0, e0);                                                       // twos: t0 remains; fours: e0
    return __popc(s2) + __popc(x7) + 2 * __popc(t0) + 4 * __popc(e0);
}
__global__ void __launch_bounds__(256)
k_bruteforce_best2(const uint4* __restrict__ query, int nq, const uint4* __restrict__ train, int nt,
                   int* __restrict__ best_idx, int* __restrict__ best_dist, int* __restrict__ second_dist) {
    // blockIdx.y = frame pair of a batch: pair p matches query[p*nq ..] against train[p*nt ..]
    query += 2 * (size_t)blockIdx.y * nq; train += 2 * (size_t)blockIdx.y * nt;
    best_idx += (size_t)blockIdx.y * nq; best_dist += (size_t)blockIdx.y * nq; second_dist += (size_t)blockIdx.y * nq;
    __shared__ uint4 tile_lo[BF_TILE], tile_hi[BF_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = (blockIdx.x * 8 + warp) * BF_QPW;
    uint32_t qd[BF_QPW][8], k1[BF_QPW], k2[BF_QPW];
#pragma unroll
    for (int u = 0; u < BF_QPW; ++u) {
        uint4 a = make_uint4(0, 0, 0, 0), b = a;
        if (q0 + u < nq) { a = __ldg(query + 2 * (q0 + u)); b = __ldg(query + 2 * (q0 + u) + 1); }
        qd[u][0] = a.x; qd[u][1] = a.y; qd[u][2] = a.z; qd[u][3] = a.w; qd[u][4] = b.x; qd[u][5] = b.y; qd[u][6] = b.z; qd[u][7] = b.w;
        k1[u] = 0xFFFFFFFFu; k2[u] = 0xFFFFFFFFu;
    }
    for (int t0 = 0; t0 < nt; t0 += BF_TILE) {
        const int tn = min(BF_TILE, nt - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn * 2; i += 256) { const uint4 v = __ldg(train + 2 * (size_t)t0 + i); if (i & 1) tile_hi[i >> 1] = v; else tile_lo[i >> 1] = v; }
        __syncthreads();
        if (q0 < nq)
            for (int j = lane; j < tn; j += 32) {
                const uint4 b0 = tile_lo[j], b1 = tile_hi[j];
#pragma unroll
                for (int u = 0; u < BF_QPW; ++u) {
                    const uint32_t key = ((uint32_t)hamming256_hs(qd[u], b0, b1) << 20) | (uint32_t)(t0 + j);
                    k2[u] = min(k2[u], max(k1[u], key)); k1[u] = min(k1[u], key);    // smallest two keys, branch-free (keys are distinct)
                }
            }
    }
#pragma unroll
    for (int u = 0; u < BF_QPW; ++u) {
        const int q = q0 + u;
        if (q >= nq) break;                                                     // uniform across the warp
        uint32_t a = k1[u], b = k2[u];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t a1 = __shfl_xor_sync(0xffffffffu, a, o), a2 = __shfl_xor_sync(0xffffffffu, b, o);
            const uint32_t lo = min(a, a1), hi = max(a, a1);
            b = min(hi, min(b, a2)); a = lo;
        }
        if (lane == 0) {
            best_idx[q] = a == 0xFFFFFFFFu ? -1 : (int)(a & 0xFFFFFu);
            best_dist[q] = a == 0xFFFFFFFFu ? INT_MAX : (int)(a >> 20);
            second_dist[q] = b == 0xFFFFFFFFu ? INT_MAX : (int)(b >> 20);
        }
    }
}
