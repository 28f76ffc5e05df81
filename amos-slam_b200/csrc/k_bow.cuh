// k_bow.cuh -- bag of words on the device.
//   DBoW2::TemplatedVocabulary::transform (descriptor -> word / node, descriptor set -> BowVector + FeatureVector)
//     /root/reference/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1124-1197, 1217-1259; FORB::distance FORB.cpp:81-101;
//     BowVector::addWeight / addIfNotExist / normalize BowVector.cpp:29-87; FeatureVector::addFeature FeatureVector.cpp:29-43
//   ORBmatcher::SearchByBoW x2   /root/reference/src/ORBmatcher.cc:230-382, 656-799
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "k_match.cuh"

struct VocDev {
    int k, L, weighting, scoring, n;                 // n = nodes incl. the root (id 0)
    const int* child_off;                            // [n + 1] into the child list
    const int* child_id;                             // child list: node ids, in the reference's push_back order
    const uint4* child_desc;                         // descriptors in child-list order (the children of a node are contiguous: one 32 k byte block)
    const double* weight; const int* word;           // by node id (word = -1 for inner nodes)
};

// one warp per descriptor: at every level the lanes take one child each, the FIRST child with the smallest distance wins (strict <)
__global__ void __launch_bounds__(128)
k_bow_descend(VocDev v, const uint4* __restrict__ desc, int n, int levelsup, int* __restrict__ word_of, int* __restrict__ node_of, double* __restrict__ weight_of) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= n) return;
    const uint4 d0 = __ldg(desc + 2 * i), d1 = __ldg(desc + 2 * i + 1);
    const int nid_level = v.L - levelsup;
    int cur = 0, level = 0, nid = 0;
    while (true) {
        const int c0 = v.child_off[cur], c1 = v.child_off[cur + 1];
        if (c0 == c1) break;                                                    // leaf
        ++level;
        uint32_t best = 0xFFFFFFFFu;
        for (int c = c0 + lane; c < c1; c += 32) {
            const uint32_t key = ((uint32_t)hamming256(d0, d1, v.child_desc + 2 * c) << 20) | (uint32_t)(c - c0);
            best = min(best, key);
        }
        best = __reduce_min_sync(0xffffffffu, best);
        cur = v.child_id[c0 + (int)(best & 0xFFFFFu)];
        if (level == nid_level) nid = cur;
    }
    if (lane == 0) { word_of[i] = v.word[cur]; node_of[i] = nid; weight_of[i] = v.weight[cur]; }
}

// ---- BowVector / FeatureVector assembly: one CTA ----
__device__ __forceinline__ void bitonic_u64(unsigned long long* a, int n, int tid, int nt) {
    for (int k = 2; (k >> 1) < n; k <<= 1) {
        for (int i = tid; i < n; i += nt) { const int j = i ^ (k - 1); if (j > i && j < n) { unsigned long long x = a[i], y = a[j]; if (x > y) { a[i] = y; a[j] = x; } } }
        __syncthreads();
        for (int s = k >> 2; s > 0; s >>= 1) {
            for (int i = tid; i < n; i += nt) { const int j = i ^ s; if (j > i && j < n) { unsigned long long x = a[i], y = a[j]; if (x > y) { a[i] = y; a[j] = x; } } }
            __syncthreads();
        }
    }
}
// exclusive scan of flags held in sc[0..n) (int), in place; returns the total.  One CTA.
__device__ int block_scan_excl(int* sc, int n, int tid, int nt, int* carry_sm) {
    // simple chunked Hillis-Steele over nt-sized chunks
    __shared__ int wsum[32];
    if (tid == 0) *carry_sm = 0;
    __syncthreads();
    for (int base = 0; base < n; base += nt) {
        const int i = base + tid;
        const int vv = i < n ? sc[i] : 0;
        int incl = vv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += t; }
        if ((tid & 31) == 31) wsum[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) { const int w = (tid < (nt >> 5)) ? wsum[tid] : 0; int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (tid >= o) wi += t; }
            wsum[tid] = wi - w; }
        __syncthreads();
        const int carry = *carry_sm;
        if (i < n) sc[i] = carry + wsum[tid >> 5] + incl - vv;
        __syncthreads();
        if (tid == nt - 1) *carry_sm = carry + wsum[tid >> 5] + incl;
        __syncthreads();
    }
    return *carry_sm;
}

#define BOW_SMEM_KEYS 4096
// keys: scratch [n + 1] u64; flags: scratch [n + 1] int; stage: scratch [n] double.  Outputs as the C ABI describes them; counts[0] = n_bow, counts[1] = n_fv.
__global__ void __launch_bounds__(1024)
k_bow_assemble(int n, int weighting, int scoring, const int* __restrict__ word_of, const int* __restrict__ node_of, const double* __restrict__ weight_of,
               unsigned long long* __restrict__ keys, int* __restrict__ flags, double* __restrict__ stage,
               int* __restrict__ bow_ids, double* __restrict__ bow_vals, int* __restrict__ fv_nodes, int* __restrict__ fv_offsets, int* __restrict__ fv_idx, int* __restrict__ counts) {
    __shared__ int carry, n_live_sm;
    __shared__ unsigned long long skeys[BOW_SMEM_KEYS + 1];             // a frame's worth of keys sorts in shared memory
    if (n <= BOW_SMEM_KEYS) keys = skeys;
    const int tid = threadIdx.x, nt = blockDim.x;
    // ---- BowVector: sort (word, feature) of the non-stopped features ----
    for (int i = tid; i < n; i += nt) keys[i] = (weight_of[i] > 0.0) ? (((unsigned long long)(uint32_t)word_of[i] << 32) | (uint32_t)i) : ~0ull;
    __syncthreads();
    bitonic_u64(keys, n, tid, nt);
    for (int i = tid; i < n; i += nt) flags[i] = (keys[i] != ~0ull && (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32))) ? 1 : 0;
    if (tid == 0) { int live = 0; /* stopped features sort to the end */ int lo = 0, hi = n; while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] != ~0ull) lo = mid + 1; else hi = mid; } live = lo; n_live_sm = live; }
    __syncthreads();
    const int n_live = n_live_sm;
    // per word: the running double sum the reference builds in feature order (first insert, then +=), or the first weight (IDF / BINARY)
    for (int i = tid; i < n_live; i += nt) {
        if (!flags[i]) continue;
        const unsigned long long w = keys[i] >> 32;
        double s = weight_of[(uint32_t)keys[i]];
        if (weighting == 0 || weighting == 1) for (int j = i + 1; j < n_live && (keys[j] >> 32) == w; ++j) s = __dadd_rn(s, weight_of[(uint32_t)keys[j]]);
        stage[i] = s;                                                   // at the head position; compacted below
    }
    __syncthreads();
    const int n_bow = block_scan_excl(flags, n_live, tid, nt, &carry);
    for (int i = tid; i < n_live; i += nt) {
        const bool head = (i == 0) || ((keys[i] >> 32) != (keys[i - 1] >> 32));
        if (head) { bow_vals[flags[i]] = stage[i]; bow_ids[flags[i]] = (int)(keys[i] >> 32); }
    }
    __syncthreads();
    // normalisation (BowVector::normalize, or the "/ size" of un-normalised TF weights): sums in map (word id) order by ONE thread
    const bool must = scoring != 5;                                     // DOT_PRODUCT is the only scoring that does not normalise
    double* sv = reinterpret_cast<double*>(keys);                       // the sorted keys are spent: reuse the buffer for the serial sum
    for (int i = tid; i < n_bow; i += nt) sv[i] = (scoring == 1) ? __dmul_rn(bow_vals[i], bow_vals[i]) : fabs(bow_vals[i]);
    __syncthreads();
    if (tid == 0) {
        double norm = 0.0;
        if (must) {
            for (int i = 0; i < n_bow; ++i) norm = __dadd_rn(norm, sv[i]);
            if (scoring == 1) norm = __dsqrt_rn(norm);
        } else if (weighting == 0 || weighting == 1) norm = (double)n_bow;
        sv[n] = norm;                                                  // keys has n + 1 slots
        counts[0] = n_bow;
    }
    __syncthreads();
    const double norm = sv[n];
    __syncthreads();
    if (norm > 0.0) for (int i = tid; i < n_bow; i += nt) bow_vals[i] = __ddiv_rn(bow_vals[i], norm);
    __syncthreads();
    // ---- FeatureVector: sort (node, feature) of the same features ----
    for (int i = tid; i < n; i += nt) keys[i] = (weight_of[i] > 0.0) ? (((unsigned long long)(uint32_t)node_of[i] << 32) | (uint32_t)i) : ~0ull;
    __syncthreads();
    bitonic_u64(keys, n, tid, nt);
    for (int i = tid; i < n_live; i += nt) { fv_idx[i] = (int)(uint32_t)keys[i]; flags[i] = (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32)) ? 1 : 0; }
    __syncthreads();
    const int n_fv = block_scan_excl(flags, n_live, tid, nt, &carry);
    for (int i = tid; i < n_live; i += nt) {
        const bool head = (i == 0) || ((keys[i] >> 32) != (keys[i - 1] >> 32));
        if (head) { fv_nodes[flags[i]] = (int)(keys[i] >> 32); fv_offsets[flags[i]] = i; }
    }
    if (tid == 0) { fv_offsets[n_fv] = n_live; counts[1] = n_fv; }
}

// ---- SearchByBoW: one warp per common node; inside, the side-1 features in list order, lanes over the side-2 candidates ----
struct BowSideDev { int n; const KpM* keys; const uint4* desc; const uint8_t* valid; const int* fv_offsets; const int* fv_idx; };

__global__ void __launch_bounds__(128)
k_bow_match(int n_pairs, const int2* __restrict__ pairs /* (position in fv1, position in fv2) of every common node */, BowSideDev s1, BowSideDev s2,
            int kf_kf, float nnratio, int checkOri, int* __restrict__ match12, int* __restrict__ match21, int* __restrict__ bin_of, int* __restrict__ hist) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (p >= n_pairs) return;
    const int2 pr = pairs[p];
    const int a0 = s1.fv_offsets[pr.x], a1 = s1.fv_offsets[pr.x + 1], b0 = s2.fv_offsets[pr.y], b1 = s2.fv_offsets[pr.y + 1];
    volatile int* m21 = match21;
    for (int e1 = a0; e1 < a1; ++e1) {
        const int i1 = s1.fv_idx[e1];
        if (!s1.valid[i1]) continue;                                            // no map point, or a bad one (:272-276 / :713-716)
        const uint4 d0 = __ldg(s1.desc + 2 * i1), d1 = __ldg(s1.desc + 2 * i1 + 1);
        uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;                             // dist << 20 | position in the node's list
        for (int e2 = b0 + lane; e2 < b1; e2 += 32) {
            const int i2 = s2.fv_idx[e2];
            if (m21[i2] >= 0) continue;                                         // already matched (:291 / :728)
            if (s2.valid && !s2.valid[i2]) continue;                          // KeyFrame x KeyFrame: pMP2 missing or bad (:728-731)
            const uint32_t key = ((uint32_t)hamming256(d0, d1, s2.desc + 2 * i2) << 20) | (uint32_t)(e2 - b0);
            if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
        }
        const uint32_t g1 = __reduce_min_sync(0xffffffffu, k1);
        const uint32_t g2 = __reduce_min_sync(0xffffffffu, (k1 == g1 && g1 != 0xFFFFFFFFu) ? k2 : k1);
        if (g1 == 0xFFFFFFFFu) continue;
        const int best1 = (int)(g1 >> 20), best2 = g2 == 0xFFFFFFFFu ? 256 : (int)(g2 >> 20);
        const bool close = kf_kf ? (best1 < 50) : (best1 <= 50);               // TH_LOW: :741 vs :313
        if (close && (float)best1 < __fmul_rn(nnratio, (float)best2)) {
            const int bidx = s2.fv_idx[b0 + (int)(g1 & 0xFFFFFu)];
            if (lane == 0) {
                match12[i1] = bidx; m21[bidx] = i1;
                if (checkOri) { const int bin = rot_bin(s1.keys[i1].angle, s2.keys[bidx].angle); bin_of[i1] = bin; atomicAdd(&hist[bin], 1); }
            }
            __syncwarp();
        }
    }
}

// ---- SearchForTriangulation (ORBmatcher.cc:810-1010): same node-by-node replay, other candidate rules ----
struct TriParams {
    float F[9];                       // F12, row-major (CV_32F)
    float ex, ey;                     // epipole of camera 1 in image 2 (:822-825)
    const float* ur1; const float* ur2;            // mvuRight of both KeyFrames (stereo iff >= 0)
    const float* scale2; const float* sigma2_2;    // pKF2->mvScaleFactors, mvLevelSigma2
    int only_stereo;
};
// CheckDistEpipolarLine (:188-215): every operation a rounded float product / sum in the reference's order; the last comparison in double
__device__ __forceinline__ bool epipolar_ok(const TriParams& T, float x1, float y1, float x2, float y2, int oct2) {
    const float a = __fadd_rn(__fadd_rn(__fmul_rn(x1, T.F[0]), __fmul_rn(y1, T.F[3])), T.F[6]);
    const float b = __fadd_rn(__fadd_rn(__fmul_rn(x1, T.F[1]), __fmul_rn(y1, T.F[4])), T.F[7]);
    const float c = __fadd_rn(__fadd_rn(__fmul_rn(x1, T.F[2]), __fmul_rn(y1, T.F[5])), T.F[8]);
    const float num = __fadd_rn(__fadd_rn(__fmul_rn(a, x2), __fmul_rn(b, y2)), c);
    const float den = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
    if (den == 0.f) return false;
    const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
    return (double)dsqr < __dmul_rn(3.84, (double)T.sigma2_2[oct2]);
}

__global__ void __launch_bounds__(128)
k_tri_match(int n_pairs, const int2* __restrict__ pairs, BowSideDev s1, BowSideDev s2, TriParams T, int checkOri,
            int* __restrict__ match12, int* __restrict__ matched2, int* __restrict__ bin_of, int* __restrict__ hist) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (p >= n_pairs) return;
    const int2 pr = pairs[p];
    const int a0 = s1.fv_offsets[pr.x], a1 = s1.fv_offsets[pr.x + 1], b0 = s2.fv_offsets[pr.y], b1 = s2.fv_offsets[pr.y + 1];
    volatile int* m2 = matched2;
    for (int e1 = a0; e1 < a1; ++e1) {
        const int i1 = s1.fv_idx[e1];
        if (!s1.valid[i1]) continue;                                            // holds a map point already (:843-845)
        const bool stereo1 = T.ur1[i1] >= 0.f;
        if (T.only_stereo && !stereo1) continue;
        const KpM kp1 = s1.keys[i1];
        const uint4 d0 = __ldg(s1.desc + 2 * i1), d1 = __ldg(s1.desc + 2 * i1 + 1);
        // the reference keeps the LAST candidate among those with the smallest distance that pass the geometric gates
        // (dist > bestDist is skipped, dist == bestDist replaces): key = dist << 20 | (0xFFFFF - position), smallest wins
        uint32_t best = 0xFFFFFFFFu;
        for (int e2 = b0 + lane; e2 < b1; e2 += 32) {
            const int i2 = s2.fv_idx[e2];
            if (m2[i2] >= 0 || !s2.valid[i2]) continue;                         // vbMatched2[idx2] || pMP2  (:862)
            const bool stereo2 = T.ur2[i2] >= 0.f;
            if (T.only_stereo && !stereo2) continue;
            const int dist = hamming256(d0, d1, s2.desc + 2 * i2);
            if (dist > 50) continue;                                            // TH_LOW (:873)
            const KpM kp2 = s2.keys[i2];
            if (!stereo1 && !stereo2) {
                const float dx = __fsub_rn(T.ex, kp2.x), dy = __fsub_rn(T.ey, kp2.y);
                if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, T.scale2[kp2.octave])) continue;   // too close to the epipole (:880-885)
            }
            if (!epipolar_ok(T, kp1.x, kp1.y, kp2.x, kp2.y, kp2.octave)) continue;
            best = min(best, ((uint32_t)dist << 20) | (0xFFFFFu - (uint32_t)(e2 - b0)));
        }
        best = __reduce_min_sync(0xffffffffu, best);
        if (best == 0xFFFFFFFFu) continue;
        const int bidx = s2.fv_idx[b0 + (int)(0xFFFFFu - (best & 0xFFFFFu))];
        if (lane == 0) {
            match12[i1] = bidx; m2[bidx] = i1;
            if (checkOri) { const int bin = rot_bin(kp1.angle, s2.keys[bidx].angle); bin_of[i1] = bin; atomicAdd(&hist[bin], 1); }
        }
        __syncwarp();
    }
}

// ---- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:359-439), batched: one warp per map point ----
// Row i of the N x N distance matrix (diagonal 0) is histogrammed over its 257 possible values in shared memory; the reference's
// vDists[0.5 * (N - 1)] of the sorted row is the first value whose cumulative count exceeds that rank.  The first row with the smallest median wins.
__global__ void __launch_bounds__(128)
k_distinctive(int n_points, const int* __restrict__ offsets, const uint4* __restrict__ desc, int* __restrict__ best_idx) {
    __shared__ int hist_s[4][264];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int p = blockIdx.x * 4 + w;
    if (p >= n_points) return;
    int* hist = hist_s[w];
    const int o = offsets[p], N = offsets[p + 1] - o;
    if (N <= 0) { if (lane == 0) best_idx[p] = -1; return; }
    const int rank = (int)(0.5 * (double)(N - 1));
    int bestMedian = 0x7FFFFFFF, bestIdx = 0;
    for (int i = 0; i < N; ++i) {
        for (int b = lane; b < 264; b += 32) hist[b] = 0;
        __syncwarp();
        const uint4 d0 = __ldg(desc + 2 * (o + i)), d1 = __ldg(desc + 2 * (o + i) + 1);
        for (int j = lane; j < N; j += 32) atomicAdd(&hist[j == i ? 0 : hamming256(d0, d1, desc + 2 * (o + j))], 1);
        __syncwarp();
        // 257 bins, 9 per lane (lane 31 has fewer): exclusive prefix over the lanes, then the bin where the cumulative count passes the rank
        int loc[9], sum = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) { const int b = lane * 9 + k; loc[k] = b < 257 ? hist[b] : 0; sum += loc[k]; }
        int incl = sum;
#pragma unroll
        for (int s2 = 1; s2 < 32; s2 <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, s2); if (lane >= s2) incl += t; }
        int cum = incl - sum, med = 0x7FFFFFFF;
#pragma unroll
        for (int k = 0; k < 9; ++k) { cum += loc[k]; if (med == 0x7FFFFFFF && cum > rank) med = lane * 9 + k; }
        med = __reduce_min_sync(0xffffffffu, (unsigned)med);
        if (med < bestMedian) { bestMedian = med; bestIdx = i; }
        __syncwarp();
    }
    if (lane == 0) best_idx[p] = bestIdx;
}

// ComputeThreeMaxima (:1866-1908) + removal of the matches outside the three strongest rotation bins; counts the survivors
__global__ void __launch_bounds__(1024)
k_bow_finish(int n1, int checkOri, const int* __restrict__ hist, const int* __restrict__ bin_of, int* __restrict__ match12, int* __restrict__ match21, int* __restrict__ nmatches) {
    __shared__ int keep[3];
    __shared__ int total;
    if (threadIdx.x == 0) {
        total = 0;
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        if (checkOri) {
            for (int i = 0; i < M_HISTO; ++i) {
                const int s = hist[i];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
                else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
                else if (s > max3) { max3 = s; ind3 = i; }
            }
            if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; } else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
        }
        keep[0] = ind1; keep[1] = ind2; keep[2] = ind3;
    }
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < n1; i += blockDim.x) {
        const int j = match12[i];
        if (j < 0) continue;
        if (checkOri) { const int b = bin_of[i]; if (b != keep[0] && b != keep[1] && b != keep[2]) { match12[i] = -1; match21[j] = -1; continue; } }
        ++mine;
    }
    atomicAdd(&total, mine);
    __syncthreads();
    if (threadIdx.x == 0) *nmatches = total;
}
