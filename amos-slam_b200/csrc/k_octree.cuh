// k_octree.cuh -- DistributeOctTree on the device (/root/reference/src/ORBextractor.cc:635-1049).
//
// Formulation.  DivideNode splits a node's x-interval and y-interval independently at
// lo + ceil((hi-lo)/2) and sends each key left/right by `pt < mid`; so every key has a fixed path
// through the quadtree that depends only on its coordinates and the root geometry.  We give each key a
// PATH CODE  root | d1 | d2 | ... | d13  (2 bits per depth: ybit*2 + xbit = child index n1..n4) and sort
// the level's keys by (code, original index) once.  Every node the reference can ever create is then a
// contiguous range of the sorted array, its children are the four sub-ranges split on the next digit,
// and "vKeys.size()" is the range length -- no key is ever moved again.
//   K3 k_octree_sort : one CTA per (level, frame): gather the cell slots in cell order (= the
//                      reference's vToDistributeKeys order), compute codes, bitonic-sort 64-bit
//                      (code<<32 | index) keys in shared memory (global memory for oversized levels).
//   K4 k_octree_tree : one warp per (level, frame) replays the reference's list surgery exactly
//                      (std::list push_front / erase order, the full pass, the "largest first" pass)
//                      on node records in shared memory; lanes cooperate on range searches and on the
//                      arg-max selection that replaces the sort at :948.  Tie-break for equal sizes is
//                      the canonical "newest node first" (creation id), see DESIGN.md.
// Output order = final std::list order, one keypoint per node: max response, first in original order.
#pragma once
#include "orbx_common.cuh"

// path code of a key; x,y are the integer coordinates relative to (minBorderX, minBorderY)
// the x half (root << 26 | x decisions on the even bits) and the y half (y decisions on the odd bits) of the code depend on one coordinate each:
// the host tabulates them per level (octree_code_tables below, same arithmetic), and a code is then two loads and an OR
__host__ __device__ inline uint32_t octree_code_half(int v, int lo, int hi, int bit) {
    uint32_t code = 0;
    for (int d = 0; d < ORBX_MAXD; ++d) {
        const int m = lo + ((hi - lo + 1) >> 1);
        const uint32_t b = v >= m;
        if (b) lo = m; else hi = m;
        code |= b << (2 * (ORBX_MAXD - 1 - d) + bit);
    }
    return code;
}
__device__ __forceinline__ uint32_t octree_code(int x, int y, const LevelGeom& g) {
    if (g.code_x) return g.code_x[min(x, g.code_nx - 1)] | g.code_y[min(y, g.code_ny - 1)];
    // root: vpIniNodes[kp.pt.x / hX]   (float division, truncation)        :766
    const int root = (int)__fdiv_rn((float)x, g.hX);
    // root geometry: UL.x = (int)(hX*i), UR.x = (int)(hX*(i+1)), y in [0, maxY-minY)   :741-745
    int xlo = (int)__fmul_rn(g.hX, (float)root), xhi = (int)__fmul_rn(g.hX, (float)(root + 1));
    int ylo = 0, yhi = g.maxBY - g.minBY;
    uint32_t code = (uint32_t)root;
#pragma unroll
    for (int d = 0; d < ORBX_MAXD; ++d) {
        const int xm = xlo + ((xhi - xlo + 1) >> 1);     // UL.x + ceil((UR.x-UL.x)/2)   :641
        const int ym = ylo + ((yhi - ylo + 1) >> 1);     // UL.y + ceil((BR.y-UL.y)/2)   :642
        const uint32_t xb = x >= xm, yb = y >= ym;        // kp.pt.x < n1.UR.x ... :680-690
        if (xb) xlo = xm; else xhi = xm;
        if (yb) ylo = ym; else yhi = ym;
        code = (code << 2) | (yb << 1) | xb;
    }
    return code;
}

// phase stamps of the (level 0, frame 0) instance for tools/qt_stamps_probe.py; compiled in only with -DORBX_QT_STAMPS (a probe build, never the product library)
#ifdef ORBX_QT_STAMPS
__device__ long long g_qt_stamps[64];
#define QT_STAMP(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && (i) < 64) g_qt_stamps[(i)] = clock64(); } while (0)
#else
#define QT_STAMP(i) do { } while (0)
#endif

#define SORT_THREADS 256
#define SORT_WARPS (SORT_THREADS / 32)

// Shared-memory layout of the radix path for `cap` keys: key ping/pong (u32), index ping/pong (u16), per-warp digit
// histograms (u16 [SORT_WARPS][256]).
__host__ __device__ inline size_t octree_sort_smem_bytes(int cap, int threads = SORT_THREADS) { return (size_t)cap * 12 + (size_t)(threads / 32) * 256 * 2 + 64; }
#define SORT_THREADS_WIDE 1024     // per-frame calls: one CTA per level, so a wide CTA shortens every pass of the level-0 sort

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_octree_sort_t(const LevelGeom* __restrict__ levels, const CellDesc* __restrict__ cells, int ncells,
              int slots_per_frame, int cand_per_frame, int nlevels, int smem_keys,
              const uint32_t* __restrict__ cand_slots, const uint16_t* __restrict__ cell_counts,
              uint32_t* __restrict__ ocand,          // [B][cand_per_frame] candidates in reference order
              unsigned long long* __restrict__ skey, // [B][cand_per_frame] sorted (code<<32 | index)
              uint32_t* __restrict__ spk,            // [B][cand_per_frame] packed candidate per sorted position
              int* __restrict__ ncand) {             // [B][nlevels]
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(16) uint8_t sort_sm[];
    __shared__ int warp_sums[32];
    __shared__ int total_sm;
    const int level = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const LevelGeom& g = levels[level];
    QT_STAMP(0);
    const uint16_t* counts = cell_counts + (long long)b * ncells + g.cell_begin;
    const uint32_t* slots = cand_slots + (long long)b * slots_per_frame;
    uint32_t* oc = ocand + (long long)b * cand_per_frame + g.cand_off;
    unsigned long long* sk_g = skey + (long long)b * cand_per_frame + g.cand_off;
    uint32_t* sp = spk + (long long)b * cand_per_frame + g.cand_off;

    // ---- ordered gather: exclusive scan of the per-cell counts, each thread owns a run of cells ----
    const int per = (g.cell_count + THREADS - 1) / THREADS;
    const int c0 = min(tid * per, g.cell_count), c1 = min(c0 + per, g.cell_count);
    int mine = 0;
    for (int c = c0; c < c1; ++c) mine += counts[c];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (tid < 32) {
        int w = tid < WARPS ? warp_sums[tid] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, wi, o); if (tid >= o) wi += v; }
        if (tid < WARPS) warp_sums[tid] = wi - w;
        if (tid == WARPS - 1) total_sm = wi;
    }
    __syncthreads();
    const int n = total_sm;
    QT_STAMP(1);
    if (n > g.cand_cap) { if (tid == 0) ncand[b * nlevels + level] = -1; return; }   // cannot happen (cap is the worst case)
    const bool radix = n <= smem_keys;
    uint32_t* keyA = reinterpret_cast<uint32_t*>(sort_sm);
    uint32_t* keyB = keyA + smem_keys;
    uint16_t* idxA = reinterpret_cast<uint16_t*>(keyB + smem_keys);
    uint16_t* idxB = idxA + smem_keys;
    uint16_t* hist = idxB + smem_keys;                                   // [WARPS][256]
    int pos = warp_sums[warp] + incl - mine;
    for (int c = c0; c < c1; ++c) {                                      // ordered copy only (few threads own cells) ...
        const int cnt = counts[c];
        const uint32_t* src = slots + cells[g.cell_begin + c].slot;
        for (int k = 0; k < cnt; ++k, ++pos) oc[pos] = src[k];
    }
    __syncthreads();
    QT_STAMP(2);
    for (int i = tid; i < n; i += THREADS) {                        // ... the path codes are computed by all threads
        const uint32_t p = oc[i];
        const uint32_t code = octree_code((int)(p & 0xFFF), (int)((p >> 12) & 0xFFF), g);
        if (radix) { keyA[i] = code; idxA[i] = (uint16_t)i; }
        else sk_g[i] = ((unsigned long long)code << 32) | (unsigned)i;
    }
    __syncthreads();
    QT_STAMP(3);

    if (radix) {
        // ---- stable LSD radix sort on the path code, 8 bits per pass.  Digits below the depth at which a node is one
        // pixel wide are constant (both halves of a 1-px interval send every key left), so they are skipped. ----
        const int span = max(g.maxBX - g.minBX, g.maxBY - g.minBY);
        int nd = 1; while ((1 << nd) < span && nd < ORBX_MAXD) ++nd;     // an interval is one pixel wide after ceil(log2(span)) splits
        nd = min(nd + 2, ORBX_MAXD);                                     // + 2: keys the float root assignment put just outside their root interval settle one level later
        const int low = 2 * (ORBX_MAXD - nd);                            // constant low bits
        int rootbits = 0; while ((1 << rootbits) < g.nIni) ++rootbits;
        const int nbits = 2 * nd + rootbits;
        const int seg = (n + WARPS - 1) / WARPS;               // contiguous segment per warp keeps the sort stable
        const int s0 = min(warp * seg, n), s1 = min(s0 + seg, n);
        const uint32_t lt = (1u << lane) - 1u;
        uint16_t* myhist = hist + warp * 256;
        int qt_pass = 0; (void)qt_pass;
        for (int shift = low; shift < low + nbits; shift += 8) {
            for (int i = lane; i < 256; i += 32) myhist[i] = 0;
            __syncwarp();
            for (int i0 = s0; i0 < s1; i0 += 32) {
                const int i = i0 + lane;
                const bool act = i < s1;
                const uint32_t d = act ? ((keyA[i] >> shift) & 255u) : 256u;
                const uint32_t peers = __match_any_sync(0xffffffffu, d);
                if (act && (peers & lt) == 0) myhist[d] += (uint16_t)__popc(peers);      // one leader per digit value
                __syncwarp();
            }
            __syncthreads();
            {   // offsets: digit-major, warp-minor exclusive scan; thread d < 256 owns digit d (the first 8 warps)
                int tot = 0, inc = 0;
                int part[WARPS];
                if (tid < 256) {
#pragma unroll
                    for (int w = 0; w < WARPS; ++w) { part[w] = tot; tot += hist[w * 256 + tid]; }
                    inc = tot;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
                    if (lane == 31) warp_sums[warp] = inc;
                }
                __syncthreads();
                if (tid < 256) {
                    int basew = 0;
#pragma unroll
                    for (int w = 0; w < 8; ++w) if (w < warp) basew += warp_sums[w];
                    const int excl = basew + inc - tot;
#pragma unroll
                    for (int w = 0; w < WARPS; ++w) hist[w * 256 + tid] = (uint16_t)(excl + part[w]);
                }
            }
            __syncthreads();
            for (int i0 = s0; i0 < s1; i0 += 32) {
                const int i = i0 + lane;
                const bool act = i < s1;
                const uint32_t k = act ? keyA[i] : 0u;
                const uint32_t d = act ? ((k >> shift) & 255u) : 256u;
                const uint32_t peers = __match_any_sync(0xffffffffu, d);
                int dst = 0;
                if (act) dst = myhist[d] + __popc(peers & lt);
                __syncwarp();
                if (act && (peers & lt) == 0) myhist[d] += (uint16_t)__popc(peers);
                if (act) { keyB[dst] = k; idxB[dst] = idxA[i]; }
                __syncwarp();
            }
            __syncthreads();
            uint32_t* tk = keyA; keyA = keyB; keyB = tk;
            uint16_t* ti = idxA; idxA = idxB; idxB = ti;
            QT_STAMP(4 + qt_pass); ++qt_pass;
        }
        for (int i = tid; i < n; i += THREADS) {
            const unsigned oi = idxA[i];
            sk_g[i] = ((unsigned long long)keyA[i] << 32) | oi;
            sp[i] = oc[oi];
        }
        QT_STAMP(10);
        if (tid == 0) ncand[b * nlevels + level] = n;
        return;
    }

    // ---- oversized level: bitonic sort in global memory, all-ascending formulation (works for any n, no padding) ----
    unsigned long long* sk = sk_g;
    for (int k = 2; (k >> 1) < n; k <<= 1) {
        for (int i = tid; i < n; i += THREADS) {          // first substage: mirror partner
            const int j = i ^ (k - 1);
            if (j > i && j < n) { unsigned long long a = sk[i], c = sk[j]; if (a > c) { sk[i] = c; sk[j] = a; } }
        }
        __syncthreads();
        for (int s = k >> 2; s > 0; s >>= 1) {
            for (int i = tid; i < n; i += THREADS) {
                const int j = i ^ s;
                if (j > i && j < n) { unsigned long long a = sk[i], c = sk[j]; if (a > c) { sk[i] = c; sk[j] = a; } }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < n; i += THREADS) sp[i] = oc[(unsigned)(sk[i] & 0xFFFFFFFFull)];
    if (tid == 0) ncand[b * nlevels + level] = n;
}

// ---- warp-cooperative lower bound on the sorted path codes: first index in [lo,hi) with code >= T ----
// codes are read as ck[pos * cs]: cs = 1 for the shared-memory copy, 2 for the high words of the 64-bit global keys
__device__ __forceinline__ int warp_lower_bound(const uint32_t* __restrict__ ck, int cs, int lo, int hi, uint32_t T, int lane) {
    while (hi - lo > 32) {
        const int step = (hi - lo + 31) >> 5;
        const int pos = lo + lane * step;
        const bool pred = pos < hi && ck[pos * cs] < T;
        const int c = __popc(__ballot_sync(0xffffffffu, pred));
        if (c == 0) return lo;
        const int nlo = lo + (c - 1) * step + 1;
        hi = min(lo + c * step, hi);
        lo = nlo;
    }
    const int pos = lo + lane;
    const bool pred = pos < hi && ck[pos * cs] < T;
    return lo + __popc(__ballot_sync(0xffffffffu, pred));
}

struct TreeSmem {
    int* lo; int* hi; unsigned* cid; short* nxt; short* prv; unsigned char* dep; short* freeStack;
    unsigned long long* vcur; unsigned long long* vprev;
};

__global__ void __launch_bounds__(32)
k_octree_tree(const LevelGeom* __restrict__ levels, int nlevels, int cand_per_frame, int kp_per_frame, int cap_max, int code_cap,
              const unsigned long long* __restrict__ skey, const uint32_t* __restrict__ spk, const int* __restrict__ ncand,
              uint32_t* __restrict__ kp_level,      // [B][kp_per_frame] packed x:12|y:12|resp:8 (relative to minBorder)
              int* __restrict__ kp_count,           // [B][nlevels]
              int* __restrict__ overflow) {
    extern __shared__ __align__(16) uint8_t tree_sm[];
    const int level = blockIdx.x, b = blockIdx.y, lane = threadIdx.x;
    const LevelGeom& g = levels[level];
    const int n = ncand[b * nlevels + level];
    const unsigned long long* key = skey + (long long)b * cand_per_frame + g.cand_off;
    const uint32_t* pk = spk + (long long)b * cand_per_frame + g.cand_off;
    uint32_t* out = kp_level + (long long)b * kp_per_frame + g.kp_off;
    if (n <= 0) { if (lane == 0) { kp_count[b * nlevels + level] = 0; if (n < 0) atomicOr(overflow, ORBX_OVF_SORT); } return; }

    const int cap = cap_max;     // node slots available (>= kp_cap + 8 for every level)
    // the serial replay probes the sorted codes thousands of times: keep them in shared memory (after the node records)
    // whenever the level's candidates fit; otherwise read the high words of the global 64-bit keys
    uint32_t* codes_sm = reinterpret_cast<uint32_t*>(tree_sm + (((size_t)cap * (8 + 8 + 4 + 4 + 4 + 2 + 2 + 2 + 1) + 15) & ~(size_t)15));
    const uint32_t* ck; int cs;
    if (n <= code_cap) {
        for (int i = lane; i < n; i += 32) codes_sm[i] = (uint32_t)(key[i] >> 32);
        ck = codes_sm; cs = 1;
    } else { ck = reinterpret_cast<const uint32_t*>(key) + 1; cs = 2; }
    __syncwarp();
    unsigned long long* vcur = reinterpret_cast<unsigned long long*>(tree_sm);
    unsigned long long* vprev = vcur + cap;
    int* nlo = reinterpret_cast<int*>(vprev + cap);
    int* nhi = nlo + cap;
    unsigned* ncid = reinterpret_cast<unsigned*>(nhi + cap);
    short* nnxt = reinterpret_cast<short*>(ncid + cap);
    short* nprv = nnxt + cap;
    short* freeStack = nprv + cap;
    unsigned char* ndep = reinterpret_cast<unsigned char*>(freeStack + cap);

    const int N = g.N;
    int head = -1, tail = -1, count = 0, nfree = cap;
    unsigned nextCid = 0;
    bool ovf = false;
    for (int i = lane; i < cap; i += 32) freeStack[i] = (short)(cap - 1 - i);
    __syncwarp();

    // all lanes run the same scalar control flow; lane 0 performs the shared-memory writes
#define NODE_ALLOC(slot) do { if (nfree == 0) { ovf = true; slot = -1; } else { slot = freeStack[--nfree]; } } while (0)
#define NODE_FREE(slot) do { if (lane == 0) freeStack[nfree] = (short)(slot); ++nfree; } while (0)

    // ---- initial nodes (:733-788): nIni roots in order (push_back), empty ones erased ----
    for (int i = 0; i < g.nIni; ++i) {
        const int rlo = warp_lower_bound(ck, cs, 0, n, (uint32_t)i << ORBX_ROOT_SHIFT, lane);
        const int rhi = warp_lower_bound(ck, cs, rlo, n, (uint32_t)(i + 1) << ORBX_ROOT_SHIFT, lane);
        ++nextCid;
        if (rhi > rlo) {
            int s; NODE_ALLOC(s);
            if (s < 0) break;
            if (lane == 0) { nlo[s] = rlo; nhi[s] = rhi; ncid[s] = nextCid - 1; ndep[s] = 0; nnxt[s] = -1; nprv[s] = (short)tail; if (tail >= 0) nnxt[tail] = (short)s; }
            if (head < 0) head = s;
            tail = s; ++count;
            __syncwarp();
        }
    }

    int nv = 0;   // entries in vcur
    // divide node `slot`: push its non-empty children to the FRONT in order n1..n4, record children with
    // more than one key in vcur (size<<40 | cid<<16 | slot), erase the node.   (:839-895 / :952-998)
    auto divide = [&](int slot, int& nToExpand) {
        const int lo = nlo[slot], hi = nhi[slot], d = ndep[slot];
        const int shift = 2 * (ORBX_MAXD - d - 1);
        int bnd[5];
        bnd[0] = lo; bnd[4] = hi;
        if (hi - lo <= 32) {
            const int pos = lo + lane;
            const unsigned digit = pos < hi ? ((ck[pos * cs] >> shift) & 3u) : 4u;
            bnd[1] = lo + __popc(__ballot_sync(0xffffffffu, digit < 1u));
            bnd[2] = lo + __popc(__ballot_sync(0xffffffffu, digit < 2u));
            bnd[3] = lo + __popc(__ballot_sync(0xffffffffu, digit < 3u));
        } else {
            const uint32_t prefix = ck[lo * cs] >> (shift + 2);
            bnd[1] = warp_lower_bound(ck, cs, lo, hi, ((prefix << 2) | 1u) << shift, lane);
            bnd[2] = warp_lower_bound(ck, cs, bnd[1], hi, ((prefix << 2) | 2u) << shift, lane);
            bnd[3] = warp_lower_bound(ck, cs, bnd[2], hi, ((prefix << 2) | 3u) << shift, lane);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int sz = bnd[k + 1] - bnd[k];
            if (sz > 0) {
                int s; NODE_ALLOC(s);
                if (s < 0) return;
                const unsigned id = nextCid++;
                if (lane == 0) {
                    nlo[s] = bnd[k]; nhi[s] = bnd[k + 1]; ncid[s] = id; ndep[s] = (unsigned char)(d + 1);
                    nprv[s] = -1; nnxt[s] = (short)head; if (head >= 0) nprv[head] = (short)s;
                    if (sz > 1) vcur[nv] = ((unsigned long long)sz << 40) | ((unsigned long long)(id & 0xFFFFFFu) << 16) | (unsigned)s;
                }
                head = s; if (tail < 0) tail = s;
                ++count;
                if (sz > 1) { ++nToExpand; ++nv; }
                __syncwarp();
            }
        }
        // erase(slot)
        const int p = nprv[slot], q = nnxt[slot];
        if (lane == 0) { if (p >= 0) nnxt[p] = (short)q; if (q >= 0) nprv[q] = (short)p; }
        if (p < 0) head = q;
        if (q < 0) tail = p;
        --count;
        NODE_FREE(slot);
        __syncwarp();
    };

    bool bFinish = false;
    while (!bFinish && !ovf) {
        const int prevSize = count;
        int nToExpand = 0;
        nv = 0;
        int cur = head;
        while (cur >= 0 && !ovf) {                                   // full pass (:824-899)
            const int nx = nnxt[cur];
            const int sz = nhi[cur] - nlo[cur];
            if (sz != 1 && ndep[cur] < ORBX_MAXD) divide(cur, nToExpand);   // bNoMore <=> one key
            else if (sz != 1) ovf = true;
            cur = nx;
        }
        if (count >= N || count == prevSize) bFinish = true;          // :907
        else if (count + nToExpand * 3 > N) {                          // :929
            while (!bFinish && !ovf) {
                const int prevSize2 = count;
                unsigned long long* t = vprev; vprev = vcur; vcur = t;
                const int np = nv; nv = 0;
                for (int it = 0; it < np; ++it) {
                    // largest size first, newest first among equals (:948-950, canonical tie-break)
                    unsigned long long best = 0; int bi = -1;
                    for (int e = lane; e < np; e += 32) { const unsigned long long v = vprev[e]; if (v > best) { best = v; bi = e; } }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const unsigned long long ov = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (ov > best) { best = ov; bi = oi; }
                    }
                    if (lane == 0) vprev[bi] = 0ull;
                    __syncwarp();
                    const int slot = (int)(best & 0xFFFFull);
                    if (ndep[slot] >= ORBX_MAXD) { ovf = true; break; }
                    int dummy = 0;
                    divide(slot, dummy);
                    if (count >= N || ovf) break;                      // :1003
                }
                if (count >= N || count == prevSize2) bFinish = true; // :1009
            }
        }
    }

    // ---- result: one keypoint per node in list order (:1018-1048) ----
    short* order = reinterpret_cast<short*>(vprev);
    if (lane == 0) { int k = 0; for (int c = head; c >= 0 && k < cap; c = nnxt[c]) order[k++] = (short)c; }
    __syncwarp();
    const int nout = min(count, g.kp_cap);
    if (count > g.kp_cap) ovf = true;
    for (int k = lane; k < nout; k += 32) {
        const int s = order[k];
        const int lo = nlo[s], hi = nhi[s];
        uint32_t bestp = 0; int bestr = -1; unsigned besti = 0xFFFFFFFFu;
        for (int i = lo; i < hi; ++i) {
            const uint32_t p = pk[i];
            const int r = (int)(p >> 24);
            const unsigned oi = (unsigned)(key[i] & 0xFFFFFFFFull);
            if (r > bestr || (r == bestr && oi < besti)) { bestr = r; besti = oi; bestp = p; }   // max response, first wins
        }
        out[k] = bestp;
    }
    if (lane == 0) {
        kp_count[b * nlevels + level] = nout;
        if (ovf) atomicOr(overflow, ORBX_OVF_TREE);
    }
#undef NODE_ALLOC
#undef NODE_FREE
}

// =================================================================================================
// K4'  k_octree_tree_par: the same replay as k_octree_tree, reorganised as a handful of block-wide rounds.
// The reference's list surgery is deterministic given the round structure:
//   * a "full pass" (:824-899) divides EVERY multi-key node of the list snapshot.  Children are push_front'ed in order
//     n1..n4, so afterwards the list is: for the divided nodes in REVERSE visiting order, their non-empty children in
//     REVERSE child order; then the undivided nodes in their old order.  Creation ids follow visiting order.
//   * a "largest first" round (:929-1011) visits the candidates of the previous round by (size, creation id) descending
//     and stops right after the divide that brings the node count to >= N.  With the per-candidate child counts known,
//     the stopping point is a prefix sum away, and the list afterwards has the same closed form as above.
// So each round is: one thread per node computes the child boundaries (binary searches on the sorted path codes),
// block-wide prefix sums give list positions / creation ids / candidate slots, and a scatter writes the next list.
// A level of the C1 workload takes ~6 rounds instead of ~100 dependent divides by a single warp.
// =================================================================================================
#define PTREE_THREADS 128
#define PTREE_MAXCAP 1024      // node capacity handled by this kernel (levels asking for more use the serial kernel)

__host__ __device__ inline size_t ptree_smem_bytes(int cap, int code_cap) {
    // codes | 2 x (lo, hi, cid: int) + 2 x dep (short) | b1 b2 b3 sM sE sU (int) | vcur vprev (u64) | order (int)
    return (size_t)code_cap * 4 + (size_t)cap * (2 * 12 + 2 * 2 + 6 * 4 + 2 * 8 + 4) + 64;
}

__device__ __forceinline__ int ptree_lower_bound(const uint32_t* __restrict__ ck, int cs, int lo, int hi, uint32_t T) {
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (ck[mid * cs] < T) lo = mid + 1; else hi = mid; }
    return lo;
}

// inclusive prefix sum of a[0..n) in shared memory, n <= 1024; every thread of the block calls it
__device__ __forceinline__ void ptree_scan(int* a, int n, int* seg) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nseg = (n + 31) >> 5;
    for (int s = warp; s < nseg; s += PTREE_THREADS / 32) {
        const int i = s * 32 + lane;
        int v = i < n ? a[i] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        if (i < n) a[i] = v;
        if (lane == 31) seg[s] = v;
    }
    __syncthreads();
    if (warp == 0) {
        int v = lane < nseg ? seg[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        seg[lane] = v;
    }
    __syncthreads();
    for (int i = tid; i < n; i += PTREE_THREADS) { const int s = i >> 5; if (s > 0) a[i] += seg[s - 1]; }
    __syncthreads();
}

__global__ void __launch_bounds__(PTREE_THREADS)
k_octree_tree_par(const LevelGeom* __restrict__ levels, int nlevels, int cand_per_frame, int kp_per_frame, int cap, int code_cap,
                  const unsigned long long* __restrict__ skey, const uint32_t* __restrict__ spk, const int* __restrict__ ncand,
                  uint32_t* __restrict__ kp_level, int* __restrict__ kp_count, int* __restrict__ overflow) {
    extern __shared__ __align__(16) uint8_t pt_sm[];
    __shared__ int seg[32];
    __shared__ int s_ovf, s_rstar;
    const int level = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const LevelGeom& g = levels[level];
    QT_STAMP(16);
    const int n = ncand[b * nlevels + level];
    const unsigned long long* key = skey + (long long)b * cand_per_frame + g.cand_off;
    const uint32_t* pk = spk + (long long)b * cand_per_frame + g.cand_off;
    uint32_t* out = kp_level + (long long)b * kp_per_frame + g.kp_off;
    if (n <= 0) { if (tid == 0) { kp_count[b * nlevels + level] = 0; if (n < 0) atomicOr(overflow, ORBX_OVF_SORT); } return; }

    uint32_t* codes_sm = reinterpret_cast<uint32_t*>(pt_sm);
    int* base = reinterpret_cast<int*>(codes_sm + code_cap);
    int* nlo[2] = {base, base + cap}; int* nhi[2] = {base + 2 * cap, base + 3 * cap}; int* ncid[2] = {base + 4 * cap, base + 5 * cap};
    int* b1 = base + 6 * cap; int* b2 = b1 + cap; int* b3 = b2 + cap; int* sM = b3 + cap; int* sE = sM + cap; int* sU = sE + cap; int* order = sU + cap;
    unsigned long long* vA = reinterpret_cast<unsigned long long*>(order + cap + (cap & 1));
    unsigned long long* vB = vA + cap;
    short* ndep[2] = {reinterpret_cast<short*>(vB + cap), reinterpret_cast<short*>(vB + cap) + cap};

    const uint32_t* ck; int cs;
    if (n <= code_cap) {
        for (int i = tid; i < n; i += PTREE_THREADS) codes_sm[i] = (uint32_t)(key[i] >> 32);
        ck = codes_sm; cs = 1;
    } else { ck = reinterpret_cast<const uint32_t*>(key) + 1; cs = 2; }
    if (tid == 0) s_ovf = 0;
    __syncthreads();
    QT_STAMP(17);
    int qt_round = 0; (void)qt_round;

    const int N = g.N;
    int cur = 0;                               // node buffer holding the current list, in list order
    unsigned long long* vcur = vA; unsigned long long* vprev = vB;
    // child boundaries of node (lo, hi, dep): c[0..4]; returns the number of non-empty children, *multi = those with > 1 key
    auto children = [&](int lo, int hi, int dep, int& c1, int& c2, int& c3, int& multi) -> int {
        const int shift = 2 * (ORBX_MAXD - dep - 1);
        const uint32_t prefix = ck[lo * cs] >> (shift + 2);
        c1 = ptree_lower_bound(ck, cs, lo, hi, ((prefix << 2) | 1u) << shift);
        c2 = ptree_lower_bound(ck, cs, c1, hi, ((prefix << 2) | 2u) << shift);
        c3 = ptree_lower_bound(ck, cs, c2, hi, ((prefix << 2) | 3u) << shift);
        const int s0 = c1 - lo, s1 = c2 - c1, s2 = c3 - c2, s3 = hi - c3;
        multi = (s0 > 1) + (s1 > 1) + (s2 > 1) + (s3 > 1);
        return (s0 > 0) + (s1 > 0) + (s2 > 0) + (s3 > 0);
    };
    // scatter the children of one divided node into list `nx`: non-empty children in order n1..n4 get creation ids cid0, cid0+1, ...
    // and list positions blockstart + m-1, blockstart + m-2, ...; children with > 1 key are appended to vcur from slot e0
    auto emit_children = [&](int nx, int lo, int c1, int c2, int c3, int hi, int dep, int m, int blockstart, int cid0, int e0) {
        const int bnd[5] = {lo, c1, c2, c3, hi};
        int q = 0, t = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int sz = bnd[k + 1] - bnd[k];
            if (sz > 0) {
                const int pos = blockstart + (m - 1 - q);
                nlo[nx][pos] = bnd[k]; nhi[nx][pos] = bnd[k + 1]; ncid[nx][pos] = cid0 + q; ndep[nx][pos] = (short)(dep + 1);
                if (sz > 1) { vcur[e0 + t] = ((unsigned long long)sz << 40) | ((unsigned long long)((unsigned)(cid0 + q) & 0xFFFFFFu) << 16) | (unsigned)pos; ++t; }
                ++q;
            }
        }
    };

    // ---- initial nodes (:733-788): nIni roots in order, empty ones dropped; creation ids 0..nIni-1 ----
    for (int i = tid; i < g.nIni; i += PTREE_THREADS) {
        const int rlo = ptree_lower_bound(ck, cs, 0, n, (uint32_t)i << ORBX_ROOT_SHIFT);
        const int rhi = ptree_lower_bound(ck, cs, rlo, n, (uint32_t)(i + 1) << ORBX_ROOT_SHIFT);
        b1[i] = rlo; b2[i] = rhi; sM[i] = rhi > rlo ? 1 : 0;
    }
    __syncthreads();
    ptree_scan(sM, g.nIni, seg);
    int count = sM[g.nIni - 1], nextCid = g.nIni, nv = 0;
    for (int i = tid; i < g.nIni; i += PTREE_THREADS)
        if (b2[i] > b1[i]) { const int pos = sM[i] - 1; nlo[0][pos] = b1[i]; nhi[0][pos] = b2[i]; ncid[0][pos] = i; ndep[0][pos] = 0; }
    __syncthreads();

    QT_STAMP(18);
    bool finish = false;
    while (!finish) {
        // ================= full pass =================
        const int prevSize = count;
        for (int i = tid; i < count; i += PTREE_THREADS) {
            const int lo = nlo[cur][i], hi = nhi[cur][i], dep = ndep[cur][i];
            int m = 0, e = 0, c1 = lo, c2 = lo, c3 = lo;
            if (hi - lo > 1) {
                if (dep >= ORBX_MAXD) s_ovf = 1;
                else m = children(lo, hi, dep, c1, c2, c3, e);
            }
            b1[i] = c1; b2[i] = c2; b3[i] = c3; sM[i] = m; sE[i] = e; sU[i] = m == 0 ? 1 : 0;
        }
        __syncthreads();
        if (s_ovf) break;
        ptree_scan(sM, count, seg); ptree_scan(sE, count, seg); ptree_scan(sU, count, seg);
        const int F = sM[count - 1], Etot = sE[count - 1], newcount = F + sU[count - 1];
        if (newcount > cap) { if (tid == 0) s_ovf = 1; __syncthreads(); break; }
        const int nx = cur ^ 1;
        for (int i = tid; i < count; i += PTREE_THREADS) {
            const int lo = nlo[cur][i], hi = nhi[cur][i], dep = ndep[cur][i];
            const int m = sM[i] - (i ? sM[i - 1] : 0);
            if (m > 0) {
                const int e = sE[i] - (i ? sE[i - 1] : 0);
                emit_children(nx, lo, b1[i], b2[i], b3[i], hi, dep, m, F - sM[i], nextCid + sM[i] - m, sE[i] - e);
            } else {
                const int pos = F + sU[i] - 1;
                nlo[nx][pos] = lo; nhi[nx][pos] = hi; ncid[nx][pos] = ncid[cur][i]; ndep[nx][pos] = (short)dep;
            }
        }
        __syncthreads();
        cur = nx; count = newcount; nextCid += F; nv = Etot;
        QT_STAMP(20 + qt_round); ++qt_round;
        const int nToExpand = Etot;
        if (count >= N || count == prevSize) { finish = true; break; }          // :907
        if (count + nToExpand * 3 > N) {                                          // :929
            // ================= "largest first" rounds =================
            while (!finish) {
                const int prevSize2 = count;
                { unsigned long long* t = vprev; vprev = vcur; vcur = t; }
                const int np = nv; nv = 0;
                if (np == 0) { finish = true; break; }                             // nothing left to divide: size unchanged (:1009)
                // rank of every candidate by (size, creation id) descending (:948-950, canonical tie-break)
                for (int e = tid; e < np; e += PTREE_THREADS) {
                    const unsigned long long v = vprev[e];
                    int r = 0;
                    for (int f = 0; f < np; ++f) r += vprev[f] > v;
                    order[r] = e;
                }
                if (tid == 0) s_rstar = np - 1;
                __syncthreads();
                for (int r = tid; r < np; r += PTREE_THREADS) {
                    const int pos = (int)(vprev[order[r]] & 0xFFFFull);
                    const int lo = nlo[cur][pos], hi = nhi[cur][pos], dep = ndep[cur][pos];
                    int m = 1, e = 0, c1 = lo, c2 = lo, c3 = lo;
                    if (dep >= ORBX_MAXD) s_ovf = 1;
                    else m = children(lo, hi, dep, c1, c2, c3, e);
                    b1[r] = c1; b2[r] = c2; b3[r] = c3; sM[r] = m; sE[r] = e;
                }
                for (int i = tid; i < count; i += PTREE_THREADS) sU[i] = 1;
                __syncthreads();
                if (s_ovf) break;
                ptree_scan(sM, np, seg); ptree_scan(sE, np, seg);
                // first candidate (in visiting order) after whose divide the list holds >= N nodes (:1003)
                for (int r = tid; r < np; r += PTREE_THREADS)
                    if (count + sM[r] - (r + 1) >= N) atomicMin(&s_rstar, r);
                __syncthreads();
                const int rstar = s_rstar, P = rstar + 1, F = sM[rstar], Etot = sE[rstar], newcount = count + F - P;
                if (newcount > cap) { if (tid == 0) s_ovf = 1; __syncthreads(); break; }
                for (int r = tid; r < P; r += PTREE_THREADS) sU[(int)(vprev[order[r]] & 0xFFFFull)] = 0;      // visited nodes leave the list
                __syncthreads();
                ptree_scan(sU, count, seg);
                const int nx = cur ^ 1;
                for (int r = tid; r < P; r += PTREE_THREADS) {
                    const int pos = (int)(vprev[order[r]] & 0xFFFFull);
                    const int m = sM[r] - (r ? sM[r - 1] : 0), e = sE[r] - (r ? sE[r - 1] : 0);
                    emit_children(nx, nlo[cur][pos], b1[r], b2[r], b3[r], nhi[cur][pos], ndep[cur][pos], m, F - sM[r], nextCid + sM[r] - m, sE[r] - e);
                }
                for (int i = tid; i < count; i += PTREE_THREADS) {
                    const bool stays = (sU[i] - (i ? sU[i - 1] : 0)) != 0;
                    if (stays) { const int pos = F + sU[i] - 1; nlo[nx][pos] = nlo[cur][i]; nhi[nx][pos] = nhi[cur][i]; ncid[nx][pos] = ncid[cur][i]; ndep[nx][pos] = ndep[cur][i]; }
                }
                __syncthreads();
                cur = nx; count = newcount; nextCid += F; nv = Etot;
                QT_STAMP(20 + qt_round); ++qt_round;
                if (count >= N || count == prevSize2) finish = true;               // :1009
            }
            break;
        }
    }
    __syncthreads();
    const bool ovf = s_ovf != 0;
    QT_STAMP(38);

    // ---- result: one keypoint per node in list order (:1018-1048): max response, first in the original order ----
    const int nout = min(count, g.kp_cap);
    for (int k = tid; k < nout; k += PTREE_THREADS) {
        const int lo = nlo[cur][k], hi = nhi[cur][k];
        uint32_t bestp = 0; int bestr = -1; unsigned besti = 0xFFFFFFFFu;
        for (int i = lo; i < hi; ++i) {
            const uint32_t p = pk[i];
            const int r = (int)(p >> 24);
            const unsigned oi = (unsigned)(key[i] & 0xFFFFFFFFull);
            if (r > bestr || (r == bestr && oi < besti)) { bestr = r; besti = oi; bestp = p; }
        }
        out[k] = bestp;
    }
    QT_STAMP(39);
    if (tid == 0) {
        kp_count[b * nlevels + level] = nout;
        if (ovf || count > g.kp_cap) atomicOr(overflow, ORBX_OVF_TREE);
    }
}
