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
__device__ __forceinline__ uint32_t octree_code(int x, int y, const LevelGeom& g) {
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

#define SORT_THREADS 256
#define SORT_WARPS (SORT_THREADS / 32)

// Shared-memory layout of the radix path for `cap` keys: key ping/pong (u32), index ping/pong (u16), per-warp digit
// histograms (u16 [SORT_WARPS][256]).
__host__ __device__ inline size_t octree_sort_smem_bytes(int cap) { return (size_t)cap * 12 + SORT_WARPS * 256 * 2 + 64; }

__global__ void __launch_bounds__(SORT_THREADS)
k_octree_sort(const LevelGeom* __restrict__ levels, const CellDesc* __restrict__ cells, int ncells,
              int slots_per_frame, int cand_per_frame, int nlevels, int smem_keys,
              const uint32_t* __restrict__ cand_slots, const uint16_t* __restrict__ cell_counts,
              uint32_t* __restrict__ ocand,          // [B][cand_per_frame] candidates in reference order
              unsigned long long* __restrict__ skey, // [B][cand_per_frame] sorted (code<<32 | index)
              uint32_t* __restrict__ spk,            // [B][cand_per_frame] packed candidate per sorted position
              int* __restrict__ ncand) {             // [B][nlevels]
    extern __shared__ __align__(16) uint8_t sort_sm[];
    __shared__ int warp_sums[SORT_WARPS];
    __shared__ int total_sm;
    const int level = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const LevelGeom& g = levels[level];
    const uint16_t* counts = cell_counts + (long long)b * ncells + g.cell_begin;
    const uint32_t* slots = cand_slots + (long long)b * slots_per_frame;
    uint32_t* oc = ocand + (long long)b * cand_per_frame + g.cand_off;
    unsigned long long* sk_g = skey + (long long)b * cand_per_frame + g.cand_off;
    uint32_t* sp = spk + (long long)b * cand_per_frame + g.cand_off;

    // ---- ordered gather: exclusive scan of the per-cell counts, each thread owns a run of cells ----
    const int per = (g.cell_count + SORT_THREADS - 1) / SORT_THREADS;
    const int c0 = min(tid * per, g.cell_count), c1 = min(c0 + per, g.cell_count);
    int mine = 0;
    for (int c = c0; c < c1; ++c) mine += counts[c];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (tid < 32) {
        int w = tid < SORT_WARPS ? warp_sums[tid] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, wi, o); if (tid >= o) wi += v; }
        if (tid < SORT_WARPS) warp_sums[tid] = wi - w;
        if (tid == SORT_WARPS - 1) total_sm = wi;
    }
    __syncthreads();
    const int n = total_sm;
    if (n > g.cand_cap) { if (tid == 0) ncand[b * nlevels + level] = -1; return; }   // cannot happen (cap is the worst case)
    const bool radix = n <= smem_keys;
    uint32_t* keyA = reinterpret_cast<uint32_t*>(sort_sm);
    uint32_t* keyB = keyA + smem_keys;
    uint16_t* idxA = reinterpret_cast<uint16_t*>(keyB + smem_keys);
    uint16_t* idxB = idxA + smem_keys;
    uint16_t* hist = idxB + smem_keys;                                   // [SORT_WARPS][256]
    int pos = warp_sums[warp] + incl - mine;
    for (int c = c0; c < c1; ++c) {                                      // ordered copy only (few threads own cells) ...
        const int cnt = counts[c];
        const uint32_t* src = slots + cells[g.cell_begin + c].slot;
        for (int k = 0; k < cnt; ++k, ++pos) oc[pos] = src[k];
    }
    __syncthreads();
    for (int i = tid; i < n; i += SORT_THREADS) {                        // ... the path codes are computed by all threads
        const uint32_t p = oc[i];
        const uint32_t code = octree_code((int)(p & 0xFFF), (int)((p >> 12) & 0xFFF), g);
        if (radix) { keyA[i] = code; idxA[i] = (uint16_t)i; }
        else sk_g[i] = ((unsigned long long)code << 32) | (unsigned)i;
    }
    __syncthreads();

    if (radix) {
        // ---- stable LSD radix sort on the path code, 8 bits per pass.  Digits below the depth at which a node is one
        // pixel wide are constant (both halves of a 1-px interval send every key left), so they are skipped. ----
        const int span = max(g.maxBX - g.minBX, g.maxBY - g.minBY);
        int nd = 1; while ((1 << nd) < span && nd < ORBX_MAXD) ++nd;     // an interval is one pixel wide after ceil(log2(span)) splits
        nd = min(nd + 2, ORBX_MAXD);                                     // + 2: keys the float root assignment put just outside their root interval settle one level later
        const int low = 2 * (ORBX_MAXD - nd);                            // constant low bits
        int rootbits = 0; while ((1 << rootbits) < g.nIni) ++rootbits;
        const int nbits = 2 * nd + rootbits;
        const int seg = (n + SORT_WARPS - 1) / SORT_WARPS;               // contiguous segment per warp keeps the sort stable
        const int s0 = min(warp * seg, n), s1 = min(s0 + seg, n);
        const uint32_t lt = (1u << lane) - 1u;
        uint16_t* myhist = hist + warp * 256;
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
            {   // offsets: digit-major, warp-minor exclusive scan; thread d owns digit d
                int tot = 0;
                int part[SORT_WARPS];
#pragma unroll
                for (int w = 0; w < SORT_WARPS; ++w) { part[w] = tot; tot += hist[w * 256 + tid]; }
                int inc = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
                if (lane == 31) warp_sums[warp] = inc;
                __syncthreads();
                int basew = 0;
#pragma unroll
                for (int w = 0; w < SORT_WARPS; ++w) if (w < warp) basew += warp_sums[w];
                const int excl = basew + inc - tot;
#pragma unroll
                for (int w = 0; w < SORT_WARPS; ++w) hist[w * 256 + tid] = (uint16_t)(excl + part[w]);
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
        }
        for (int i = tid; i < n; i += SORT_THREADS) {
            const unsigned oi = idxA[i];
            sk_g[i] = ((unsigned long long)keyA[i] << 32) | oi;
            sp[i] = oc[oi];
        }
        if (tid == 0) ncand[b * nlevels + level] = n;
        return;
    }

    // ---- oversized level: bitonic sort in global memory, all-ascending formulation (works for any n, no padding) ----
    unsigned long long* sk = sk_g;
    for (int k = 2; (k >> 1) < n; k <<= 1) {
        for (int i = tid; i < n; i += SORT_THREADS) {          // first substage: mirror partner
            const int j = i ^ (k - 1);
            if (j > i && j < n) { unsigned long long a = sk[i], c = sk[j]; if (a > c) { sk[i] = c; sk[j] = a; } }
        }
        __syncthreads();
        for (int s = k >> 2; s > 0; s >>= 1) {
            for (int i = tid; i < n; i += SORT_THREADS) {
                const int j = i ^ s;
                if (j > i && j < n) { unsigned long long a = sk[i], c = sk[j]; if (a > c) { sk[i] = c; sk[j] = a; } }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < n; i += SORT_THREADS) sp[i] = oc[(unsigned)(sk[i] & 0xFFFFFFFFull)];
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
