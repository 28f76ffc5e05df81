// k_octree_fused.cuh -- DistributeOctTree (/root/reference/src/ORBextractor.cc:635-1049) as ONE kernel.  One CTA per (level, frame).
//   * wide form (QF_THREADS, a handful of frames: what Tracking calls once per frame): gather, path codes, sort and the tree all stay in shared memory, so the chain
//     k_octree_sort -> skey / spk / ocand in HBM -> k_octree_tree_par  (two launches, ~130 block barriers, 39 us for a 640 x 480 level 0 with its 2.5 k
//     candidates) becomes one launch of 17 us (phase times: tools/qt_stamps_probe.py);
//   * lean form (QF_THREADS_BATCH, batches): only small levels keep their keys in shared memory, the others run the same code on L2-resident global scratch -- more
//     resident CTAs hide the latencies better than shared-memory keys do (0.59 -> 0.40 ms per 1024 VGA frames).
// Levels with more candidates than the shared-memory key capacity of either form take the global-scratch path too (32-bit index and radix counters).
//
// The tree is not replayed round by round any more.  With the keys sorted by path code (k_octree.cuh), let dd[i] be the first digit in
// which key i differs from key i-1 (0 = different root, 1..13 = tree digit, dd[0] = dd[n] = 0).  Then:
//   * FULL PASSES (:824-899) divide EVERY multi-key node, so after q passes the nodes are exactly the maximal runs between the positions
//     with dd <= q: the node count C(q) and the number of multi-key nodes M(q) = C(q) - #{i : max(dd[i], dd[i+1]) <= q} are two cumulative
//     histograms, and the reference's loop control (:907 count >= N or unchanged -> finish; :929 count + 3 nToExpand > N -> largest-first)
//     picks the number of passes Q from them in one go.  A node's depth D is Q if it still has several keys, else max(dd[lo], dd[lo+1]).
//   * LIST ORDER.  Children are push_front'ed in order n1..n4 while the list is walked from the front, so after a pass the list reads
//     (children in REVERSE creation order) + (undivided nodes in their old order).  Unrolled over the passes: the nodes created at depth D
//     form one block, sorted by (root, d1, .., dD) with the last digit descending and the direction of every earlier digit flipped once per
//     later pass -- i.e. ascending order of  prefix ^ xormask(D)  (qf_tkey) --, and the blocks follow each other newest first.
//   * LARGEST-FIRST ROUNDS (:929-1011) visit the multi-key nodes created by the previous round by (size desc, newest first) and stop at
//     the divide that reaches N.  "Newest first" is the list order of the newest block (reverse creation order), so no creation ids are
//     needed: rank by (size desc, block key asc), prefix-sum the growth in rank order, cut at the first rank that reaches N.  The children
//     form the next block with key (rank desc, child digit desc).
//   * The final list is the living nodes sorted by (block age, block key); each emits its best key (max response, first in the original
//     order, :1018-1048).
// Everything order-dependent is therefore a sort key; no list is ever walked.  Equal to k_octree_tree / k_octree_tree_par (and to the
// reference build under the canonical tie-break) on every golden and adversarial input of tests/.
#pragma once
#include "k_octree.cuh"

#define QF_BUCKET_BITS 11
#define QF_BUCKETS (1 << QF_BUCKET_BITS)
#define QF_BUCKET_MAX 48            // largest bucket the in-bucket ranking takes (n / 64 for large levels); above it the radix passes sort the level
#define QF_THREADS 1024             // latency form: one wide CTA per level, alone on its SM
#define QF_THREADS_BATCH 256        // batched form: lean CTAs, several per SM
// u16 per digit row of the radix histogram: one counter per warp + 8 padding (1024 threads: 80 B rows, conflict-free 16-byte reads)
__host__ __device__ constexpr int qf_hs(int threads) { return threads / 32 + 8; }
#define QF_MAXPOOL 2048             // node pool entries (N + 3 per level at most): more features per level than this fall back to the sort + tree pair

#define QF_MAXLEVELS 16
struct QfLevels { LevelGeom lv[QF_MAXLEVELS]; };   // the level geometry travels as a kernel parameter: no dependent global load in front of the counts
struct QfPlan { int key_cap, pool_cap, cell_cap, tab_cap, threads, smem_bytes; };   // tab_cap: path-code table entries staged in shared memory (0 = read them from global memory)

// shared-memory layout: [pool: rk (u64) lo hi (int) key (u32) clist order b1 b2 b3 gr (int) dep alive (u8), pool_cap each] [2 x hist] [cell_off cell_slot (int, cell_cap + 1 each)] [path-code tables (u32, tab_cap)]
//                       [cand keyA keyB (u32, key_cap each)] [idxA idxB (u16, key_cap each)]
__host__ __device__ inline size_t qf_pool_bytes(int pool_cap) { return (size_t)pool_cap * (4 + 4 + 4 + 8 + 6 * 4 + 1 + 1); }
__host__ __device__ inline size_t qf_fixed_bytes(int pool_cap, int cell_cap, int tab_cap, int threads) {
    const size_t hist = (size_t)2 * 256 * qf_hs(threads) * 2, buckets = (size_t)(QF_BUCKETS + 1) * 4;      // the bucket table overlays the radix histograms
    return ((qf_pool_bytes(pool_cap) + 15) & ~(size_t)15) + (((hist > buckets ? hist : buckets) + 15) & ~(size_t)15) + (((size_t)(cell_cap + 1) * 8 + 15) & ~(size_t)15) + (((size_t)tab_cap * 4 + 15) & ~(size_t)15) + 256;
}
__host__ __device__ inline size_t qf_smem_bytes(const QfPlan& q) { return qf_fixed_bytes(q.pool_cap, q.cell_cap, q.tab_cap, q.threads) + (size_t)q.key_cap * 16; }

// lanes with the same 8-bit digit as this one (among the active lanes)
__device__ __forceinline__ uint32_t qf_peers(uint32_t d, bool act) {
    uint32_t p = __ballot_sync(0xffffffffu, act);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
        p &= bit ? bal : ~bal;
    }
    return p;
}

// first digit in which two path codes differ: 0 = root, 1..13 = tree digit, 14 = identical codes
__device__ __forceinline__ int qf_diff_digit(uint32_t a, uint32_t b) {
    const uint32_t x = a ^ b;
    if (x == 0u) return ORBX_MAXD + 1;
    const int hb = 31 - __clz(x);
    return hb >= ORBX_ROOT_SHIFT ? 0 : ORBX_MAXD - (hb >> 1);
}

// List-order key of a node, 32 bits, ascending = list order.  Blocks newest first: the largest-first rounds (round r: QF_LF_SPACE slots at
// (13 - r) * QF_LF_SPACE), then the depth-D blocks of the full passes for D = Q down to 0 (4^D * 16 slots each: 4-bit root index + D digits).
// Inside a full-pass block: prefix ^ xormask(D), the digits D, D-2, .. (and the root with digit 1) descending, the others ascending.
#define QF_LF_SPACE (1u << 14)
__device__ __forceinline__ uint32_t qf_tkey(uint32_t code, int D, int Q) {
    const uint32_t prefix = code >> (2 * (ORBX_MAXD - D));
    const uint32_t low = D ? (0x33333333u & ((1u << (2 * D)) - 1u)) : 0u;
    const uint32_t root = (D & 1) ? (0xFu << (2 * D)) : 0u;
    const uint32_t base = 14u * QF_LF_SPACE + 16u * (((1u << (2 * (Q + 1))) - (1u << (2 * (D + 1)))) / 3u);   // blocks Q, Q-1, .., D+1 lie in front
    return base + ((prefix ^ (low | root)) & ((16u << (2 * D)) - 1u));
}

struct QfShared {
    int wsum[33];
    int cumC[16], cumS[16];
    int counter[4];
    int rstar, ovf;
};

// inclusive block scan of one value per thread; returns the inclusive prefix, *total = sum over the block.  Two barriers.
template <int THREADS>
__device__ __forceinline__ int qf_block_scan(int v, int* wsum, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    __syncthreads();                                                          // wsum may still be read by the previous user
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int w = lane < THREADS / 32 ? wsum[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
    *total = __shfl_sync(0xffffffffu, wi, 31);
    return incl + __shfl_sync(0xffffffffu, wi - w, warp);
}

// inclusive scan in place of a[0..n) (shared memory), every thread owns a run of consecutive elements
template <int THREADS>
__device__ __forceinline__ void qf_scan_array(int* a, int n, int* wsum) {
    const int per = (n + THREADS - 1) / THREADS;
    const int i0 = min((int)threadIdx.x * per, n), i1 = min(i0 + per, n);
    int s = 0;
    for (int i = i0; i < i1; ++i) s += a[i];
    int total;
    int run = qf_block_scan<THREADS>(s, wsum, &total) - s;
    for (int i = i0; i < i1; ++i) { run += a[i]; a[i] = run; }
    __syncthreads();
}

// unordered compaction: every thread may append items; returns nothing, the count sits in *counter after the next barrier
__device__ __forceinline__ void qf_append(bool pred, int value, int* list, int* counter) {
    const uint32_t m = __ballot_sync(0xffffffffu, pred);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) list[base + __popc(m & ((1u << lane) - 1u))] = value;
}

// rank of key[e] among key[0..np) (number of smaller keys; keys are distinct), several threads per element when there are few elements
template <int THREADS>
__device__ __forceinline__ void qf_rank(const unsigned long long* key, int np, int* order) {
    int lg = 0;
    while (lg < 5 && (np << (lg + 1)) <= THREADS) ++lg;
    const int G = 1 << lg;
    for (int t = threadIdx.x; t < (((np << lg) + 31) & ~31); t += THREADS) {
        const int e = t >> lg, g = t & (G - 1);
        int r = 0;
        unsigned long long v = 0;
        if (e < np) {
            v = key[e];
            for (int f = g; f < np; f += G) r += key[f] < v;
        }
        for (int o = 1; o < G; o <<= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (e < np && g == 0) order[r] = e;
    }
}

// the same for 32-bit keys padded with 0xFFFFFFFF to a multiple of 4: four keys per load
template <int THREADS>
__device__ __forceinline__ void qf_rank32(const uint32_t* key, int np, int* order) {
    int lg = 0;
    while (lg < 5 && (np << (lg + 1)) <= THREADS) ++lg;
    const int G = 1 << lg, n4 = (np + 3) >> 2;
    const uint4* key4 = reinterpret_cast<const uint4*>(key);
    for (int t = threadIdx.x; t < (((np << lg) + 31) & ~31); t += THREADS) {
        const int e = t >> lg, g = t & (G - 1);
        int r = 0;
        if (e < np) {
            const uint32_t v = key[e];
            for (int f = g; f < n4; f += G) { const uint4 k = key4[f]; r += (k.x < v) + (k.y < v) + (k.z < v) + (k.w < v); }
        }
        for (int o = 1; o < G; o <<= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (e < np && g == 0) order[r] = e;
    }
}

template <typename IdxT, int THREADS>
__device__ __forceinline__ void qf_core(const LevelGeom& g, int n, int level, int b, int nlevels, int kp_per_frame, int pool_cap,
                                        uint8_t* pool_sm, uint16_t* hist, const int* cell_off, const int* cell_slot, const uint32_t* tab,
                                        uint32_t* cand, uint32_t* keyA, uint32_t* keyB, IdxT* idxA, IdxT* idxB,
                                        const uint32_t* __restrict__ slots, uint32_t* __restrict__ oc,
                                        uint32_t* __restrict__ kp_level, int* __restrict__ kp_count, int* __restrict__ overflow, QfShared& sh) {
    constexpr int WARPS = THREADS / 32, QF_HS = qf_hs(THREADS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;

    // ---- gather in the reference's vToDistributeKeys order (cells row-major, raster inside a cell) + path codes (two table reads and an OR) ----
    const int span = max(g.maxBX - g.minBX, g.maxBY - g.minBY);
    int nd = 1; while ((1 << nd) < span && nd < ORBX_MAXD) ++nd;
    nd = min(nd + 2, ORBX_MAXD);                                               // (k_octree.cuh: digits below that depth are constant)
    const int low = 2 * (ORBX_MAXD - nd);
    int rootbits = 0; while ((1 << rootbits) < g.nIni) ++rootbits;
    const int nbits = 2 * nd + rootbits;
    const int bshift = max(low, low + nbits - QF_BUCKET_BITS);                  // bucket = the top QF_BUCKET_BITS bits of the code's live range
    int* bcnt = reinterpret_cast<int*>(hist);                                   // [QF_BUCKETS + 1] bucket sizes, then exclusive offsets
    for (int i = tid; i <= QF_BUCKETS; i += THREADS) bcnt[i] = 0;
    if (tid == 0) sh.counter[3] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += THREADS) {
        int lo = 0, hi = g.cell_count;                                         // last cell with cell_off <= i
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (cell_off[mid] <= i) lo = mid; else hi = mid; }
        const uint32_t p = slots[cell_slot[lo] + (i - cell_off[lo])];
        cand[i] = p; oc[i] = p;
        const int x = (int)(p & 0xFFF), y = (int)((p >> 12) & 0xFFF);
        const uint32_t code = tab ? (tab[min(x, g.code_nx - 1)] | tab[g.code_nx + min(y, g.code_ny - 1)]) : octree_code(x, y, g);
        keyA[i] = code;
        idxA[i] = (IdxT)atomicAdd(&bcnt[(code >> bshift) & (QF_BUCKETS - 1)], 1);      // arrival rank inside the bucket (any order: the ranking below restores it)
    }
    __syncthreads();
    QT_STAMP(3);

    // ---- sort by (path code, original index).  Candidates are spread over the image, so a bucket (= a quadtree node ~5 levels down) holds a
    // handful of them: counting sort into the buckets, then every key ranks itself inside its bucket by direct comparison.  ~10 x fewer
    // instructions than the radix passes below, which stay as the path for clustered sets (a bucket of more than QF_BUCKET_MAX keys). ----
    bool sorted = false;
    {
        constexpr int PER = (QF_BUCKETS + THREADS - 1) / THREADS;
        int v[PER], mine = 0, big = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) { const int bi = tid * PER + k; v[k] = bi < QF_BUCKETS ? bcnt[bi] : 0; mine += v[k]; big = max(big, v[k]); }
        big = __reduce_max_sync(0xffffffffu, big);
        if (lane == 0 && big > max(QF_BUCKET_MAX, n >> 6)) sh.counter[3] = 1;   // (large levels tolerate larger buckets: the ranking stays cheaper than four radix passes over global scratch)
        int total;
        int run = qf_block_scan<THREADS>(mine, sh.wsum, &total) - mine;
#pragma unroll
        for (int k = 0; k < PER; ++k) { const int bi = tid * PER + k; if (bi < QF_BUCKETS) bcnt[bi] = run; run += v[k]; }
        if (tid == 0) bcnt[QF_BUCKETS] = n;
        __syncthreads();
        if (sh.counter[3] == 0) {
            for (int i = tid; i < n; i += THREADS) {
                const uint32_t code = keyA[i];
                const int pos = bcnt[(code >> bshift) & (QF_BUCKETS - 1)] + (int)idxA[i];
                keyB[pos] = code; idxB[pos] = (IdxT)i;
            }
            __syncthreads();
            for (int p = tid; p < n; p += THREADS) {
                const uint32_t code = keyB[p]; const uint32_t oi = (uint32_t)idxB[p];
                const int bi = (code >> bshift) & (QF_BUCKETS - 1);
                const int b0 = bcnt[bi], b1 = bcnt[bi + 1];
                int r = b0;
                for (int j = b0; j < b1; ++j) { const uint32_t cj = keyB[j]; r += (cj < code) || (cj == code && (uint32_t)idxB[j] < oi); }
                keyA[r] = code; idxA[r] = (IdxT)oi;
            }
            __syncthreads();
            sorted = true;
        } else {
            for (int i = tid; i < n; i += THREADS) idxA[i] = (IdxT)i;
        }
    }
    QT_STAMP(4);
    // ---- stable LSD radix sort on the path code, 8 bits per pass; the digits below the depth at which a node is one pixel wide are
    // constant and skipped (k_octree.cuh).  Warp w owns a contiguous segment (stability), ranks inside a step come from ballots. ----
    if (!sorted && sizeof(IdxT) == 4) {
        // levels on global scratch can hold more than 65535 keys: the same passes with 32-bit counters (one histogram, zeroed per pass)
        uint32_t* h32 = reinterpret_cast<uint32_t*>(hist);                     // [256][QF_HS]: exactly the bytes of the two 16-bit histograms
        const int seg = ((n + WARPS - 1) / WARPS + 31) & ~31;
        const int s0 = min(warp * seg, n), s1 = min(s0 + seg, n);
        for (int shift = low; shift < low + nbits; shift += 8) {
            for (int i = tid; i < 256 * QF_HS; i += THREADS) h32[i] = 0u;
            __syncthreads();
            for (int i0 = s0; i0 < s1; i0 += 32) {
                const int i = i0 + lane;
                const bool act = i < s1;
                const uint32_t d = act ? ((keyA[i] >> shift) & 255u) : 0u;
                const uint32_t peers = qf_peers(d, act);
                if (act && (peers & lt) == 0u) h32[d * QF_HS + warp] += (uint32_t)__popc(peers);
                __syncwarp();
            }
            __syncthreads();
            int tot = 0, inc = 0;
            if (tid < 256) {
                uint32_t* row = h32 + tid * QF_HS;
                for (int w = 0; w < WARPS; ++w) { const int c = (int)row[w]; row[w] = (uint32_t)tot; tot += c; }
                inc = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
                if (lane == 31) sh.wsum[warp] = inc;
            }
            __syncthreads();
            if (tid < 256) {
                int basew = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) if (w < warp) basew += sh.wsum[w];
                const uint32_t excl = (uint32_t)(basew + inc - tot);
                uint32_t* row = h32 + tid * QF_HS;
                for (int w = 0; w < WARPS; ++w) row[w] += excl;
            }
            __syncthreads();
            for (int i0 = s0; i0 < s1; i0 += 32) {
                const int i = i0 + lane;
                const bool act = i < s1;
                const uint32_t k = act ? keyA[i] : 0u;
                const uint32_t d = (k >> shift) & 255u;
                const uint32_t peers = qf_peers(d, act);
                uint32_t dst = 0;
                if (act) dst = h32[d * QF_HS + warp] + (uint32_t)__popc(peers & lt);
                __syncwarp();
                if (act && (peers & lt) == 0u) h32[d * QF_HS + warp] += (uint32_t)__popc(peers);
                if (act) { keyB[dst] = k; idxB[dst] = idxA[i]; }
                __syncwarp();
            }
            __syncthreads();
            { uint32_t* t = keyA; keyA = keyB; keyB = t; }
            { IdxT* t = idxA; idxA = idxB; idxB = t; }
        }
        sorted = true;
    }
    if (!sorted) {
        for (int i = tid; i < 256 * QF_HS; i += THREADS) reinterpret_cast<uint32_t*>(hist)[i] = 0u;  // both radix histograms (2 x 256 x QF_HS u16; they overlay the bucket table)
        __syncthreads();
        const int seg = ((n + WARPS - 1) / WARPS + 31) & ~31;
        const int s0 = min(warp * seg, n), s1 = min(s0 + seg, n);
        uint16_t* hcur = hist; uint16_t* hnext = hist + 256 * QF_HS;
        int qt_pass = 1; (void)qt_pass;
        for (int shift = low; shift < low + nbits; shift += 8) {
            for (int i0 = s0; i0 < s1; i0 += 32) {
                const int i = i0 + lane;
                const bool act = i < s1;
                const uint32_t d = act ? ((keyA[i] >> shift) & 255u) : 0u;
                const uint32_t peers = qf_peers(d, act);
                if (act && (peers & lt) == 0u) hcur[d * QF_HS + warp] += (uint16_t)__popc(peers);
                __syncwarp();
            }
            __syncthreads();
            // offsets: digit-major, warp-minor exclusive scan; thread d < 256 owns digit d
            int tot = 0, inc = 0;
            uint32_t part[WARPS / 2];
            if (tid < 256) {
                const uint4* row = reinterpret_cast<const uint4*>(hcur + tid * QF_HS);
#pragma unroll
                for (int k = 0; k < WARPS / 8; ++k) { const uint4 v = row[k]; part[4 * k] = v.x; part[4 * k + 1] = v.y; part[4 * k + 2] = v.z; part[4 * k + 3] = v.w; }
#pragma unroll
                for (int k = 0; k < WARPS / 2; ++k) {                          // two u16 counters per word -> their exclusive prefixes
                    const int c0 = (int)(part[k] & 0xFFFFu), c1 = (int)(part[k] >> 16);
                    part[k] = (uint32_t)tot | ((uint32_t)(tot + c0) << 16);
                    tot += c0 + c1;
                }
                inc = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
                if (lane == 31) sh.wsum[warp] = inc;
            } else if (THREADS > 256) {
                for (int i = tid - 256; i < 128 * QF_HS; i += THREADS - 256) reinterpret_cast<uint32_t*>(hnext)[i] = 0u;
            }
            if (THREADS <= 256) for (int i = tid; i < 128 * QF_HS; i += THREADS) reinterpret_cast<uint32_t*>(hnext)[i] = 0u;
            __syncthreads();
            if (tid < 256) {
                int basew = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) if (w < warp) basew += sh.wsum[w];
                const uint32_t excl = (uint32_t)(basew + inc - tot) * 0x00010001u;
                uint4* row = reinterpret_cast<uint4*>(hcur + tid * QF_HS);
#pragma unroll
                for (int k = 0; k < WARPS / 8; ++k) row[k] = make_uint4(part[4 * k] + excl, part[4 * k + 1] + excl, part[4 * k + 2] + excl, part[4 * k + 3] + excl);
            }
            __syncthreads();
            for (int i0 = s0; i0 < s1; i0 += 32) {
                const int i = i0 + lane;
                const bool act = i < s1;
                const uint32_t k = act ? keyA[i] : 0u;
                const uint32_t d = (k >> shift) & 255u;
                const uint32_t peers = qf_peers(d, act);
                int dst = 0;
                if (act) dst = hcur[d * QF_HS + warp] + __popc(peers & lt);
                __syncwarp();
                if (act && (peers & lt) == 0u) hcur[d * QF_HS + warp] += (uint16_t)__popc(peers);
                if (act) { keyB[dst] = k; idxB[dst] = idxA[i]; }
                __syncwarp();
            }
            __syncthreads();
            { uint32_t* t = keyA; keyA = keyB; keyB = t; }
            { IdxT* t = idxA; idxA = idxB; idxB = t; }
            { uint16_t* t = hcur; hcur = hnext; hnext = t; }
            QT_STAMP(4 + qt_pass); ++qt_pass;
        }
    }
    const uint32_t* ck = keyA;                                                 // sorted path codes
    const IdxT* sidx = idxA;                                                   // original index (position in the reference's order) per sorted position
    uint8_t* dd = reinterpret_cast<uint8_t*>(keyB);                            // n + 1 entries (keyB is free now)

    // pool of node records (pool index = creation slot; the list order lives in nkey)
    unsigned long long* rk = reinterpret_cast<unsigned long long*>(pool_sm);   // rank keys of the round (64 bits: size | list key); the final ranking reads them as u32
    int* nlo = reinterpret_cast<int*>(rk + pool_cap);
    int* nhi = nlo + pool_cap;
    uint32_t* nkey = reinterpret_cast<uint32_t*>(nhi + pool_cap);              // list key (qf_tkey): ascending = list order
    int* clist = reinterpret_cast<int*>(nkey + pool_cap);                      // candidates of the current largest-first round / living nodes at the end (pool indices)
    int* order = clist + pool_cap;                                             // element of rank r
    int* b1 = order + pool_cap; int* b2 = b1 + pool_cap; int* b3 = b2 + pool_cap;       // child boundaries of the candidate of rank r
    int* gr = b3 + pool_cap;                                                   // growth (children - 1) of rank r, then its inclusive prefix
    uint8_t* ndep = reinterpret_cast<uint8_t*>(gr + pool_cap);
    uint8_t* nalive = ndep + pool_cap;

    // ---- dd + the two cumulative histograms: C(q) = #{dd <= q} (nodes after q passes), S(q) = #{max(dd[i], dd[i+1]) <= q} (single-key nodes).
    // Warp w owns the positions [s0, s1); its own histogram row (wacc) later gives the rank of its first node without a second counting pass. ----
    const int seg = ((n + WARPS - 1) / WARPS + 31) & ~31;
    const int s0 = min(warp * seg, n), s1 = min(s0 + seg, n);
    int* wacc = order;                                                          // [WARPS][33]: C counts of the values 0..15, S counts of 0..15, one pad (pool_cap >= 1056)
    {
        uint32_t pc[4] = {0u, 0u, 0u, 0u}, ps[4] = {0u, 0u, 0u, 0u};            // 16 + 16 private 8-bit counters
        int acc = 0, pending = 0;
        auto flush = [&]() {                                                    // warp sums stay below 256 per field: at most 7 positions per lane between flushes
            uint32_t mine = 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t rc = __reduce_add_sync(0xffffffffu, pc[k]), rs = __reduce_add_sync(0xffffffffu, ps[k]);
                if (((lane & 15) >> 2) == k) mine = lane < 16 ? rc : rs;
                pc[k] = 0u; ps[k] = 0u;
            }
            acc += (int)((mine >> (8 * (lane & 3))) & 255u);
            pending = 0;
        };
        for (int i0 = s0; i0 < s1; i0 += 32) {
            const int i = i0 + lane;
            if (i < s1) {
                const uint32_t c = ck[i];
                const int d = i == 0 ? 0 : qf_diff_digit(c, ck[i - 1]);
                const int dn = i + 1 == n ? 0 : qf_diff_digit(ck[i + 1], c);
                dd[i] = (uint8_t)d;
                const int e = max(d, dn);
#pragma unroll
                for (int k = 0; k < 4; ++k) { pc[k] += (d >> 2) == k ? 1u << (8 * (d & 3)) : 0u; ps[k] += (e >> 2) == k ? 1u << (8 * (e & 3)) : 0u; }
            }
            if (++pending == 7) flush();
        }
        if (pending) flush();
        wacc[warp * 33 + lane] = acc;
        if (tid == 0) dd[n] = 0;
        __syncthreads();
        if (tid < 32) {
            int t = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) t += wacc[w * 33 + tid];
            int inc = t;                                                        // cumulative over the digit value inside each half-warp
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o, 16); if ((lane & 15) >= o) inc += u; }
            if (tid < 16) sh.cumC[tid] = inc; else sh.cumS[tid - 16] = inc;
            if (tid == 0) { sh.counter[0] = 0; sh.counter[1] = 0; sh.ovf = 0; }
        }
        __syncthreads();
    }
    QT_STAMP(16);

    // ---- number of full passes Q (:824-929), evaluated by every thread ----
    const int N = g.N;
    int Q = 0, count = sh.cumC[0];
    bool lf = false, fin = false;
    for (int q = 1; q <= ORBX_MAXD; ++q) {
        const int prev = count;
        count = sh.cumC[q];
        Q = q;
        if (count >= N || count == prev) { fin = true; break; }                // :907
        if (count + 3 * (count - sh.cumS[q]) > N) { lf = true; break; }        // :929  (nToExpand = multi-key nodes = nodes - single-key nodes)
    }
    bool ovf = (!fin && !lf) || count > pool_cap;                              // a 14th pass would be needed (identical codes) / more nodes than the pool holds
    int poolN = 0, np = 0;

    // ---- nodes after Q passes, in code order: pool index k = rank of the boundary position ----
    if (!ovf) {
        int base;
        {   // nodes in front of this warp's segment: lane l sums row l of the per-warp histograms up to Q
            int mine = 0;
            if (lane < WARPS) for (int v = 0; v <= Q; ++v) mine += wacc[lane * 33 + v];
            int inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
            base = __shfl_sync(0xffffffffu, inc - mine, warp);
        }
        for (int i0 = s0; i0 < s1; i0 += 32) {
            const int i = i0 + lane;
            const bool f = i < s1 && dd[i] <= Q;
            const uint32_t m = __ballot_sync(0xffffffffu, f);
            if (f) nlo[base + __popc(m & lt)] = i;
            base += __popc(m);
        }
        __syncthreads();
        for (int k0 = 0; k0 < count; k0 += THREADS) {
            const int k = k0 + tid;
            bool multi = false;
            if (k < count) {
                const int lo = nlo[k], hi = k + 1 < count ? nlo[k + 1] : n;
                nhi[k] = hi;
                multi = hi - lo > 1;
                const int D = multi ? Q : max((int)dd[lo], (int)dd[lo + 1]);
                ndep[k] = (uint8_t)D; nalive[k] = 1;
                nkey[k] = qf_tkey(ck[lo], D, Q);
            }
            if (lf) qf_append(multi, k, clist, &sh.counter[0]);
        }
        __syncthreads();
        poolN = count;
        np = lf ? sh.counter[0] : 0;
    }
    QT_STAMP(17);

    // ---- largest-first rounds (:929-1011) ----
    int round = 0;
    while (lf && !ovf && np > 0) {                                              // np == 0: nothing to divide, the size does not change (:1009)
        const int prevSize = count;
        // rank by (size desc, list order = newest first)
        for (int e = tid; e < np; e += THREADS) {
            const int k = clist[e];
            rk[e] = ((unsigned long long)(0xFFFFFFu - (unsigned)(nhi[k] - nlo[k])) << 32) | nkey[k];
        }
        if (tid == 0) { sh.rstar = np - 1; sh.counter[(round + 1) & 1] = 0; }
        __syncthreads();
        qf_rank<THREADS>(rk, np, order);
        __syncthreads();
        if (round == 0) QT_STAMP(24);
        // children of the candidate of rank r: boundaries c1..c3 (three interleaved lower bounds on the sorted codes), growth = non-empty children - 1
        for (int r = tid; r < np; r += THREADS) {
            const int k = clist[order[r]];
            const int lo = nlo[k], hi = nhi[k], dep = ndep[k];
            int m = 1, c1 = lo, c2 = lo, c3 = lo;
            if (dep >= ORBX_MAXD) sh.ovf = 1;
            else {
                const int shift = 2 * (ORBX_MAXD - dep - 1);
                const uint32_t prefix = ck[lo] >> (shift + 2);
                const uint32_t t1 = ((prefix << 2) | 1u) << shift, t2 = ((prefix << 2) | 2u) << shift, t3 = ((prefix << 2) | 3u) << shift;
                int l1 = lo, h1 = hi, l2 = lo, h2 = hi, l3 = lo, h3 = hi;
                while (l1 < h1 || l2 < h2 || l3 < h3) {
                    if (l1 < h1) { const int mid = (l1 + h1) >> 1; if (ck[mid] < t1) l1 = mid + 1; else h1 = mid; }
                    if (l2 < h2) { const int mid = (l2 + h2) >> 1; if (ck[mid] < t2) l2 = mid + 1; else h2 = mid; }
                    if (l3 < h3) { const int mid = (l3 + h3) >> 1; if (ck[mid] < t3) l3 = mid + 1; else h3 = mid; }
                }
                c1 = l1; c2 = l2; c3 = l3;
                m = (c1 > lo) + (c2 > c1) + (c3 > c2) + (hi > c3);
            }
            order[r] = k;                                                       // from here on: pool index of the candidate of rank r
            gr[r] = m - 1; b1[r] = c1; b2[r] = c2; b3[r] = c3;
        }
        __syncthreads();
        if (round == 0) QT_STAMP(25);
        if (sh.ovf) { ovf = true; break; }
        qf_scan_array<THREADS>(gr, np, sh.wsum);
        for (int r = tid; r < np; r += THREADS)
            if (count + gr[r] >= N && (r == 0 || count + gr[r - 1] < N)) sh.rstar = r;   // first candidate after whose divide the list holds >= N nodes (:1003)
        __syncthreads();
        if (round == 0) QT_STAMP(26);
        const int rstar = sh.rstar, P = rstar + 1, growth = gr[rstar];
        if (poolN + growth > pool_cap) { ovf = true; break; }                   // cannot happen: the list never holds more than N + 3 nodes
        int* cnt_next = &sh.counter[(round + 1) & 1];
        for (int r0 = 0; r0 < P; r0 += THREADS) {                               // the visited candidates leave the list, their children form the next block
            const int r = r0 + tid;
            const bool actv = r < P;
            int k = 0, lo = 0, hi = 0, dep = 0, c1 = 0, c2 = 0, c3 = 0, extra = 0;
            if (actv) {
                k = order[r];
                lo = nlo[k]; hi = nhi[k]; dep = ndep[k]; c1 = b1[r]; c2 = b2[r]; c3 = b3[r];
                extra = poolN + (r ? gr[r - 1] : 0);                            // the first child takes over the erased node's record (:991), the others get new ones
            }
            const int bnd[5] = {lo, c1, c2, c3, hi};
            bool first = true;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int sz = bnd[q + 1] - bnd[q];
                const bool have = actv && sz > 0;
                int slot = 0;
                if (have) {
                    slot = first ? k : extra++;
                    first = false;
                    nlo[slot] = bnd[q]; nhi[slot] = bnd[q + 1]; ndep[slot] = (uint8_t)(dep + 1); nalive[slot] = 1;
                    nkey[slot] = (uint32_t)(13 - round) * QF_LF_SPACE + (uint32_t)(((np - 1 - r) << 2) | (3 - q));   // reverse creation order
                }
                qf_append(have && sz > 1, slot, clist, cnt_next);               // clist is dead since the barrier above: the next round's candidates go there
            }
        }
        __syncthreads();
        count += growth; poolN += growth;
        np = *cnt_next;
        ++round;
        QT_STAMP(17 + round);
        if (count >= N || count == prevSize) break;                             // :1009
    }

    // ---- final list: living nodes by (block age, block key); one keypoint per node (:1018-1048) ----
    uint32_t* out = kp_level + (long long)b * kp_per_frame + g.kp_off;
    int nout = 0;
    __syncthreads();
    if (!ovf) {
        if (tid == 0) sh.counter[2] = 0;
        __syncthreads();
        for (int k0 = 0; k0 < poolN; k0 += THREADS) { const int k = k0 + tid; qf_append(k < poolN && nalive[k], k, clist, &sh.counter[2]); }
        __syncthreads();
        const int K = sh.counter[2];                                            // == count
        uint32_t* rk32 = reinterpret_cast<uint32_t*>(rk);
        for (int e = tid; e < ((K + 3) & ~3); e += THREADS) rk32[e] = e < K ? nkey[clist[e]] : 0xFFFFFFFFu;
        __syncthreads();
        QT_STAMP(30);
        qf_rank32<THREADS>(rk32, K, order);
        __syncthreads();
        QT_STAMP(31);
        nout = min(K, g.kp_cap);
        if (K > g.kp_cap) ovf = true;
        for (int t0 = 0; t0 < nout * 8; t0 += THREADS) {                        // eight lanes per node
            const int r = (t0 + tid) >> 3, sub = tid & 7;
            uint32_t best = 0u;
            if (r < nout) {
                const int k = clist[order[r]];
                const int lo = nlo[k], hi = nhi[k];
                for (int i = lo + sub; i < hi; i += 8) {
                    const uint32_t oi = (uint32_t)sidx[i];
                    best = max(best, (cand[oi] & 0xFF000000u) | (0xFFFFFFu - oi));   // max response, first in the original order
                }
            }
            best = max(best, __shfl_xor_sync(0xffffffffu, best, 1));
            best = max(best, __shfl_xor_sync(0xffffffffu, best, 2));
            best = max(best, __shfl_xor_sync(0xffffffffu, best, 4));
            if (r < nout && sub == 0) out[r] = cand[0xFFFFFFu - (best & 0xFFFFFFu)];
        }
    }
    QT_STAMP(39);
    if (tid == 0) {
        kp_count[b * nlevels + level] = nout;
        if (ovf) atomicOr(overflow, ORBX_OVF_TREE);
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_octree_fused(const __grid_constant__ QfLevels P, const CellDesc* __restrict__ cells, int ncells, int slots_per_frame, int cand_per_frame,
               int kp_per_frame, int nlevels, QfPlan plan,
               uint32_t* cand_slots,                    // [B][slots_per_frame] FAST candidates per cell slot (an oversized level reuses its slice as scratch once gathered)
               const uint16_t* __restrict__ cell_counts,
               uint32_t* __restrict__ ocand,            // [B][cand_per_frame] candidates in reference order (debug taps; candidate array of oversized levels)
               unsigned long long* __restrict__ skey,   // [B][cand_per_frame] global scratch of oversized levels (code ping / pong)
               uint32_t* __restrict__ spk,              // [B][cand_per_frame] global scratch of oversized levels (index ping)
               int* __restrict__ ncand, uint32_t* __restrict__ kp_level, int* __restrict__ kp_count, int* __restrict__ overflow) {
    extern __shared__ __align__(16) uint8_t qf_sm[];
    __shared__ QfShared sh;
    const int level = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    QT_STAMP(0);
    const LevelGeom& g = P.lv[level];
    uint8_t* pool_sm = qf_sm;
    uint16_t* hist = reinterpret_cast<uint16_t*>(qf_sm + ((qf_pool_bytes(plan.pool_cap) + 15) & ~(size_t)15));
    int* cell_off = reinterpret_cast<int*>(qf_sm + qf_fixed_bytes(plan.pool_cap, 0, 0, THREADS) - 256 - 16);
    int* cell_slot = cell_off + plan.cell_cap + 1;
    uint32_t* tab_sm = reinterpret_cast<uint32_t*>(qf_sm + qf_fixed_bytes(plan.pool_cap, plan.cell_cap, 0, THREADS) - 256);
    uint8_t* keys_sm = qf_sm + qf_fixed_bytes(plan.pool_cap, plan.cell_cap, plan.tab_cap, THREADS);
    // path-code tables of the level: staged beside the counts when they fit (x entries, then y entries: contiguous in the level's global table)
    const uint32_t* tab = nullptr;
    if (g.code_x) {
        tab = g.code_x;
        if (g.code_nx + g.code_ny <= plan.tab_cap) { for (int i = tid; i < g.code_nx + g.code_ny; i += THREADS) tab_sm[i] = g.code_x[i]; tab = tab_sm; }
    }

    // ---- exclusive scan of the per-cell counts (cells in row-major order); slot offsets of the cells staged beside them ----
    const uint16_t* counts = cell_counts + (long long)b * ncells + g.cell_begin;
    const int per = (g.cell_count + THREADS - 1) / THREADS;
    const int c0 = min(tid * per, g.cell_count), c1 = min(c0 + per, g.cell_count);
    int mine = 0;
    for (int c = c0; c < c1; ++c) { const int v = counts[c]; cell_off[c] = v; cell_slot[c] = cells[g.cell_begin + c].slot; mine += v; }
    int n;
    int run = qf_block_scan<THREADS>(mine, sh.wsum, &n) - mine;
    for (int c = c0; c < c1; ++c) { const int v = cell_off[c]; cell_off[c] = run; run += v; }
    if (tid == 0) { cell_off[g.cell_count] = n; ncand[b * nlevels + level] = n; }
    __syncthreads();
    QT_STAMP(1);
    if (n == 0) { if (tid == 0) kp_count[b * nlevels + level] = 0; return; }
    const uint32_t* slots = cand_slots + (long long)b * slots_per_frame;
    uint32_t* oc = ocand + (long long)b * cand_per_frame + g.cand_off;
    if (n <= plan.key_cap) {
        uint32_t* cand = reinterpret_cast<uint32_t*>(keys_sm);
        uint32_t* keyA = cand + plan.key_cap; uint32_t* keyB = keyA + plan.key_cap;
        uint16_t* idxA = reinterpret_cast<uint16_t*>(keyB + plan.key_cap); uint16_t* idxB = idxA + plan.key_cap;
        qf_core<uint16_t, THREADS>(g, n, level, b, nlevels, kp_per_frame, plan.pool_cap, pool_sm, hist, cell_off, cell_slot, tab, cand, keyA, keyB, idxA, idxB,
                                   slots, oc, kp_level, kp_count, overflow, sh);
    } else {
        // oversized level: the same code on global scratch (the level's slices of skey = two u32 arrays, spk, and the cell slots themselves, which are dead once
        // gathered; ocand doubles as the candidate array).
        uint32_t* keyA = reinterpret_cast<uint32_t*>(skey + (long long)b * cand_per_frame + g.cand_off); uint32_t* keyB = keyA + g.cand_cap;
        uint32_t* idxA = spk + (long long)b * cand_per_frame + g.cand_off; uint32_t* idxB = cand_slots + (long long)b * slots_per_frame + g.slot_off;
        if (n >= (1 << 24) || n > g.cand_cap) { if (tid == 0) { kp_count[b * nlevels + level] = 0; atomicOr(overflow, ORBX_OVF_SORT); } return; }
        qf_core<uint32_t, THREADS>(g, n, level, b, nlevels, kp_per_frame, plan.pool_cap, pool_sm, hist, cell_off, cell_slot, tab, oc, keyA, keyB, idxA, idxB,
                                   slots, oc, kp_level, kp_count, overflow, sh);
    }
}
