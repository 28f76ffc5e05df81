// k_fast.cuh -- per-cell FAST-9/16 detection with the iniThFAST / minThFAST retry and 3x3 non-max suppression
// (/root/reference/src/ORBextractor.cc:1089-1157 + cv::FAST semantics), in score-map form:
//   S(p) = max over the 16 arcs of 9 contiguous ring pixels of min(v - p_k)  or  min(p_k - v);
//   corner at threshold t <=> S > t;  response = S - 1;
//   a cell's keypoints at threshold t = strict 3x3 local maxima of S inside the cell's zone with S > t
//   (neighbours outside the zone count as 0), in raster order; if the cell yields none at iniTh, the
//   same set at minTh is used.
//
// Shape of the kernel (what the profiles asked for, profiles/r02_fast_notes.md): the stage is bound by instruction issue AND by how many
// warps fit next to each other (shared memory per warp), so both the instruction stream and the per-warp footprint are kept small:
//   * one warp per cell, each warp walks cells cell0, cell0 + W, ...; the cell's ROI (widened to the left to a 16-byte boundary: the
//     innermost TMA coordinate must be 16-byte aligned, measured with tools/probes/tma_canon.cu) is fetched by ONE bulk tensor copy
//     (TMA, cp.async.bulk.tensor.3d over (x, y, frame)) issued by lane 0 and awaited on an mbarrier: no copy loop, no address
//     arithmetic.  The patch is single-buffered: the next cell's copy is issued as soon as the last stage that reads the patch is
//     over and lands behind the non-max suppression / emission of the current cell;
//   * stage A, byte-SIMD pre-test, 4 pixels (one word) per lane per step: any 9-arc contains ring point 0 or 8 and ring point 4 or
//     12, so a corner needs (|N-v| > t or |S-v| > t) and (|E-v| > t or |W-v| > t) -- exact thresholds by a carry trick.  A lane owns
//     one word column of a strip of rows and keeps the 7 rows of that column in registers (3 shared-memory loads per word).
//     Surviving words (flags of ~17 % of the pixels) are compacted into a word list by one ballot per step;
//   * the word list is expanded 32 words at a time into a small ring of pixel codes (one prefix sum per 32 words), which feeds
//   * stage S with full warps, one survivor per lane: the FAST score of BOTH polarities at once on packed 16-bit pairs
//     (lo = p_k - v + 256, hi = v - p_k + 256, one IMAD per ring pixel) through a 3-input min/max network of 40 VIMNMX3.U16x2;
//     corner <=> max(lo, hi) - 256 > t.  No separate segment test: the score is the test;
//   * stage D, strict 3x3 maxima among the corners into a bitmap of the zone, from which the candidates are emitted in raster
//     order (the order of the reference's vToDistributeKeys) by one prefix sum over the rows.  Corners are kept in a 256-entry list;
//     a cell with more corners than that (a quarter of its pixels) scans its score map instead.
// Candidates go to the cell's fixed slot range (capacity = max possible local maxima): no global atomics, deterministic layout.
#pragma once
#include "orbx_common.cuh"
#include "tma.cuh"

#define FAST_WARPS 4                // (a 64-register cap for 32 warps per SM was measured: 2.06 vs 1.93 ms, the spills cost more than the warps give)
#define FAST_RING 256          // pixel-code ring (entries): >= 31 left over + 128 from one expansion step
#define FAST_CLIST 256         // corner list (entries)

// per-warp shared memory: [patch | S | wq | ring | clist | bm | mbarrier]   (byte sizes; per_warp a multiple of 128)
// CTA form (one cell per CTA, see below): [patch | S | bm | mbarrier + 2 ints] shared (cta_shared bytes), then [wq | ring | clist] per warp (cta_per_warp)
struct FastLayout { int patch_cap, s_cap, wq_cap, bm_cap, per_warp, cta_shared, cta_per_warp; };

// exact byte-wise "d > T" (T <= 126) up to the final & 0x80808080: bit 7 of each byte of the result
__device__ __forceinline__ uint32_t fast_gt(uint32_t d, uint32_t kc) { return ((d & 0x7F7F7F7Fu) + kc) | d; }
__device__ __forceinline__ int max(int a, int b, int c) { return max(max(a, b), c); }

// CTA = false: throughput form, one WARP per cell (each warp walks several cells).  CTA = true: latency form for a handful of frames, one CTA per
// cell: the four warps share the patch, the score map and the bitmap and split the zone's rows between them for the two expensive stages
// (pre-test and score), so a cell takes a quarter of the dependent instruction chain (a lone warp per scheduler runs at ~25 cycles per
// instruction: 29 us for the FAST stage of one 640 x 480 frame in the warp form).
template <bool CTA>
__global__ void __launch_bounds__(FAST_WARPS * 32)
k_fast_cells(const __grid_constant__ CUtensorMap map_l0, const CUtensorMap* __restrict__ maps, int b0,
             const LevelGeom* __restrict__ levels, const CellDesc* __restrict__ cells, int ncells,
             int slots_per_frame, FastLayout lay, int iniTh, int minTh,
             uint32_t* __restrict__ cand_slots,      // [B][slots_per_frame]  packed x:12|y:12|resp:8 (x,y relative to minBorder)
             uint16_t* __restrict__ cell_counts) {   // [B][ncells]
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int W = CTA ? (int)gridDim.x : (int)gridDim.x * FAST_WARPS;
    int cell = CTA ? (int)blockIdx.x : (int)blockIdx.x * FAST_WARPS + warp;
    if (cell >= ncells) return;
    uint8_t* smw = CTA ? smem_raw : smem_raw + (size_t)warp * lay.per_warp;
    uint8_t* priv = CTA ? smem_raw + lay.cta_shared + (size_t)warp * lay.cta_per_warp : smw + lay.patch_cap + lay.s_cap;   // [wq | ring | clist]
    const uint8_t* patch = smw;                                             // patch column pc0 = zone x 0, patch row 3 = zone y 0
    uint8_t* S = smw + lay.patch_cap;                                       // score map of the zone with a 1-px zero ring, row stride sst
    uint32_t* wq = reinterpret_cast<uint32_t*>(priv);                       // surviving words: flags (bits 7,15,23,31) | word column << 8 | zone row
    uint16_t* ring = reinterpret_cast<uint16_t*>(priv + lay.wq_cap);        // pixel codes y << 6 | x (zone coordinates)
    uint16_t* clist = ring + FAST_RING;                                     // corners
    uint32_t* bm = CTA ? reinterpret_cast<uint32_t*>(S + lay.s_cap) : reinterpret_cast<uint32_t*>(clist + FAST_CLIST);   // 64 bits per zone row: local maxima
    uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bm) + lay.bm_cap);
    int* sh_n = reinterpret_cast<int*>(bar + 1);                            // CTA form: the cell's candidate count, shared by the warps
    const uint32_t lt = (1u << lane) - 1u;
    const int tid0 = CTA ? (int)threadIdx.x : lane, tstep = CTA ? FAST_WARPS * 32 : 32;
    auto sync = [&]() { if (CTA) __syncthreads(); else __syncwarp(); };

    for (int i = tid0; i < (lay.s_cap >> 2); i += tstep) reinterpret_cast<uint32_t*>(S)[i] = 0u;
    for (int i = tid0; i < (lay.bm_cap >> 2); i += tstep) bm[i] = 0u;
    if (tid0 == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    sync();

    auto issue = [&](const CellDesc& cd) {
        if (tid0 == 0) {
            const LevelGeom& g = levels[cd.level];
            const CUtensorMap* m = cd.level == 0 ? &map_l0 : maps + cd.level;
            mbar_expect_tx(bar, (uint32_t)(g.fast_bw * g.fast_bh));
            tma_load_3d(smw, m, (cd.x0 - 1) & ~15, cd.y0, b0 + b, bar);
        }
    };

    CellDesc c = cells[cell];
    issue(c);
    uint32_t phase = 0;
    for (; cell < ncells; cell += W) {
    const int next = cell + W;
    CellDesc cnext = c;
    if (next < ncells) cnext = cells[next];
    bool issued = next >= ncells;                                           // nothing to prefetch after the last cell

    const int pc0 = c.x0 + 3 - ((c.x0 - 1) & ~15);                          // 4 .. 19
    const int pitch = levels[c.level].fast_bw;
    const int wpr = pitch >> 2;
    const uint32_t* p32 = reinterpret_cast<const uint32_t*>(patch);
    const int zw = c.cw - 6, zh = c.ch - 6;
    const int sst = zw + 2;
    // stage-A geometry (host-computed, build_plan): lane = (strip s, word column j of the patch words wi0 .. wi0 + nwz - 1 that hold zone pixels)
    const int wi0 = pc0 >> 2, xoff = pc0 & 3;
    const int nwz = c.geo & 0xFF, strips = c.geo >> 16;
    const int r_lo = CTA ? warp * zh / FAST_WARPS : 0, r_hi = CTA ? (warp + 1) * zh / FAST_WARPS : zh;   // zone rows of this warp
    const int rps = CTA ? (r_hi - r_lo + strips - 1) / strips : (c.geo >> 8) & 0xFF;
    const int s_ = (lane * c.rcp) >> 16, j = lane - s_ * nwz;               // lane / nwz, lane % nwz
    const bool owner = s_ < strips && r_lo + s_ * rps < r_hi;
    const int ys = owner ? r_lo + s_ * rps : 0;
    const int ye = owner ? min(ys + rps, r_hi) : 0;
    // bytes of word j inside the zone: zone x = 4 j + k - xoff in [0, zw)
    const int kfirst = max(0, xoff - 4 * j), klast = min(3, zw - 1 + xoff - 4 * j);
    const uint32_t colmask = (0x80808080u << (8 * kfirst)) & (0x80808080u >> (8 * (3 - klast)));
    uint32_t* out = cand_slots + (long long)b * slots_per_frame + c.slot;

    mbar_wait(bar, phase);
    phase ^= 1u;

    int n = 0;
    for (int pass = 0; pass < 2; ++pass) {
    const int T = pass ? minTh : iniTh;
    const uint32_t kc = (uint32_t)(127 - min(T, 126)) * 0x01010101u;
    // ---- stage A: pre-test, surviving words compacted by one ballot per step ----
    int nW = 0;
    {
        const uint32_t* rp = p32 + ys * wpr + wi0 + j;                      // this lane's word column, patch row of the current N pixel
        const int w3 = 3 * wpr, w6 = 6 * wpr;
        uint32_t rg[7];
#pragma unroll
        for (int k = 0; k < 6; ++k) rg[k] = rp[k * wpr];
        rg[6] = 0;
        uint32_t ent = ((uint32_t)j << 8) | (uint32_t)ys;                   // entry without flags; the row advances with the loop
        const uint32_t ent_end = ((uint32_t)j << 8) | (uint32_t)ye;
        for (int r = 0; r < rps; r += 7) {
#pragma unroll
            for (int u = 0; u < 7; ++u) {
                if (r + u < rps) {                                          // warp-uniform
                    uint32_t f = 0;
                    if (ent < ent_end) {
                        rg[(u + 6) % 7] = rp[w6];
                        const uint32_t Cl = rp[w3 - 1], Cr = rp[w3 + 1];
                        const uint32_t C = rg[(u + 3) % 7], Nn = rg[u % 7], Ss = rg[(u + 6) % 7];
                        const uint32_t E = __byte_perm(C, Cr, 0x6543), Wn = __byte_perm(Cl, C, 0x4321);
                        f = (fast_gt(__vabsdiffu4(C, Nn), kc) | fast_gt(__vabsdiffu4(C, Ss), kc)) &
                            (fast_gt(__vabsdiffu4(C, E), kc) | fast_gt(__vabsdiffu4(C, Wn), kc)) & colmask;
                    }
                    const uint32_t any = __ballot_sync(0xffffffffu, f != 0);
                    if (f) wq[nW + __popc(any & lt)] = f | ent;
                    nW += __popc(any);
                    rp += wpr; ++ent;
                }
            }
        }
    }
    __syncwarp();
    // ---- expansion of the word list into the pixel ring + stage S (score of both polarities) on full warps ----
    int cn = 0;                                                             // corners (clist holds the first FAST_CLIST of them)
    {
        int head = 0, tail = 0, wpos = 0;
        const uint8_t* pbase = patch + 3 * pitch + pc0;
        while (true) {
            while (tail - head < 32 && wpos < nW) {                         // warp-uniform
                const int k = wpos + lane;
                const uint32_t e = k < nW ? wq[k] : 0u;
                const int cnt = __popc(e & 0x80808080u);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                int off = tail + incl - cnt;
                const int code = (int)((e & 63u) << 6) + 4 * (int)((e >> 8) & 31u) - xoff;   // + not |: byte 0 of the first word may lie left of the zone
                if (e & 0x00000080u) ring[(off++) & (FAST_RING - 1)] = (uint16_t)code;
                if (e & 0x00008000u) ring[(off++) & (FAST_RING - 1)] = (uint16_t)(code + 1);
                if (e & 0x00800000u) ring[(off++) & (FAST_RING - 1)] = (uint16_t)(code + 2);
                if (e & 0x80000000u) ring[(off++) & (FAST_RING - 1)] = (uint16_t)(code + 3);
                tail += __shfl_sync(0xffffffffu, incl, 31);
                wpos += 32;
                __syncwarp();
            }
            const int navail = min(32, tail - head);
            if (navail == 0) break;
            int code = 0, sc = 0;
            if (lane < navail) {
                code = ring[(head + lane) & (FAST_RING - 1)];
                const int y = code >> 6, x = code & 63;
                const uint8_t* p = pbase + y * pitch + x;
                const uint8_t* pu = p - 3 * pitch;
                const uint8_t* pd = p + 3 * pitch;
                const uint32_t v = p[0];
                uint32_t P[16];
                // ring order (dx,dy): (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
                P[0] = pd[0];            P[1] = pd[1];            P[2] = p[2 * pitch + 2];   P[3] = p[pitch + 3];
                P[4] = p[3];             P[5] = p[3 - pitch];     P[6] = p[2 - 2 * pitch];   P[7] = pu[1];
                P[8] = pu[0];            P[9] = pu[-1];           P[10] = p[-2 * pitch - 2]; P[11] = p[-pitch - 3];
                P[12] = p[-3];           P[13] = p[pitch - 3];    P[14] = p[2 * pitch - 2];  P[15] = pd[-1];
                const uint32_t cst = (256u - v) + ((256u + v) << 16);
#pragma unroll
                for (int i = 0; i < 16; ++i) P[i] = P[i] * 0xFFFF0001u + cst;    // lo = p_k - v + 256, hi = v - p_k + 256
                uint32_t m3[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) m3[i] = __vimin3_u16x2(P[i], P[(i + 1) & 15], P[(i + 2) & 15]);
                uint32_t best = 0;
#pragma unroll
                for (int i = 0; i < 16; i += 2)
                    best = __vimax3_u16x2(best, __vimin3_u16x2(m3[i], m3[(i + 3) & 15], m3[(i + 6) & 15]),
                                          __vimin3_u16x2(m3[(i + 1) & 15], m3[(i + 4) & 15], m3[(i + 7) & 15]));
                sc = (int)max(best & 0xFFFFu, best >> 16) - 256;
            }
            head += navail;
            const bool corner = sc > T;                                      // sc = 0 on idle lanes, T >= 0
            const uint32_t m = __ballot_sync(0xffffffffu, corner);
            if (corner) {
                S[((code >> 6) + 1) * sst + (code & 63) + 1] = (uint8_t)sc;
                const int ci = cn + __popc(m & lt);
                if (ci < FAST_CLIST) clist[ci] = (uint16_t)code;
            }
            cn += __popc(m);
            __syncwarp();
        }
    }
    if (CTA) __syncthreads();                                                // every warp's scores are in S before anybody looks at neighbours
    // the patch is dead unless this cell may need its second pass: let the next cell's copy start now
    const bool may_retry = pass == 0 && minTh != iniTh;
    if (!may_retry && !issued) { issue(cnext); issued = true; }
    // ---- stage D: strict 3x3 local maxima among the corners -> bitmap ----
    int nmax = 0;
    if (cn <= FAST_CLIST) {
        for (int k0 = 0; k0 < cn; k0 += 32) {
            const int k = k0 + lane;
            bool mx = false;
            if (k < cn) {
                const int code = clist[k], y = code >> 6, x = code & 63;
                const uint8_t* q = S + (y + 1) * sst + x + 1;
                const uint8_t* qu = q - sst;
                const uint8_t* qd = q + sst;
                const int s = q[0];                                          // branch-free: all 8 neighbours, one 3-input max tree
                const int nb = max(max(max((int)qu[-1], (int)qu[0], (int)qu[1]), (int)q[-1], (int)q[1]), max((int)qd[-1], (int)qd[0], (int)qd[1]));
                mx = s > nb;
                if (mx) atomicOr(&bm[2 * y + (x >> 5)], 1u << (x & 31));
            }
            nmax += __popc(__ballot_sync(0xffffffffu, mx));
        }
    } else {                                                                 // very dense cell: walk the score map itself
        for (int i = lane; i < zw * (r_hi - r_lo); i += 32) {
            const int y = r_lo + i / zw, x = i % zw;
            const uint8_t* q = S + (y + 1) * sst + x + 1;
            const int s = q[0];
            if (s && s > q[-1] && s > q[1] && s > q[-sst - 1] && s > q[-sst] && s > q[-sst + 1] && s > q[sst - 1] && s > q[sst] && s > q[sst + 1]) {
                atomicOr(&bm[2 * y + (x >> 5)], 1u << (x & 31));
                nmax = 1;
            }
        }
        nmax = __any_sync(0xffffffffu, nmax != 0);
    }
    sync();
    if (!CTA && may_retry && nmax > 0 && !issued) { issue(cnext); issued = true; }  // no second pass: the patch is dead
    // ---- ordered emission: rows in raster order, one row per lane (CTA form: warp 0 emits for the cell, then shares the count) ----
    if (CTA ? warp == 0 : nmax > 0) {
        for (int y0 = 0; y0 < zh; y0 += 32) {
            const int y = y0 + lane;
            uint32_t lo = 0, hi = 0;
            if (y < zh) { lo = bm[2 * y]; hi = bm[2 * y + 1]; }
            const int cnt = __popc(lo) + __popc(hi);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            int pos = n + incl - cnt;
            n += __shfl_sync(0xffffffffu, incl, 31);
            if (cnt) {
                bm[2 * y] = 0u; bm[2 * y + 1] = 0u;
                const uint8_t* srow = S + (y + 1) * sst + 1;
                const uint32_t ybits = (uint32_t)(y + 3 + c.sy) << 12;
                while (lo) { const int x = __ffs(lo) - 1; lo &= lo - 1; out[pos++] = (uint32_t)(x + 3 + c.sx) | ybits | ((uint32_t)(srow[x] - 1) << 24); }
                while (hi) { const int x = 32 + __ffs(hi) - 1; hi &= hi - 1; out[pos++] = (uint32_t)(x + 3 + c.sx) | ybits | ((uint32_t)(srow[x] - 1) << 24); }
            }
        }
    }
    if (CTA) {                                                               // all warps take the same decision
        if (threadIdx.x == 0) sh_n[0] = n;
        __syncthreads();
        n = sh_n[0];
    }
    if (n > 0 || !may_retry) {
        sync();
        if (cn <= FAST_CLIST) { for (int k = lane; k < cn; k += 32) { const int code = clist[k]; S[((code >> 6) + 1) * sst + (code & 63) + 1] = 0; } }
        else if (!CTA) { for (int i = lane; i < (lay.s_cap >> 2); i += 32) reinterpret_cast<uint32_t*>(S)[i] = 0u; }
        else { for (int i = lane; i < sst * (r_hi - r_lo); i += 32) S[(r_lo + 1) * sst + i] = 0; }      // this warp's rows of the score map
        break;                                                               // S is all zero again
    }
    sync();                                                                  // lists are rebuilt by the second pass (the scores already in S are a subset of its scores)
    }
    if (!issued) issue(cnext);                                              // (a second pass that found nothing issues here)
    if (tid0 == 0) cell_counts[(long long)b * ncells + cell] = (uint16_t)n;
    c = cnext;
    sync();                                                                  // S / lists are reused by the next cell
    }
}
