// orbx_extractor.cu -- host side of the B200-native ORBextractor: geometry plan, device buffers,
// kernel orchestration and the C ABI declared in include/orbx_b200.h.
//
// Mirrors ORB_SLAM2::ORBextractor (/root/reference/include/ORBextractor.h:93-168,
// /root/reference/src/ORBextractor.cc).  There is no CPU path here: every stage is a CUDA kernel
// (k_pyramid_fast.cuh, k_octree.cuh, k_describe.cuh, k_cull.cuh); the host only computes the small
// per-geometry tables exactly as the reference's constructor / ComputePyramid / cell loop do.
#include <memory>
#include <mutex>
#include <thread>
#include <atomic>
#include <utility>
#include "../../include/orbx_b200.h"
#include "../../include/orbx_b200_testtaps.h"
#include "orbx_common.cuh"
#include "orbx_internal.h"
#include "k_pyramid_fast.cuh"
#include "k_fast.cuh"
#include "k_octree.cuh"
#include "k_octree_fused.cuh"
#include "k_describe.cuh"
#include "k_cull.cuh"
#include "brief_pattern.inc"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static_assert(sizeof(orbx_keypoint) == 28, "orbx_keypoint must match cv::KeyPoint");
static_assert(sizeof(KpOut) == 28, "KpOut must match cv::KeyPoint");

static thread_local std::string g_last_error;
extern "C" const char* orbx_last_error(void) { return g_last_error.c_str(); }
extern "C" int orbx_version(void) { return 100; }
void orbx_set_error(const std::string& s) { g_last_error = s; }

#define CU_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { \
    orbx_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); return ORBX_E_CUDA; } } while (0)
#define FAIL(code, msg) do { orbx_set_error(msg); return (code); } while (0)

static inline int cv_round_f(float v) { return (int)lrintf(v); }      // cvRound: round half to even
static inline int cv_floor_f(float v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil_f(float v) { int i = (int)v; return i + (i < v); }
static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// Guard zones around every device buffer of an extractor handle (ORBX_CANARY=1 in the environment): compute-sanitizer is closed on this
// pool, so the parity suite has one test that runs the whole extractor with 256 guard bytes of a known pattern before and after each
// allocation and checks afterwards that no kernel wrote into them (orbx_debug_canary_check, test taps).  Off by default: no cost.
#define ORBX_GUARD 256
struct CanaryRegistry {
    std::mutex mu; std::vector<std::pair<uint8_t*, size_t>> blocks;                 // raw pointer, payload bytes
    static CanaryRegistry& get() { static CanaryRegistry r; return r; }
    static bool on() { static const bool v = [] { const char* e = std::getenv("ORBX_CANARY"); return e && std::atoi(e) != 0; }(); return v; }
};
static cudaError_t guarded_malloc(void** p, size_t bytes) {
    if (!CanaryRegistry::on()) return cudaMalloc(p, bytes);
    uint8_t* raw = nullptr;
    cudaError_t e = cudaMalloc((void**)&raw, bytes + 2 * ORBX_GUARD);
    if (e != cudaSuccess) return e;
    cudaMemset(raw, 0xA5, ORBX_GUARD); cudaMemset(raw + ORBX_GUARD + bytes, 0xA5, ORBX_GUARD);
    { std::lock_guard<std::mutex> lk(CanaryRegistry::get().mu); CanaryRegistry::get().blocks.emplace_back(raw, bytes); }
    *p = raw + ORBX_GUARD;
    return cudaSuccess;
}
static void guarded_free(void* p) {
    if (!p) return;
    if (!CanaryRegistry::on()) { cudaFree(p); return; }
    uint8_t* raw = (uint8_t*)p - ORBX_GUARD;
    { std::lock_guard<std::mutex> lk(CanaryRegistry::get().mu); auto& b = CanaryRegistry::get().blocks;
      for (size_t i = 0; i < b.size(); ++i) if (b[i].first == raw) { b.erase(b.begin() + i); break; } }
    cudaFree(raw);
}

template <typename T> struct DevBuf {
    T* p = nullptr; size_t n = 0;
    int ensure(size_t count) {
        if (count <= n) return ORBX_OK;
        if (p) guarded_free(p);
        p = nullptr; n = 0;
        CU_TRY(guarded_malloc((void**)&p, count * sizeof(T)));
        n = count; return ORBX_OK;
    }
    void release() { if (p) guarded_free(p); p = nullptr; n = 0; }
};

struct orbx_extractor {
    // ---- parameters and tables of ORBextractor::ORBextractor (ORBextractor.cc:492-609) ----
    int nfeatures; double scaleFactor; int nlevels, iniThFAST, minThFAST, device;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<int> mnFeaturesPerLevel, umax;
    cudaStream_t stream = nullptr;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;   // copy engines of the host-pointer batch call (overlap with compute)
    cudaStream_t s_alt = nullptr;                    // second compute stream: odd chunks of the host batch call run here, so the
                                                     // latency-bound quadtree kernels of one chunk overlap the stencils of the next
    cudaStream_t s_more[2] = {nullptr, nullptr};     // third / fourth compute stream of the host batch call
    cudaStream_t cur = nullptr;                      // stream the run_* helpers launch on (stream or s_alt)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_block = nullptr;   // ev_block: blocking-sync event the host pipeline sleeps on
    std::vector<cudaEvent_t> ev_h2d, ev_done;        // one pair per chunk
    long long launches = 0;

    // ---- geometry plan for (rows, cols) ----
    int rows = 0, cols = 0;
    std::vector<LevelGeom> levels; std::vector<CellDesc> cells; std::vector<BlurTile> tiles, tiles_s;   // tiles_s: BLUR_STRIP_SMALL tiling
    long long pyr_fstride = 0; int cand_per_frame = 0, kp_per_frame = 0, max_kp = 0;
    FastLayout fast_lay{}; int tree_cap = 0, sort_smem_keys = 4096;
    QfPlan qf{}, qfb{}; QfLevels qf_levels{}; bool qf_ok = false, qfb_ok = false;                       // k_octree_fused.cuh: the one-launch quadtree of the latency form
    DevBuf<CUtensorMap> d_tmaps; const void* tmaps_base = nullptr; int tmaps_B = 0;          // FAST tensor maps of levels >= 1 (by level), valid for (d_pyr.p, Bcap)
    CUtensorMap map_l0, map_l0_blur, map_l0_blur_s, map_l0_resize; const void* map_l0_sig[4] = {nullptr};                                // level-0 map of the current view (pointer, frame stride, pitch, frames)
    DevBuf<LevelGeom> d_levels; DevBuf<CellDesc> d_cells; DevBuf<BlurTile> d_tiles, d_tiles_s; DevBuf<int> d_tabs;
    std::vector<ResizeTabs> resize_tabs;
    DevBuf<TilePyrLevel> d_tp_lv; DevBuf<int> d_tp_xr, d_tp_yr; int tp_nx = 0, tp_ny = 0, tp_smem = 0; bool tp_ok = false;   // k_pyr_tiles plan
    DevBuf<ChainLevel> d_chain; bool chain_ok = false;   // k_pyr_chain's level table (all levels in the word-load form)
    std::vector<int> resize_bw, resize_bh;           // TMA box of the source tile per destination level (k_pyr_resize_t); 0 = use the per-thread kernels

    // ---- per-batch device state (the "stateful extractor": pyramid stays resident) ----
    int Bcap = 0, lastB = 0;
    DevBuf<uint8_t> d_pyr, d_blur, d_l0;
    DevBuf<uint32_t> d_slots, d_ocand, d_spk, d_kp_level, d_codetab;
    DevBuf<unsigned long long> d_skey;
    DevBuf<uint16_t> d_cell_counts;
    DevBuf<int> d_ncand, d_kp_count, d_counts, d_level_counts, d_overflow;
    DevBuf<KpOut> d_kp_out; DevBuf<uint8_t> d_desc_out; int out_cap = 0;
    uint32_t* h_bits = nullptr; size_t h_bits_cap = 0;   // pinned staging of host-packed masks (host_pack.cpp), words
    DevBuf<uint8_t> d_gather; uint8_t* h_gather = nullptr; size_t h_gather_cap = 0;   // single-frame result block and its pinned landing buffer
    const KpOut* lb_kp = nullptr; const uint8_t* lb_desc = nullptr; const int* lb_counts = nullptr; int lb_cap = 0, lb_B = 0;   // outputs of the last batched call (orbx_compute_stereo_matches_batch)
    const KpOut* last_kp = nullptr; const uint8_t* last_desc = nullptr; int last_n = -1;   // frame 0 of the last extract / describe (orbx_frame_assign)
    PyrView view{}; bool have_pyramid = false, blur_valid = false;
    // single-frame operator(): the whole per-geometry chain (upload from a pinned staging frame, 7 resizes, FAST, quadtree, blur on the
    // second stream, orientation + descriptors, result gather, download) captured once as a CUDA graph and replayed per call
    cudaGraphExec_t graph1 = nullptr; int graph_launches = 0; uint8_t* h_in = nullptr; size_t h_in_cap = 0; const void* graph_sig[28] = {nullptr};
    // optional per-stage CUDA-event timing (bench.py's roofline): one event set per profiled call
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;     // ORBX_NSTAGES+1 events per call
    // small staging buffers for the single-frame calls
    DevBuf<KpOut> d_kp_tmp; DevBuf<uint8_t> d_desc_tmp; DevBuf<uint8_t> d_mask; DevBuf<uint32_t> d_bits0, d_bits1; DevBuf<double> d_label; DevBuf<int> d_ids, d_culled; DevBuf<uint16_t> d_label16; DevBuf<uint8_t> d_lflags;
};

// -------------------------------------------------------------------------------------------------
// cv::resize(INTER_LINEAR) coefficient tables, padded to `padded` entries
// -------------------------------------------------------------------------------------------------
static void linear_coefs(int ssize, int dsize, int padded, std::vector<int>& ofs, std::vector<short>& w) {
    ofs.assign(padded, 0); w.assign((size_t)padded * 2, 0);
    const double inv_scale = (double)dsize / ssize, scale = 1.0 / inv_scale;
    for (int d = 0; d < padded; ++d) {
        const int dd = std::min(d, dsize - 1);
        float f = (float)((dd + 0.5) * scale - 0.5);
        int s = cv_floor_f(f);
        f -= s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        ofs[d] = s;
        w[2 * d] = (short)cv_round_f((1.f - f) * 2048.f);
        w[2 * d + 1] = (short)cv_round_f(f * 2048.f);
    }
}

// path-code tables of one level (k_octree.cuh): nx entries for x, then ny entries for y, appended to `out`.  Same float arithmetic as octree_code
// (IEEE single division / product, truncation), so a tabulated code equals the computed one.
static void octree_code_tables(const LevelGeom& g, int nx, int ny, std::vector<uint32_t>& out) {
    for (int x = 0; x < nx; ++x) {
        const volatile float q = (float)x / g.hX;                              // vpIniNodes[kp.pt.x / hX]   ORBextractor.cc:766
        const int root = (int)q;
        const volatile float a = g.hX * (float)root, b = g.hX * (float)(root + 1);   // UL.x = (int)(hX*i), UR.x = (int)(hX*(i+1))   :741-745
        out.push_back(((uint32_t)root << ORBX_ROOT_SHIFT) | octree_code_half(x, (int)a, (int)b, 0));
    }
    for (int y = 0; y < ny; ++y) out.push_back(octree_code_half(y, 0, g.maxBY - g.minBY, 1));
}

static int build_plan(orbx_extractor* h, int rows, int cols) {
    if (h->rows == rows && h->cols == cols) return ORBX_OK;
    const int L = h->nlevels;
    std::vector<LevelGeom> lv(L);
    std::vector<CellDesc> cells; std::vector<BlurTile> tiles, tiles_s;
    long long off = 0; int cand_off = 0, kp_off = 0, patch_cap = 0, s_cap = 0, wq_words = 1, zh_max = 1, tree_cap = 0;
    for (int l = 0; l < L; ++l) {
        LevelGeom& g = lv[l];
        std::memset(&g, 0, sizeof(g));
        const float inv = h->mvInvScaleFactor[l];
        g.w = cv_round_f((float)cols * inv); g.h = cv_round_f((float)rows * inv);     // ORBextractor.cc:1834
        if (g.w > ORBX_MAX_DIM || g.h > ORBX_MAX_DIM) FAIL(ORBX_E_INVALID, "image larger than 4095 px is not supported");
        g.pitch = align_up(g.w, 128);
        g.off = off; off += (long long)g.pitch * g.h;
        g.scale = h->mvScaleFactor[l];
        g.kp_size = (float)(int)(31 * h->mvScaleFactor[l]);                             // :1175
        g.N = h->mnFeaturesPerLevel[l];
        // cell grid :1067-1086
        g.minBX = ORBX_EDGE - 3; g.minBY = g.minBX; g.maxBX = g.w - ORBX_EDGE + 3; g.maxBY = g.h - ORBX_EDGE + 3;
        const float W = 30;
        const float width = (float)(g.maxBX - g.minBX), height = (float)(g.maxBY - g.minBY);
        if (g.w < 1 || g.h < 1) FAIL(ORBX_E_INVALID, "image too small: a pyramid level is empty");
        int nCols = (int)(width / W), nRows = (int)(height / W);
        // a level narrower than one 30-px cell has no detection cells: the reference's cell loops do not run and the
        // level contributes no keypoints (ORBextractor.cc:1082-1089)
        if (nCols <= 0 || nRows <= 0) nCols = nRows = 0;
        const int wCell = nCols ? (int)std::ceil(width / nCols) : 0, hCell = nRows ? (int)std::ceil(height / nRows) : 0;
        g.cell_begin = (int)cells.size(); g.slot_off = cand_off; g.cand_off = cand_off;
        int slot = cand_off, bw_max = 0, ch_max = 0;
        for (int i = 0; i < nRows; i++) {
            const float iniY = (float)(g.minBY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= g.maxBY - 3) continue;                                          // :1099
            if (maxY > g.maxBY) maxY = (float)g.maxBY;
            for (int j = 0; j < nCols; j++) {
                const float iniX = (float)(g.minBX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= g.maxBX - 6) continue;                                      // :1116
                if (maxX > g.maxBX) maxX = (float)g.maxBX;
                CellDesc c; std::memset(&c, 0, sizeof(c));
                c.x0 = (short)iniX; c.y0 = (short)iniY; c.cw = (short)((int)maxX - (int)iniX); c.ch = (short)((int)maxY - (int)iniY);
                if (c.cw < 7 || c.ch < 7) continue;                                     // cv::FAST returns nothing
                if (c.cw - 6 >= 64 || c.ch - 6 >= 64) FAIL(ORBX_E_INVALID, "unsupported cell size");
                c.sx = (short)(j * wCell); c.sy = (short)(i * hCell); c.level = (short)l;
                const int zw = c.cw - 6, zh = c.ch - 6;
                c.cap = (short)(((zw + 1) / 2) * ((zh + 1) / 2));
                c.slot = slot; slot += c.cap;
                cells.push_back(c);
                // k_fast.cuh: zone word columns, row strips per warp, rows per strip
                // (the patch starts at the 16-byte boundary left of x0 - 1: TMA's innermost coordinate must be 16-byte aligned)
                const int pc0 = c.x0 + 3 - ((c.x0 - 1) & ~15), wi0 = pc0 >> 2, wi1 = (pc0 + zw - 1) >> 2;
                const int nwz = wi1 - wi0 + 1, strips = 32 / nwz, rps = (zh + strips - 1) / strips;
                cells.back().geo = nwz | (rps << 8) | (strips << 16); cells.back().rcp = 65536 / nwz + 1;
                bw_max = std::max(bw_max, 4 * (wi1 + 2)); ch_max = std::max(ch_max, (int)c.ch);
                s_cap = std::max(s_cap, (zw + 2) * (zh + 2)); wq_words = std::max(wq_words, nwz * zh); zh_max = std::max(zh_max, zh);
            }
        }
        g.cell_count = (int)cells.size() - g.cell_begin;
        // TMA box of the level's cells: ROI from the 16-byte boundary on its left to the right neighbour of its last zone word
        // -- as an ODD number of 16-byte units: the patch pitch is then an odd multiple of 4 banks, so that a column of the patch walks over 8
        // different bank groups instead of 2 (pitch 64) and the ring gathers of vertically adjacent survivors do not collide
        g.fast_bw = align_up(std::max(bw_max, 16), 16); { static const int odd = [] { const char* e = std::getenv("ORBX_FAST_ODD"); return e ? std::atoi(e) : 0; }(); if (odd && (g.fast_bw / 16) % 2 == 0) g.fast_bw += 16; }
        g.fast_bh = std::max(ch_max, 1);
        patch_cap = std::max(patch_cap, g.fast_bw * g.fast_bh);
        g.cand_cap = slot - cand_off; cand_off = slot;
        if (g.cand_cap >= (1 << 20)) FAIL(ORBX_E_INVALID, "level too large");
        // quadtree roots :719-722
        if (g.maxBY - g.minBY == 0) FAIL(ORBX_E_INVALID, "a pyramid level is exactly 32 px high: the reference's nIni = width/0 is undefined");
        g.nIni = (int)std::round(static_cast<float>(g.maxBX - g.minBX) / (g.maxBY - g.minBY));
        if (g.cell_count == 0) {
            // no candidates can exist on this level; the reference still evaluates vpIniNodes.resize(nIni) and dies on nIni < 0
            if (g.nIni < 0) FAIL(ORBX_E_INVALID, "degenerate level geometry (negative nIni in the reference)");
            g.nIni = 1;
        }
        if (g.nIni < 1 || g.nIni > 15) FAIL(ORBX_E_INVALID, "unsupported aspect ratio (nIni must be 1..15)");
        g.hX = static_cast<float>(g.maxBX - g.minBX) / g.nIni;
        g.kp_off = kp_off; g.kp_cap = std::max(g.N + 2, 4 * g.nIni) + 2; kp_off += g.kp_cap;
        tree_cap = std::max(tree_cap, g.kp_cap + 8);
        for (int st = 0; st < (g.h + BLUR_STRIP - 1) / BLUR_STRIP; ++st)
            for (int xc = 0; xc < (g.w + BLUR_TILE_W - 1) / BLUR_TILE_W; ++xc) { BlurTile t; t.level = (short)l; t.xc = (short)xc; t.strip = (short)st; t.pad = 0; tiles.push_back(t); }
        for (int st = 0; st < (g.h + BLUR_STRIP_SMALL - 1) / BLUR_STRIP_SMALL; ++st)
            for (int xc = 0; xc < (g.w + BLUR_TILE_W - 1) / BLUR_TILE_W; ++xc) { BlurTile t; t.level = (short)l; t.xc = (short)xc; t.strip = (short)st; t.pad = 0; tiles_s.push_back(t); }
    }
    if (tree_cap > 32000) FAIL(ORBX_E_INVALID, "too many features per level");
    // resize tables for levels >= 1
    std::vector<int> tabs; std::vector<size_t> tab_off((size_t)L * 4, 0), xg_off(L, 0); std::vector<int> wide_ok(L, 0), rbw(L, 0), rbh(L, 0);
    for (int l = 1; l < L; ++l) {
        std::vector<int> xo, yo; std::vector<short> xw, yw;
        linear_coefs(lv[l - 1].w, lv[l].w, lv[l].pitch, xo, xw);
        linear_coefs(lv[l - 1].h, lv[l].h, align_up(lv[l].h, 8), yo, yw);
        auto push = [&](const void* p, size_t bytes) { size_t o = tabs.size(); tabs.resize(o + (bytes + 15) / 16 * 4); std::memcpy(&tabs[o], p, bytes); return o; };
        tab_off[l * 4 + 0] = push(xo.data(), xo.size() * 4); tab_off[l * 4 + 1] = push(xw.data(), xw.size() * 2);
        tab_off[l * 4 + 2] = push(yo.data(), yo.size() * 4); tab_off[l * 4 + 3] = push(yw.data(), yw.size() * 2);
        // word-load form: per group of 4 destination columns, first source column + PRMT selectors of the (left,right) pairs
        const int ngroups = lv[l].pitch / 4, sw = lv[l - 1].w;
        std::vector<int> xg((size_t)ngroups * 2);
        bool wide = true;
        for (int gx = 0; gx < ngroups; ++gx) {
            const int base = xo[4 * gx];
            unsigned sel = 0;
            for (int k = 0; k < 4; ++k) {
                const int il = xo[4 * gx + k] - base, ir = std::min(xo[4 * gx + k] + 1, sw - 1) - base;
                if (il < 0 || il > 7 || ir < 0 || ir > 7) wide = false;
                sel |= ((unsigned)(il & 7) | ((unsigned)(ir & 7) << 4)) << (8 * k);
            }
            xg[2 * gx] = base; xg[2 * gx + 1] = (int)sel;
        }
        xg_off[l] = push(xg.data(), xg.size() * 4); wide_ok[l] = wide && (lv[l - 1].pitch % 4 == 0);
        // k_pyr_resize_t: source rectangle of a 128 x RESIZE_ROWS destination tile, from the 16-byte boundary left of its first source column
        int bwm = 0, bhm = 0;
        for (int g0 = 0; g0 < ngroups; g0 += 32) {
            const int bx0 = xg[2 * g0] & ~15;
            for (int gx = g0; gx < std::min(g0 + 32, ngroups); ++gx) bwm = std::max(bwm, 4 * (((xg[2 * gx] - bx0) >> 2) + 3));
        }
        for (int y0 = 0; y0 < lv[l].h; y0 += RESIZE_ROWS) {
            const int y1 = std::min(y0 + RESIZE_ROWS, lv[l].h) - 1;
            bhm = std::max(bhm, std::min(yo[y1] + 1, lv[l - 1].h - 1) - yo[y0] + 1);
        }
        bwm = align_up(bwm, 16); if ((bwm / 16) % 2 == 0) bwm += 16;            // odd number of 16-byte units: rows of a column spread over the banks
        const long long per_warp = ((long long)bwm * bhm + 16 + 127) / 128 * 128;
        rbw[l] = rbh[l] = 0;
        if (wide_ok[l] && bwm <= 256 && bhm <= 256 && per_warp * RESIZE_WARPS <= 100 * 1024) { rbw[l] = bwm; rbh[l] = bhm; }
    }
    {   // quadtree path-code tables, one (x, y) pair per level
        std::vector<uint32_t> ct; std::vector<size_t> at(L);
        for (int l = 0; l < L; ++l) {
            at[l] = ct.size();
            lv[l].code_nx = std::max(lv[l].maxBX - lv[l].minBX, 1) + 8; lv[l].code_ny = std::max(lv[l].maxBY - lv[l].minBY, 1) + 8;
            octree_code_tables(lv[l], lv[l].code_nx, lv[l].code_ny, ct);
        }
        if (h->d_codetab.ensure(ct.size())) return ORBX_E_CUDA;
        CU_TRY(cudaMemcpyAsync(h->d_codetab.p, ct.data(), ct.size() * 4, cudaMemcpyHostToDevice, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        for (int l = 0; l < L; ++l) { lv[l].code_x = h->d_codetab.p + at[l]; lv[l].code_y = lv[l].code_x + lv[l].code_nx; }
    }
    if (h->d_levels.ensure(L)) return ORBX_E_CUDA;
    if (h->d_cells.ensure(cells.size())) return ORBX_E_CUDA;
    if (h->d_tiles.ensure(tiles.size()) || h->d_tiles_s.ensure(tiles_s.size())) return ORBX_E_CUDA;
    if (h->d_tabs.ensure(std::max<size_t>(tabs.size(), 4))) return ORBX_E_CUDA;
    CU_TRY(cudaMemcpyAsync(h->d_levels.p, lv.data(), sizeof(LevelGeom) * L, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(h->d_cells.p, cells.data(), sizeof(CellDesc) * cells.size(), cudaMemcpyHostToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(h->d_tiles.p, tiles.data(), sizeof(BlurTile) * tiles.size(), cudaMemcpyHostToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(h->d_tiles_s.p, tiles_s.data(), sizeof(BlurTile) * tiles_s.size(), cudaMemcpyHostToDevice, h->stream));
    if (!tabs.empty()) CU_TRY(cudaMemcpyAsync(h->d_tabs.p, tabs.data(), tabs.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));   // host vectors go out of scope below
    h->resize_tabs.assign(L, ResizeTabs());
    for (int l = 1; l < L; ++l) {
        h->resize_tabs[l].xofs = h->d_tabs.p + tab_off[l * 4 + 0];
        h->resize_tabs[l].xw = reinterpret_cast<const short2*>(h->d_tabs.p + tab_off[l * 4 + 1]);
        h->resize_tabs[l].yofs = h->d_tabs.p + tab_off[l * 4 + 2];
        h->resize_tabs[l].yw = reinterpret_cast<const short2*>(h->d_tabs.p + tab_off[l * 4 + 3]);
        h->resize_tabs[l].xg = reinterpret_cast<const int2*>(h->d_tabs.p + xg_off[l]);
        h->resize_tabs[l].wide = wide_ok[l];
    }
    {   // k_pyr_chain: per-level geometry + tables in one device array
        std::vector<ChainLevel> ch((size_t)L); std::memset(ch.data(), 0, sizeof(ChainLevel) * (size_t)L);
        bool ok = L > 1;
        for (int l = 1; l < L; ++l) {
            ChainLevel& c = ch[l];
            c.sw = lv[l - 1].w; c.sh = lv[l - 1].h; c.spitch = lv[l - 1].pitch; c.soff = lv[l - 1].off;
            c.dw = lv[l].w; c.dh = lv[l].h; c.dpitch = lv[l].pitch; c.doff = lv[l].off; c.t = h->resize_tabs[l];
            ok = ok && h->resize_tabs[l].wide;
        }
        if (h->d_chain.ensure((size_t)L)) return ORBX_E_CUDA;
        CU_TRY(cudaMemcpyAsync(h->d_chain.p, ch.data(), sizeof(ChainLevel) * (size_t)L, cudaMemcpyHostToDevice, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        h->chain_ok = ok;
    }
    {   // k_pyr_tiles: one CTA per 64 x 64 rectangle of level 0; per level its owned rectangle and the (slightly larger) one it computes
        const int TP = 64;
        const int nx = (cols + TP - 1) / TP, ny = (rows + TP - 1) / TP;
        bool ok = L > 1;
        for (int l = 1; l < L; ++l) ok = ok && h->resize_tabs[l].wide;
        std::vector<std::vector<int>> xo(L), yo(L);
        for (int l = 1; l < L && ok; ++l) { std::vector<short> w; linear_coefs(lv[l - 1].w, lv[l].w, lv[l].pitch, xo[l], w); linear_coefs(lv[l - 1].h, lv[l].h, align_up(lv[l].h, 8), yo[l], w); }
        std::vector<int> xr((size_t)L * nx * 4, 0), yr((size_t)L * ny * 4, 0);
        auto X = [&](int l, int i) { return &xr[((size_t)l * nx + i) * 4]; };
        auto Y = [&](int l, int j) { return &yr[((size_t)l * ny + j) * 4]; };
        for (int l = 1; l < L && ok; ++l) {
            const int G = (lv[l].w + 3) / 4;
            for (int i = 0; i < nx; ++i) { int* r = X(l, i); r[2] = (int)((long long)i * G / nx); r[3] = (int)((long long)(i + 1) * G / nx); r[0] = r[2]; r[1] = r[3]; }
            for (int j = 0; j < ny; ++j) { int* r = Y(l, j); r[2] = (int)((long long)j * lv[l].h / ny); r[3] = (int)((long long)(j + 1) * lv[l].h / ny); r[0] = r[2]; r[1] = r[3]; }
        }
        for (int l = L - 1; l >= 2 && ok; --l) {                      // what level l - 1 must hold for this CTA's region of level l
            for (int i = 0; i < nx; ++i) {
                const int* d = X(l, i); int* sr = X(l - 1, i);
                if (d[1] > d[0]) {
                    const int lo = xo[l][4 * d[0]] / 4, hi = std::min(xo[l][4 * d[1] - 1] + 1, lv[l - 1].w - 1) / 4 + 1;
                    if (sr[1] > sr[0]) { sr[0] = std::min(sr[0], lo); sr[1] = std::max(sr[1], hi); } else { sr[0] = lo; sr[1] = hi; }
                }
            }
            for (int j = 0; j < ny; ++j) {
                const int* d = Y(l, j); int* sr = Y(l - 1, j);
                if (d[1] > d[0]) {
                    const int lo = yo[l][d[0]], hi = std::min(yo[l][d[1] - 1] + 1, lv[l - 1].h - 1) + 1;
                    if (sr[1] > sr[0]) { sr[0] = std::min(sr[0], lo); sr[1] = std::max(sr[1], hi); } else { sr[0] = lo; sr[1] = hi; }
                }
            }
        }
        std::vector<TilePyrLevel> tl((size_t)L); std::memset(tl.data(), 0, sizeof(TilePyrLevel) * (size_t)L);
        int off = 0;
        for (int l = 1; l < L && ok; ++l) {
            int gw = 0, gh = 0;
            for (int i = 0; i < nx; ++i) gw = std::max(gw, X(l, i)[1] - X(l, i)[0]);
            for (int j = 0; j < ny; ++j) gh = std::max(gh, Y(l, j)[1] - Y(l, j)[0]);
            TilePyrLevel& c = tl[l];
            c.sw = lv[l - 1].w; c.sh = lv[l - 1].h; c.dw = lv[l].w; c.dh = lv[l].h; c.dpitch = lv[l].pitch; c.spitch = lv[l - 1].pitch; c.doff = lv[l].off;
            c.sm_off = off; c.sm_pitch = gw * 4; c.t = h->resize_tabs[l];
            off += align_up(std::max(gw * 4 * gh, 16), 16);
            c.tab_rows = align_up(std::max(gh, 1), 4); c.tab_groups = align_up(std::max(gw, 1), 2);      // keeps the int2 / uint4 slices aligned
            c.tab_off = off; off += align_up(c.tab_rows * 8 + c.tab_groups * 24, 16);
        }
        ok = ok && off <= 200 * 1024;
        if (ok) {
            if (h->d_tp_lv.ensure((size_t)L) || h->d_tp_xr.ensure(xr.size()) || h->d_tp_yr.ensure(yr.size())) return ORBX_E_CUDA;
            CU_TRY(cudaMemcpyAsync(h->d_tp_lv.p, tl.data(), sizeof(TilePyrLevel) * (size_t)L, cudaMemcpyHostToDevice, h->stream));
            CU_TRY(cudaMemcpyAsync(h->d_tp_xr.p, xr.data(), xr.size() * 4, cudaMemcpyHostToDevice, h->stream));
            CU_TRY(cudaMemcpyAsync(h->d_tp_yr.p, yr.data(), yr.size() * 4, cudaMemcpyHostToDevice, h->stream));
            CU_TRY(cudaStreamSynchronize(h->stream));
        }
        h->tp_ok = ok; h->tp_nx = nx; h->tp_ny = ny; h->tp_smem = off;
    }
    h->levels.swap(lv); h->cells.swap(cells); h->tiles.swap(tiles); h->tiles_s.swap(tiles_s); h->resize_bw.swap(rbw); h->resize_bh.swap(rbh);
    h->pyr_fstride = (off + 255) / 256 * 256;
    h->cand_per_frame = cand_off; h->kp_per_frame = kp_off;
    h->max_kp = 0; for (int l = 0; l < L; ++l) h->max_kp += h->levels[l].kp_cap;
    {   // [patch | S | wq | ring | clist | bm | mbarrier] per warp (k_fast.cuh)
        FastLayout& f = h->fast_lay;
        f.patch_cap = align_up(patch_cap, 128); f.s_cap = align_up(s_cap, 16); f.wq_cap = align_up(4 * wq_words, 16); f.bm_cap = align_up(8 * zh_max, 16);
        static const int padx = [] { const char* e = std::getenv("ORBX_FAST_PAD"); return e ? std::atoi(e) : 0; }();
        f.per_warp = align_up(f.patch_cap + f.s_cap + f.wq_cap + 2 * FAST_RING + 2 * FAST_CLIST + f.bm_cap + 16 + padx, 128);
        f.cta_shared = align_up(f.patch_cap + f.s_cap + f.bm_cap + 32, 128); f.cta_per_warp = align_up(f.wq_cap + 2 * FAST_RING + 2 * FAST_CLIST, 16);
        if ((long long)f.per_warp * FAST_WARPS > 227 * 1024) FAIL(ORBX_E_INVALID, "FAST cells too large for shared memory");
    }
    h->tmaps_base = nullptr; h->map_l0_sig[0] = nullptr;
    h->tree_cap = tree_cap;
    // radix-sort capacity per (level, frame): ~1 candidate per 100 px of level 0 is generous for real images; larger levels fall back
    // to the global-memory bitonic path inside the kernel
    { const long long want = (long long)rows * cols / 100; int k = 4096; if (want > 4096) k = (int)std::min<long long>(18432, (want + want / 4 + 2047) / 2048 * 2048); h->sort_smem_keys = k; }
    // one-launch quadtree (k_octree_fused.cuh): node pool of the list's worst case (N + 3 nodes), the rest of the shared memory for keys; levels with more
    // candidates than that run the same code on global scratch.
    {
        int cellmax = 1, tabmax = 0;
        for (int l = 0; l < L; ++l) { cellmax = std::max(cellmax, h->levels[l].cell_count); tabmax = std::max(tabmax, h->levels[l].code_nx + h->levels[l].code_ny); }
        QfPlan q{}; q.threads = QF_THREADS; q.pool_cap = std::max(1056, align_up(tree_cap, 32)); q.cell_cap = cellmax; q.tab_cap = align_up(tabmax, 4);
        const size_t budget = 224 * 1024, fixed = qf_fixed_bytes(q.pool_cap, q.cell_cap, q.tab_cap, q.threads);
        h->qf_ok = L <= QF_MAXLEVELS && q.pool_cap <= QF_MAXPOOL && fixed + 2048 * 16 <= budget;
        if (h->qf_ok) for (int l = 0; l < L; ++l) h->qf_levels.lv[l] = h->levels[l];
        if (h->qf_ok) {
            q.key_cap = (int)std::min<size_t>(8192, ((budget - fixed) / 16) & ~(size_t)31);
            q.smem_bytes = (int)qf_smem_bytes(q);
        }
        h->qf = q;
        // batched form: lean CTAs (QF_THREADS_BATCH threads).  Measured on 1024 VGA frames (quadtree stage, ms): key_cap 4096 (2 CTAs per SM) 0.59 = the
        // sort + tree pair, 2048 0.42, 1024 0.41, 512 without staged code tables 0.40: residency beats shared-memory keys (the global scratch of a level is
        // L2-resident), so the batched form keeps only the small levels' keys in shared memory.
        QfPlan qb{}; qb.threads = QF_THREADS_BATCH; qb.pool_cap = std::max(align_up(QF_THREADS_BATCH / 32 * 33, 32), align_up(tree_cap, 32)); qb.cell_cap = cellmax; qb.tab_cap = 0;
        { static const int btab = [] { const char* e = std::getenv("ORBX_QT_BTAB"); return e ? std::atoi(e) : 0; }(); if (btab) qb.tab_cap = q.tab_cap; }
        static const int bkeys = [] { const char* e = std::getenv("ORBX_QT_BKEYS"); return e ? std::atoi(e) : 0; }();
        const size_t bbudget = 112 * 1024, bfixed = qf_fixed_bytes(qb.pool_cap, qb.cell_cap, qb.tab_cap, qb.threads);
        h->qfb_ok = h->qf_ok && bfixed + 1024 * 16 <= bbudget;
        if (h->qfb_ok) {
            qb.key_cap = bkeys >= 256 ? (int)std::min<size_t>(bkeys & ~31, ((bbudget - bfixed) / 16) & ~(size_t)31) : 512;
            qb.smem_bytes = (int)qf_smem_bytes(qb);
        }
        h->qfb = qb;
    }
    h->rows = rows; h->cols = cols; h->Bcap = 0; h->have_pyramid = false;
    if (h->graph1) { cudaGraphExecDestroy(h->graph1); h->graph1 = nullptr; }       // the captured pointers / launch shapes belong to the old geometry
    return ORBX_OK;
}

static int ensure_capacity(orbx_extractor* h, int B, int out_cap) {
    h->last_n = -1; h->lb_B = 0;                                 // every detect / extract path passes here: the previous result is about to be overwritten
    if (B > h->Bcap) {
        const size_t b = (size_t)B;
        if (h->d_pyr.ensure(b * h->pyr_fstride) || h->d_blur.ensure(b * h->pyr_fstride) ||
            h->d_slots.ensure(b * h->cand_per_frame) || h->d_ocand.ensure(b * h->cand_per_frame) ||
            h->d_spk.ensure(b * h->cand_per_frame) || h->d_skey.ensure(b * h->cand_per_frame) ||
            h->d_cell_counts.ensure(b * h->cells.size()) || h->d_ncand.ensure(b * h->nlevels) ||
            h->d_kp_level.ensure(b * h->kp_per_frame) || h->d_kp_count.ensure(b * h->nlevels) ||
            h->d_counts.ensure(b) || h->d_level_counts.ensure(b * h->nlevels) || h->d_overflow.ensure(4))
            return ORBX_E_CUDA;
        CU_TRY(cudaMemsetAsync(h->d_overflow.p, 0, 16, h->stream));
        h->Bcap = B;
        if (h->graph1) { cudaGraphExecDestroy(h->graph1); h->graph1 = nullptr; }   // buffers may have moved
    }
    if (out_cap > 0 && ((size_t)B * out_cap > h->d_kp_out.n)) {
        if (h->d_kp_out.ensure((size_t)B * out_cap) || h->d_desc_out.ensure((size_t)B * out_cap * 32)) return ORBX_E_CUDA;
        if (h->graph1) { cudaGraphExecDestroy(h->graph1); h->graph1 = nullptr; }
    }
    return ORBX_OK;
}

#define ORBX_NSTAGES 6   // resize, fast_cells, octree_sort, octree_tree, gauss7, orient_describe
static inline void prof_mark(orbx_extractor* h) {
    if (!h->profiling) return;
    cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, h->cur); h->prof_events.push_back(e);
}

#define LAUNCH_CHECK() do { cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) { \
    orbx_set_error(std::string("kernel launch: ") + cudaGetErrorString(_e)); return ORBX_E_CUDA; } ++h->launches; } while (0)

static int run_blur_range(orbx_extractor* h, int b0, int B);

// re-pitch B frames on the device in one launch (k_cull.cuh); dst_pitch must be a multiple of 4
static int repitch_frames(orbx_extractor* h, uint8_t* dst, long long dst_fstride, int dst_pitch, const uint8_t* src, long long src_fstride, long long src_step, int cols, int rows, int B, cudaStream_t s) {
    if (B <= 0) return ORBX_OK;
    k_repitch<<<dim3((cols + 1023) / 1024, rows, B), 256, 0, s>>>(src, src_fstride, src_step, dst, dst_fstride, dst_pitch, cols, rows);
    LAUNCH_CHECK();
    return ORBX_OK;
}

// Tensor maps of the FAST and blur stages (k_fast.cuh, k_describe.cuh).  Levels >= 1 live in d_pyr: one map per level in a device array, rebuilt when the pyramid
// block moves.  Level 0 follows the current view (internal copy, pinned mirror or the caller's device frames) and travels as a kernel
// parameter.  Must not be first called inside a stream capture (extract_graph prepares before capturing).
static int fast_prepare(orbx_extractor* h) {
    const int L = h->nlevels;
    const long long frames = 1 << 16;                      // bound of the frame coordinate only; kernels index frames < B
    if (h->tmaps_base != (const void*)h->d_pyr.p || !h->d_tmaps.p) {
        std::vector<CUtensorMap> m((size_t)4 * L);                       // [0, L): FAST cells, [L, 2L): blur tiles, [2L, 3L): resize source tiles (level l reads l - 1), [3L, 4L): small blur tiles
        std::memset(m.data(), 0, sizeof(CUtensorMap) * (size_t)4 * L);
        for (int l = 1; l < L; ++l) {
            const LevelGeom& g = h->levels[l];
            if (g.cell_count && !orbx_tmap_image(&m[l], h->d_pyr.p + g.off, g.w, g.h, frames, g.pitch, h->pyr_fstride, g.fast_bw, g.fast_bh)) FAIL(ORBX_E_CUDA, "cuTensorMapEncodeTiled failed (pyramid level)");
            if (!orbx_tmap_image(&m[L + l], h->d_pyr.p + g.off, g.w, g.h, frames, g.pitch, h->pyr_fstride, BLUR_BOX_W, BLUR_STRIP + 6) ||
                !orbx_tmap_image(&m[3 * L + l], h->d_pyr.p + g.off, g.w, g.h, frames, g.pitch, h->pyr_fstride, BLUR_BOX_W, BLUR_STRIP_SMALL + 6)) FAIL(ORBX_E_CUDA, "cuTensorMapEncodeTiled failed (pyramid level, blur)");
            if (l + 1 < L && h->resize_bw[l + 1] && !orbx_tmap_image(&m[2 * L + l + 1], h->d_pyr.p + g.off, g.w, g.h, frames, g.pitch, h->pyr_fstride, h->resize_bw[l + 1], h->resize_bh[l + 1]))
                FAIL(ORBX_E_CUDA, "cuTensorMapEncodeTiled failed (pyramid level, resize)");
        }
        if (h->d_tmaps.ensure((size_t)4 * L)) return ORBX_E_CUDA;
        CU_TRY(cudaMemcpyAsync(h->d_tmaps.p, m.data(), sizeof(CUtensorMap) * (size_t)4 * L, cudaMemcpyHostToDevice, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        h->tmaps_base = h->d_pyr.p;
    }
    const void* sig[4] = {h->view.l0, (const void*)(uintptr_t)h->view.l0_fstride, (const void*)(uintptr_t)h->view.l0_pitch, (const void*)(uintptr_t)(h->rows * 65536 + h->cols)};
    if (std::memcmp(sig, h->map_l0_sig, sizeof(sig)) != 0) {
        const LevelGeom& g = h->levels[0];
        std::memset(&h->map_l0, 0, sizeof(h->map_l0));
        if (g.cell_count && !orbx_tmap_image(&h->map_l0, h->view.l0, g.w, g.h, frames, h->view.l0_pitch, h->view.l0_fstride, g.fast_bw, g.fast_bh))
            FAIL(ORBX_E_CUDA, "cuTensorMapEncodeTiled failed (level 0: frames must be 16-byte aligned in pointer, row step and frame stride)");
        if (!orbx_tmap_image(&h->map_l0_blur, h->view.l0, g.w, g.h, frames, h->view.l0_pitch, h->view.l0_fstride, BLUR_BOX_W, BLUR_STRIP + 6) ||
            !orbx_tmap_image(&h->map_l0_blur_s, h->view.l0, g.w, g.h, frames, h->view.l0_pitch, h->view.l0_fstride, BLUR_BOX_W, BLUR_STRIP_SMALL + 6)) FAIL(ORBX_E_CUDA, "cuTensorMapEncodeTiled failed (level 0, blur)");
        std::memset(&h->map_l0_resize, 0, sizeof(h->map_l0_resize));
        if (L > 1 && h->resize_bw[1] && !orbx_tmap_image(&h->map_l0_resize, h->view.l0, g.w, g.h, frames, h->view.l0_pitch, h->view.l0_fstride, h->resize_bw[1], h->resize_bh[1]))
            FAIL(ORBX_E_CUDA, "cuTensorMapEncodeTiled failed (level 0, resize)");
        std::memcpy(h->map_l0_sig, sig, sizeof(sig));
    }
    return ORBX_OK;
}

// pyramid (levels >= 1) + FAST cells + quadtree; leaves per-level keypoints in d_kp_level / d_kp_count.
// fork_blur: the blur only needs the pyramid, so it is launched on the second compute stream -- after FAST, which fills the GPU
// by itself -- and runs next to the quadtree kernels (latency-bound: one serial warp per level, most SMs idle);
// the caller joins on ev_join before the descriptor stage.
static int run_detect(orbx_extractor* h, int b0, int B, bool fork_blur = false) {
    cudaStream_t s = h->cur;
    const int L = h->nlevels;
    PyrView view = h->view;
    view.l0 += (long long)b0 * view.l0_fstride; view.pyr += (long long)b0 * view.pyr_fstride;
    uint8_t* pyr = h->d_pyr.p + (size_t)b0 * h->pyr_fstride;
    uint32_t* slots = h->d_slots.p + (size_t)b0 * h->cand_per_frame;
    uint16_t* cell_counts = h->d_cell_counts.p + (size_t)b0 * h->cells.size();
    const size_t co = (size_t)b0 * h->cand_per_frame;
    { const int rc = fast_prepare(h); if (rc) return rc; }
    prof_mark(h);
    static const int chain_env = [] { const char* e = std::getenv("ORBX_CHAIN"); return e ? std::atoi(e) : -1; }();
    const bool chain = h->chain_ok && (h->view.l0_pitch & 3) == 0 && (chain_env > 0);
    static const int tp_env = [] { const char* e = std::getenv("ORBX_TILEPYR"); return e ? std::atoi(e) : -1; }();
    const bool tilepyr = !chain && h->tp_ok && (h->view.l0_pitch & 3) == 0 && (tp_env >= 0 ? tp_env != 0 : (long long)B * h->tp_nx * h->tp_ny <= 2048);
    if (tilepyr) {                                                   // a handful of frames: the whole pyramid in ONE launch, seams recomputed instead of synchronised
        TilePyrPlan P; P.lv = h->d_tp_lv.p; P.xr = h->d_tp_xr.p; P.yr = h->d_tp_yr.p; P.L = L; P.nx = h->tp_nx; P.ny = h->tp_ny;
        k_pyr_tiles<<<dim3(h->tp_nx, h->tp_ny, B), TILEPYR_THREADS, h->tp_smem, s>>>(h->view, b0, P);
        LAUNCH_CHECK();
    }
    if (chain) {                                                     // a handful of frames: the whole chain in one launch, one 8-CTA cluster per frame
        k_pyr_chain<<<dim3(CHAIN_CTAS, B), CHAIN_THREADS, 0, s>>>(h->view, b0, h->d_chain.p, L);
        LAUNCH_CHECK();
    }
    for (int l = 1; l < L && !chain && !tilepyr; ++l) {
        const LevelGeom& g = h->levels[l]; const LevelGeom& gp = h->levels[l - 1];
        const uint8_t* src; long long sfs; int sp;
        if (l == 1) { src = view.l0; sfs = view.l0_fstride; sp = view.l0_pitch; }
        else { src = pyr + gp.off; sfs = h->pyr_fstride; sp = gp.pitch; }
        static const int rt_env = [] { const char* e = std::getenv("ORBX_RESIZE_TMA"); return e ? std::atoi(e) : -1; }();
        const long long tile_warps = (long long)((g.w + 127) / 128) * ((g.h + RESIZE_ROWS - 1) / RESIZE_ROWS) * B;
        if (h->resize_bw[l] && (rt_env >= 0 ? rt_env != 0 : tile_warps >= 2048)) {      // few tiles (a handful of frames): the per-thread form spreads them over more SMs
            const int bw = h->resize_bw[l], bh = h->resize_bh[l];
            const int per_warp = (bw * bh + 16 + 127) / 128 * 128;
            dim3 grid((g.w + 127) / 128, ((g.h + RESIZE_ROWS - 1) / RESIZE_ROWS + RESIZE_WARPS - 1) / RESIZE_WARPS, B);
            k_pyr_resize_t<<<grid, RESIZE_WARPS * 32, per_warp * RESIZE_WARPS, s>>>(h->map_l0_resize, l == 1 ? nullptr : h->d_tmaps.p + 2 * L + l, b0, gp.h,
                                                                                     h->d_pyr.p + g.off, h->pyr_fstride, g.pitch, g.w, g.h, h->resize_tabs[l], bw, bh);
        } else {
        dim3 grid((g.w + 127) / 128, (g.h + 7) / 8, B), block(32, 8);
        if (h->resize_tabs[l].wide && (sp & 3) == 0)
            k_pyr_resize_w<<<grid, block, 0, s>>>(src, sfs, sp, gp.w, gp.h, pyr + g.off, h->pyr_fstride, g.pitch, g.w, g.h, h->resize_tabs[l]);
        else
            k_pyr_resize<<<grid, block, 0, s>>>(src, sfs, sp, gp.w, gp.h, pyr + g.off, h->pyr_fstride, g.pitch, g.w, g.h, h->resize_tabs[l]);
        }
        LAUNCH_CHECK();
    }
    prof_mark(h);
    const int ncells = (int)h->cells.size();
    if (ncells > 0) {
        // each warp walks FAST_CELLS_PER_WARP cells of its frame with the next ROI prefetched while the current one is processed
        // (4 cells per warp for throughput; a small batch would leave most SMs idle that way, so it gets one CTA per 4 cells: 148 SMs x 7 CTAs per wave)
        static const int cpw_env = [] { const char* e = std::getenv("ORBX_FAST_CPW"); int v = e ? std::atoi(e) : 0; return v < 0 ? 0 : v; }();
        const long long cells_total = (long long)ncells * B;
        const int cpw = cpw_env ? cpw_env : (cells_total >= 64LL * 2072 ? 8 : (cells_total >= 16LL * 2072 ? 4 : (cells_total >= 8LL * 2072 ? 2 : 1)));
        { const int rc = fast_prepare(h); if (rc) return rc; }
        static const int cta_env = [] { const char* e = std::getenv("ORBX_FAST_CTA"); return e ? std::atoi(e) : -1; }();
        if (cta_env >= 0 ? cta_env != 0 : cells_total <= 4LL * 2072) {   // a handful of frames: one CTA per cell, its four warps split the zone's rows (latency form)
            const int smem = h->fast_lay.cta_shared + FAST_WARPS * h->fast_lay.cta_per_warp;
            k_fast_cells<true><<<dim3(ncells, B), FAST_WARPS * 32, smem, s>>>(h->map_l0, h->d_tmaps.p, b0, h->d_levels.p, h->d_cells.p, ncells, h->cand_per_frame,
                                                                             h->fast_lay, h->iniThFAST, h->minThFAST, slots, cell_counts);
        } else {
            dim3 grid((ncells + FAST_WARPS * cpw - 1) / (FAST_WARPS * cpw), B);
            const int smem = h->fast_lay.per_warp * FAST_WARPS;
            k_fast_cells<false><<<grid, FAST_WARPS * 32, smem, s>>>(h->map_l0, h->d_tmaps.p, b0, h->d_levels.p, h->d_cells.p, ncells, h->cand_per_frame,
                                                                     h->fast_lay, h->iniThFAST, h->minThFAST, slots, cell_counts);
        }
        LAUNCH_CHECK();
    }
    prof_mark(h);
    if (fork_blur) CU_TRY(cudaEventRecord(h->ev_fork, s));        // the blur may start once FAST is done ...
    static const int qf_env = [] { const char* e = std::getenv("ORBX_QT_FUSED"); return e ? std::atoi(e) : -1; }();
    // ORBX_QT_FUSED: 0 = sort + tree pair, 1 = one-launch kernel in its wide (latency) form, 2 = in its lean (batched) form; default: by batch size
    // (a batch of large frames with few (level, frame) instances keeps the wide form: its big levels would crawl on 256 threads)
    const bool wide_qt = B <= 4 || ((long long)h->rows * h->cols > 500000 && (long long)B * L <= 4 * 148);
    if (h->qfb_ok && (qf_env >= 0 ? qf_env == 2 : !wide_qt)) {
        k_octree_fused<QF_THREADS_BATCH><<<dim3(L, B), QF_THREADS_BATCH, h->qfb.smem_bytes, s>>>(h->qf_levels, h->d_cells.p, ncells, h->cand_per_frame, h->cand_per_frame, h->kp_per_frame, L, h->qfb,
            slots, cell_counts, h->d_ocand.p + co, h->d_skey.p + co, h->d_spk.p + co, h->d_ncand.p + (size_t)b0 * L, h->d_kp_level.p + (size_t)b0 * h->kp_per_frame,
            h->d_kp_count.p + (size_t)b0 * L, h->d_overflow.p);
        LAUNCH_CHECK();
        prof_mark(h);                                                    // (the stage table keeps its sort / tree columns: the second one reads 0)
    } else if (h->qf_ok && (qf_env >= 0 ? qf_env != 0 : wide_qt)) {
        // a handful of frames (what Tracking calls): gather + path codes + sort + tree of a level in ONE launch, everything in shared memory
        k_octree_fused<QF_THREADS><<<dim3(L, B), QF_THREADS, h->qf.smem_bytes, s>>>(h->qf_levels, h->d_cells.p, ncells, h->cand_per_frame, h->cand_per_frame, h->kp_per_frame, L, h->qf,
            slots, cell_counts, h->d_ocand.p + co, h->d_skey.p + co, h->d_spk.p + co, h->d_ncand.p + (size_t)b0 * L, h->d_kp_level.p + (size_t)b0 * h->kp_per_frame,
            h->d_kp_count.p + (size_t)b0 * L, h->d_overflow.p);
        LAUNCH_CHECK();
        prof_mark(h);                                                    // (the stage table keeps its sort / tree columns: the second one reads 0)
    } else {
        dim3 grid(L, B);
        // a handful of frames: one CTA per level cannot fill the GPU anyway, so each CTA is made wide and the level-0 sort gets 4x the threads per pass
        static const int wide_env = [] { const char* e = std::getenv("ORBX_SORT_WIDE"); return e ? std::atoi(e) : -1; }();
        const bool wide = (wide_env < 0 ? B <= 4 : wide_env != 0) && h->sort_smem_keys <= 8192;
        if (wide)
            k_octree_sort_t<SORT_THREADS_WIDE><<<grid, SORT_THREADS_WIDE, octree_sort_smem_bytes(h->sort_smem_keys, SORT_THREADS_WIDE), s>>>(
                h->d_levels.p, h->d_cells.p, ncells, h->cand_per_frame, h->cand_per_frame, L, h->sort_smem_keys, slots, cell_counts, h->d_ocand.p + co, h->d_skey.p + co,
                h->d_spk.p + co, h->d_ncand.p + (size_t)b0 * L);
        else
            k_octree_sort_t<SORT_THREADS><<<grid, SORT_THREADS, octree_sort_smem_bytes(h->sort_smem_keys), s>>>(
                h->d_levels.p, h->d_cells.p, ncells, h->cand_per_frame, h->cand_per_frame, L, h->sort_smem_keys, slots, cell_counts, h->d_ocand.p + co, h->d_skey.p + co,
                h->d_spk.p + co, h->d_ncand.p + (size_t)b0 * L);
        LAUNCH_CHECK();
        prof_mark(h);
        // sorted path codes staged in shared memory per (level, frame) instance: the extra 16 KB lowers the number of resident
        // instances per SM, so it only pays while all instances fit in one wave (small batches, latency)
        static const int tree_mode = [] { const char* e = std::getenv("ORBX_TREE"); return e ? std::atoi(e) : 0; }();   // 1 = force the serial kernel (A/B testing)
        static const int cc_env = [] { const char* e = std::getenv("ORBX_TREE_CODECAP"); return e ? std::atoi(e) : -1; }();
        const int code_cap = cc_env >= 0 ? cc_env : 3072;          // 12 KB: covers ~2.4 k candidates of a VGA level 0; larger instances search the global keys
        if (h->tree_cap <= PTREE_MAXCAP && tree_mode != 1) {
            k_octree_tree_par<<<grid, PTREE_THREADS, ptree_smem_bytes(h->tree_cap, code_cap), s>>>(h->d_levels.p, L, h->cand_per_frame, h->kp_per_frame, h->tree_cap, code_cap,
                h->d_skey.p + co, h->d_spk.p + co, h->d_ncand.p + (size_t)b0 * L, h->d_kp_level.p + (size_t)b0 * h->kp_per_frame, h->d_kp_count.p + (size_t)b0 * L, h->d_overflow.p);
        } else {
            const size_t tsm = (((size_t)h->tree_cap * (8 + 8 + 4 + 4 + 4 + 2 + 2 + 2 + 1) + 15) & ~(size_t)15) + (size_t)code_cap * 4 + 16;
            k_octree_tree<<<grid, 32, tsm, s>>>(h->d_levels.p, L, h->cand_per_frame, h->kp_per_frame, h->tree_cap, code_cap, h->d_skey.p + co, h->d_spk.p + co, h->d_ncand.p + (size_t)b0 * L,
                                                h->d_kp_level.p + (size_t)b0 * h->kp_per_frame, h->d_kp_count.p + (size_t)b0 * L, h->d_overflow.p);
        }
        LAUNCH_CHECK();
    }
    prof_mark(h);
    if (fork_blur) {                                                     // ... but is submitted after the quadtree kernels, so it fills the SMs they leave idle
        CU_TRY(cudaStreamWaitEvent(h->s_alt, h->ev_fork, 0));
        cudaStream_t keep = h->cur; h->cur = h->s_alt;
        const int rc = run_blur_range(h, b0, B);
        h->cur = keep;
        if (rc) return rc;
        CU_TRY(cudaEventRecord(h->ev_join, h->s_alt));
    }
    h->lastB = b0 + B; h->have_pyramid = true; h->blur_valid = false;
    return ORBX_OK;
}

// blur of frames [b0, b0+B).  The single-frame entry points (describe, debug taps) use the cached form below.
static int run_blur_range(orbx_extractor* h, int b0, int B) {
    { const int rc = fast_prepare(h); if (rc) return rc; }
    // a handful of frames: short strips, so that the few tiles there are spread over all SMs (latency form)
    static const int small_env = [] { const char* e = std::getenv("ORBX_BLUR_SMALL"); return e ? std::atoi(e) : -1; }();
    const bool small = small_env >= 0 ? small_env != 0 : (long long)h->tiles.size() * B < 4096;
    const int ntiles = (int)(small ? h->tiles_s.size() : h->tiles.size()), strip = small ? BLUR_STRIP_SMALL : BLUR_STRIP;
    dim3 grid((ntiles + BLUR_WARPS - 1) / BLUR_WARPS, B);
    k_gauss7<<<grid, BLUR_WARPS * 32, BLUR_WARPS * blur_smem_per_warp(strip), h->cur>>>(small ? h->map_l0_blur_s : h->map_l0_blur, h->d_tmaps.p + (small ? 3 : 1) * h->nlevels, b0,
        h->d_levels.p, small ? h->d_tiles_s.p : h->d_tiles.p, ntiles, strip, h->d_blur.p, h->pyr_fstride);
    LAUNCH_CHECK();
    prof_mark(h);
    return ORBX_OK;
}
static int run_blur(orbx_extractor* h, int B) {
    if (h->blur_valid) return ORBX_OK;
    int rc = run_blur_range(h, 0, B); if (rc) return rc;
    h->blur_valid = true;
    return ORBX_OK;
}

// d_kp / d_desc / d_counts / d_level_counts are the buffers of the WHOLE batch; the range offset is applied here
static int run_orient(orbx_extractor* h, int b0, int B, bool describe, KpOut* d_kp, uint8_t* d_desc, int cap, int* d_counts, int* d_level_counts) {
    dim3 grid((h->kp_per_frame + 3) / 4, B);
    PyrView view = h->view;
    view.l0 += (long long)b0 * view.l0_fstride; view.pyr += (long long)b0 * view.pyr_fstride;
    const uint32_t* kp_level = h->d_kp_level.p + (size_t)b0 * h->kp_per_frame;
    const int* kp_count = h->d_kp_count.p + (size_t)b0 * h->nlevels;
    KpOut* kp = d_kp + (size_t)b0 * cap;
    int* counts = d_counts ? d_counts + b0 : nullptr;
    int* lcounts = d_level_counts ? d_level_counts + (size_t)b0 * h->nlevels : nullptr;
    if (describe)
        k_orient_describe<true><<<grid, 128, 0, h->cur>>>(view, h->d_levels.p, h->nlevels, h->kp_per_frame, kp_level, kp_count,
                                                            h->d_blur.p + (size_t)b0 * h->pyr_fstride, h->pyr_fstride, kp, d_desc + (size_t)b0 * cap * 32, cap, counts, lcounts);
    else
        k_orient_describe<false><<<grid, 128, 0, h->cur>>>(view, h->d_levels.p, h->nlevels, h->kp_per_frame, kp_level, kp_count,
                                                             nullptr, 0, kp, nullptr, cap, counts, lcounts);
    LAUNCH_CHECK();
    if (describe) prof_mark(h);
    return ORBX_OK;
}

// level 0 := host frames (H2D straight into the resident pyramid block)
static int upload_level0(orbx_extractor* h, const uint8_t* images, int B, int rows, int cols, size_t step, size_t frame_stride) {
    const LevelGeom& g0 = h->levels[0];
    for (int b = 0; b < B; ++b)
        CU_TRY(cudaMemcpy2DAsync(h->d_pyr.p + (size_t)b * h->pyr_fstride + g0.off, g0.pitch, images + (size_t)b * frame_stride, step,
                                 cols, rows, cudaMemcpyHostToDevice, h->stream));
    h->view.l0 = h->d_pyr.p + g0.off; h->view.l0_fstride = h->pyr_fstride; h->view.l0_pitch = g0.pitch;
    h->view.pyr = h->d_pyr.p; h->view.pyr_fstride = h->pyr_fstride;
    return ORBX_OK;
}

static int check_args(orbx_extractor* h, const void* image, int rows, int cols, size_t step) {
    if (!h) FAIL(ORBX_E_INVALID, "null handle");
    if (!image || rows <= 0 || cols <= 0 || step < (size_t)cols) FAIL(ORBX_E_INVALID, "bad image arguments");
    CU_TRY(cudaSetDevice(h->device));
    return ORBX_OK;
}

// operator()(image, mask, keypoints, descriptors) for ONE frame through a CUDA graph.  Tracking calls this once per frame, and the chain is
// ~20 short dependent kernels: issued one by one the host's launch cost (~4 us each) is longer than most of the kernels, so the call was
// launch-bound.  The graph is captured on first use per (geometry, buffer set) -- every pointer and launch shape in it is a function of
// the plan -- and a call then costs: memcpy of the frame into the pinned staging frame, one cudaGraphLaunch, one synchronisation.
static int extract_graph(orbx_extractor* h, const uint8_t* image, int rows, int cols, size_t step, orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* n_out, int icap) {
    const int pitch = align_up(cols, 16);                  // TMA: 16-byte row step
    const size_t fbytes = (size_t)pitch * rows;
    const size_t kb = (size_t)icap * sizeof(KpOut), db = (size_t)icap * 32, blk = 16 + kb + db;
    if (h->h_in_cap < fbytes) {
        if (h->h_in) cudaFreeHost(h->h_in);
        h->h_in = nullptr; h->h_in_cap = 0;
        CU_TRY(cudaHostAlloc((void**)&h->h_in, fbytes, cudaHostAllocDefault));
        h->h_in_cap = fbytes;
        if (h->graph1) { cudaGraphExecDestroy(h->graph1); h->graph1 = nullptr; }
    }
    if (h->h_gather_cap < blk) {
        if (h->h_gather) cudaFreeHost(h->h_gather);
        h->h_gather = nullptr; h->h_gather_cap = 0;
        CU_TRY(cudaHostAlloc((void**)&h->h_gather, blk, cudaHostAllocDefault));
        h->h_gather_cap = blk;
        if (h->graph1) { cudaGraphExecDestroy(h->graph1); h->graph1 = nullptr; }
    }
    if (h->d_gather.n < blk || h->d_l0.n < fbytes) {
        if (h->d_gather.ensure(blk) || h->d_l0.ensure(fbytes)) return ORBX_E_CUDA;
        if (h->graph1) { cudaGraphExecDestroy(h->graph1); h->graph1 = nullptr; }
    }
    h->view.l0 = h->d_l0.p; h->view.l0_fstride = (long long)fbytes; h->view.l0_pitch = pitch;
    h->view.pyr = h->d_pyr.p; h->view.pyr_fstride = h->pyr_fstride;
    { const int rp = fast_prepare(h); if (rp) return rp; }    // tensor maps are built outside the capture
    // every buffer the captured nodes point at: any of them may have been re-allocated by another entry point since the capture
    const void* sig[28] = {h->d_tp_lv.p, h->d_chain.p, h->d_tmaps.p, h->d_tiles_s.p, h->d_l0.p, h->d_pyr.p, h->d_blur.p, h->d_slots.p, h->d_ocand.p, h->d_spk.p, h->d_skey.p, h->d_cell_counts.p, h->d_ncand.p, h->d_kp_level.p,
                           h->d_kp_count.p, h->d_counts.p, h->d_overflow.p, h->d_kp_out.p, h->d_desc_out.p, h->d_gather.p, h->h_gather, h->h_in, h->d_levels.p, h->d_cells.p,
                           h->d_tiles.p, h->d_tabs.p, (const void*)(uintptr_t)pitch, (const void*)(uintptr_t)blk};
    if (h->graph1 && std::memcmp(sig, h->graph_sig, sizeof(sig)) != 0) { cudaGraphExecDestroy(h->graph1); h->graph1 = nullptr; }
    // stage the frame (the caller's buffer is pageable in general: a DMA straight from it would be a blocking, staged copy anyway)
    if (step == (size_t)pitch) std::memcpy(h->h_in, image, fbytes - (size_t)(pitch - cols));
    else for (int y = 0; y < rows; ++y) std::memcpy(h->h_in + (size_t)y * pitch, image + (size_t)y * step, (size_t)cols);
    int rc = ORBX_OK;
    if (!h->graph1) {
        cudaGraph_t g = nullptr;
        CU_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        const long long l0 = h->launches;
        do {
            if (cudaMemcpyAsync(h->d_l0.p, h->h_in, fbytes, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) { rc = ORBX_E_CUDA; break; }
            h->cur = h->stream;
            if ((rc = run_detect(h, 0, 1, true))) break;
            if (cudaStreamWaitEvent(h->stream, h->ev_join, 0) != cudaSuccess) { rc = ORBX_E_CUDA; break; }
            if ((rc = run_orient(h, 0, 1, true, h->d_kp_out.p, h->d_desc_out.p, icap, h->d_counts.p, nullptr))) break;
            k_gather_result<<<(int)((kb + db + 16 + 4095) / 4096), 256, 0, h->stream>>>(h->d_counts.p, h->d_overflow.p, reinterpret_cast<const uint32_t*>(h->d_kp_out.p), (int)(kb / 4),
                                                                                       reinterpret_cast<const uint32_t*>(h->d_desc_out.p), (int)(db / 4), reinterpret_cast<uint32_t*>(h->d_gather.p));
            if (cudaGetLastError() != cudaSuccess) { rc = ORBX_E_CUDA; break; }
            ++h->launches;
            if (cudaMemcpyAsync(h->h_gather, h->d_gather.p, blk, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) { rc = ORBX_E_CUDA; break; }
        } while (0);
        h->graph_launches = (int)(h->launches - l0);
        const cudaError_t ee = cudaStreamEndCapture(h->stream, &g);
        if (rc || ee != cudaSuccess || !g) { if (g) cudaGraphDestroy(g); if (!rc) orbx_set_error(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ee)); return rc ? rc : ORBX_E_CUDA; }
        const cudaError_t ei = cudaGraphInstantiate(&h->graph1, g, 0);
        cudaGraphDestroy(g);
        if (ei != cudaSuccess) { h->graph1 = nullptr; FAIL(ORBX_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ei)); }
        std::memcpy(h->graph_sig, sig, sizeof(sig));
        h->launches = l0;                                   // counted per replay below
    }
    CU_TRY(cudaGraphLaunch(h->graph1, h->stream));
    h->launches += h->graph_launches;
    CU_TRY(cudaStreamSynchronize(h->stream));
    h->lastB = 1; h->have_pyramid = true; h->blur_valid = true;
    int n, ovf;
    std::memcpy(&n, h->h_gather, 4); std::memcpy(&ovf, h->h_gather + 4, 4);
    if (ovf) { CU_TRY(cudaMemsetAsync(h->d_overflow.p, 0, 16, h->stream)); FAIL(ORBX_E_OVERFLOW, "internal bound exceeded in the quadtree stage"); }
    *n_out = n;
    h->last_kp = h->d_kp_out.p; h->last_desc = h->d_desc_out.p; h->last_n = n;
    if (n > cap) FAIL(ORBX_E_CAPACITY, "keypoint buffer too small");
    std::memcpy(kp_out, h->h_gather + 16, (size_t)n * sizeof(KpOut));
    std::memcpy(desc_out, h->h_gather + 16 + kb, (size_t)n * 32);
    return ORBX_OK;
}

static int upload_constants() {
    // __constant__ tables are per-device module state; upload on every create (cheap, idempotent)
    char4 pt[8 * 32];
    for (int i = 0; i < 32; ++i) for (int k = 0; k < 8; ++k) {
        const signed char* p = k_brief_pattern_host + (size_t)(8 * i + k) * 4;
        pt[k * 32 + i] = make_char4(p[0], p[1], p[2], p[3]);
    }
    CU_TRY(cudaMemcpyToSymbol(g_pattern_t, pt, sizeof(pt)));
    return ORBX_OK;
}

extern "C" {

int orbx_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device, orbx_extractor** out) {
    if (!out) FAIL(ORBX_E_INVALID, "null out");
    *out = nullptr;
    if (nfeatures <= 0 || nlevels <= 0 || nlevels > ORBX_MAX_LEVELS || !(scaleFactor > 1.0f) || iniThFAST < 0 || minThFAST < 0 || minThFAST > 255)
        FAIL(ORBX_E_INVALID, "bad extractor parameters");
    int ndev = 0;
    CU_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) FAIL(ORBX_E_CUDA, "no such CUDA device (this library has no CPU fallback)");
    CU_TRY(cudaSetDevice(device));
    orbx_extractor* h = new orbx_extractor();
    h->nfeatures = nfeatures; h->scaleFactor = scaleFactor; h->nlevels = nlevels; h->iniThFAST = iniThFAST; h->minThFAST = minThFAST; h->device = device;
    // scale tables: double product stored to float (ORBextractor.cc:506-524, include/ORBextractor.h:213)
    h->mvScaleFactor.resize(nlevels); h->mvLevelSigma2.resize(nlevels); h->mvInvScaleFactor.resize(nlevels); h->mvInvLevelSigma2.resize(nlevels);
    h->mvScaleFactor[0] = 1.0f; h->mvLevelSigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) { h->mvScaleFactor[i] = (float)(h->mvScaleFactor[i - 1] * h->scaleFactor); h->mvLevelSigma2[i] = h->mvScaleFactor[i] * h->mvScaleFactor[i]; }
    for (int i = 0; i < nlevels; i++) { h->mvInvScaleFactor[i] = 1.0f / h->mvScaleFactor[i]; h->mvInvLevelSigma2[i] = 1.0f / h->mvLevelSigma2[i]; }
    // per-level quotas (:534-554)
    h->mnFeaturesPerLevel.resize(nlevels);
    float factor = (float)(1.0f / h->scaleFactor);
    float nDesired = (float)(nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels)));
    int sum = 0;
    for (int level = 0; level < nlevels - 1; level++) { h->mnFeaturesPerLevel[level] = cv_round_f(nDesired); sum += h->mnFeaturesPerLevel[level]; nDesired *= factor; }
    h->mnFeaturesPerLevel[nlevels - 1] = std::max(nfeatures - sum, 0);
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; FAIL(ORBX_E_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); }
    h->cur = h->stream;
    cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
    if (cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->s_alt, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->s_more[0], cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&h->s_more[1], cudaStreamNonBlocking) != cudaSuccess) {
        cudaStreamDestroy(h->stream); delete h; FAIL(ORBX_E_CUDA, "cudaStreamCreate (copy streams)");
    }
    if (upload_constants() != ORBX_OK) { cudaStreamDestroy(h->stream); cudaStreamDestroy(h->s_h2d); cudaStreamDestroy(h->s_d2h); delete h; return ORBX_E_CUDA; }
    cudaFuncSetAttribute(k_octree_sort_t<SORT_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)octree_sort_smem_bytes(18432));
    cudaFuncSetAttribute(k_octree_sort_t<SORT_THREADS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)octree_sort_smem_bytes(8192, SORT_THREADS_WIDE));
    cudaFuncSetAttribute(k_octree_tree, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_octree_tree_par, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ptree_smem_bytes(PTREE_MAXCAP, 4096));
    cudaFuncSetAttribute(k_octree_fused<QF_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    cudaFuncSetAttribute(k_octree_fused<QF_THREADS_BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    cudaFuncSetAttribute(k_fast_cells<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(k_fast_cells<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(k_pyr_resize_t, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_pyr_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    *out = h;
    return ORBX_OK;
}

void orbx_destroy(orbx_extractor* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->s_h2d) { cudaStreamSynchronize(h->s_h2d); cudaStreamDestroy(h->s_h2d); }
    if (h->s_d2h) { cudaStreamSynchronize(h->s_d2h); cudaStreamDestroy(h->s_d2h); }
    if (h->s_alt) { cudaStreamSynchronize(h->s_alt); cudaStreamDestroy(h->s_alt); }
    for (cudaStream_t x : h->s_more) if (x) { cudaStreamSynchronize(x); cudaStreamDestroy(x); }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_block) cudaEventDestroy(h->ev_block);
    h->d_gather.release(); if (h->h_gather) cudaFreeHost(h->h_gather);
    if (h->graph1) cudaGraphExecDestroy(h->graph1);
    if (h->h_in) cudaFreeHost(h->h_in);
    if (h->h_bits) cudaFreeHost(h->h_bits);
    for (cudaEvent_t e : h->ev_h2d) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_done) cudaEventDestroy(e);
    h->d_l0.release();
    h->d_tp_lv.release(); h->d_tp_xr.release(); h->d_tp_yr.release(); h->d_chain.release(); h->d_tmaps.release(); h->d_levels.release(); h->d_codetab.release(); h->d_cells.release(); h->d_tiles.release(); h->d_tiles_s.release(); h->d_tabs.release();
    h->d_pyr.release(); h->d_blur.release(); h->d_slots.release(); h->d_ocand.release(); h->d_spk.release(); h->d_kp_level.release();
    h->d_skey.release(); h->d_cell_counts.release(); h->d_ncand.release(); h->d_kp_count.release(); h->d_counts.release();
    h->d_level_counts.release(); h->d_overflow.release(); h->d_kp_out.release(); h->d_desc_out.release();
    h->d_kp_tmp.release(); h->d_desc_tmp.release(); h->d_mask.release(); h->d_bits0.release(); h->d_bits1.release(); h->d_culled.release(); h->d_label.release(); h->d_ids.release(); h->d_label16.release(); h->d_lflags.release();
    cudaStreamDestroy(h->stream);
    delete h;
}

int orbx_get_levels(const orbx_extractor* h) { return h ? h->nlevels : 0; }
float orbx_get_scale_factor(const orbx_extractor* h) { return h ? (float)h->scaleFactor : 0.f; }
#define GETTER(name, member, T) int name(const orbx_extractor* h, T* out) { if (!h || !out) return ORBX_E_INVALID; for (int i = 0; i < h->nlevels; ++i) out[i] = h->member[i]; return ORBX_OK; }
GETTER(orbx_get_scale_factors, mvScaleFactor, float)
GETTER(orbx_get_inverse_scale_factors, mvInvScaleFactor, float)
GETTER(orbx_get_scale_sigma_squares, mvLevelSigma2, float)
GETTER(orbx_get_inverse_scale_sigma_squares, mvInvLevelSigma2, float)
GETTER(orbx_get_features_per_level, mnFeaturesPerLevel, int)
void* orbx_stream(orbx_extractor* h) { return h ? (void*)h->stream : nullptr; }
long long orbx_launch_count(const orbx_extractor* h) { return h ? h->launches : 0; }

// Per-stage device timing for bench.py: while enabled, every batched extract records CUDA events at the
// stage boundaries on the handle's stream; orbx_profile_collect synchronises, sums the elapsed times per
// stage over all profiled calls (ms), and clears the events.  Stage order: resize, fast_cells, octree_sort,
// octree_tree, gauss7, orient_describe.
int orbx_profile_enable(orbx_extractor* h, int on) { if (!h) return ORBX_E_INVALID; h->profiling = on != 0; return ORBX_OK; }
int orbx_profile_collect(orbx_extractor* h, double* stage_ms, int* ncalls) {
    if (!h || !stage_ms || !ncalls) return ORBX_E_INVALID;
    if (cudaSetDevice(h->device) != cudaSuccess) return ORBX_E_CUDA;
    CU_TRY(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < ORBX_NSTAGES; ++i) stage_ms[i] = 0.0;
    const size_t per = ORBX_NSTAGES + 1;
    const size_t calls = h->prof_events.size() / per;
    for (size_t c = 0; c < calls; ++c)
        for (int i = 0; i < ORBX_NSTAGES; ++i) { float ms = 0.f; cudaEventElapsedTime(&ms, h->prof_events[c * per + i], h->prof_events[c * per + i + 1]); stage_ms[i] += ms; }
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    h->prof_events.clear();
    *ncalls = (int)calls;
    return ORBX_OK;
}

int orbx_max_keypoints(orbx_extractor* h, int rows, int cols) {
    if (!h || rows <= 0 || cols <= 0) FAIL(ORBX_E_INVALID, "bad arguments");
    if (cudaSetDevice(h->device) != cudaSuccess) return ORBX_E_CUDA;
    int rc = build_plan(h, rows, cols);
    return rc ? rc : h->max_kp;
}

int orbx_check_overflow(orbx_extractor* h) {
    if (!h || !h->d_overflow.p) return 0;
    int v = 0;
    cudaSetDevice(h->device);
    if (cudaMemcpyAsync(&v, h->d_overflow.p, 4, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) return ORBX_E_CUDA;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return ORBX_E_CUDA;
    if (v && cudaMemsetAsync(h->d_overflow.p, 0, 16, h->stream) != cudaSuccess) return ORBX_E_CUDA;   // reported once, then cleared
    return v;
}

int orbx_extract_batch_device(orbx_extractor* h, const uint8_t* d_images, int B, int rows, int cols, size_t step, size_t frame_stride,
                              orbx_keypoint* d_kp_out, uint8_t* d_desc_out, int cap, int* d_counts_out) {
    int rc = check_args(h, d_images, rows, cols, step); if (rc) return rc;
    if (B <= 0 || !d_kp_out || !d_desc_out || !d_counts_out || cap <= 0) FAIL(ORBX_E_INVALID, "bad batch arguments");
    if ((rc = build_plan(h, rows, cols))) return rc;
    if ((rc = ensure_capacity(h, B, 0))) return rc;
    if (((uintptr_t)d_images & 15) == 0 && (step & 15) == 0 && (frame_stride & 15) == 0) {      // TMA reads level 0: 16-byte alignment
        h->view.l0 = d_images; h->view.l0_fstride = (long long)frame_stride; h->view.l0_pitch = (int)step;       // alias the caller's frames
    } else {
        const LevelGeom& g0 = h->levels[0];
        if ((rc = repitch_frames(h, h->d_pyr.p + g0.off, h->pyr_fstride, g0.pitch, d_images, (long long)frame_stride, (long long)step, cols, rows, B, h->stream))) return rc;
        h->view.l0 = h->d_pyr.p + g0.off; h->view.l0_fstride = h->pyr_fstride; h->view.l0_pitch = g0.pitch;
    }
    h->view.pyr = h->d_pyr.p; h->view.pyr_fstride = h->pyr_fstride;
    if (h->profiling) {                                  // per-stage timing wants the stages back to back on one stream
        if ((rc = run_detect(h, 0, B))) return rc;
        if ((rc = run_blur(h, B))) return rc;
        h->lb_kp = reinterpret_cast<KpOut*>(d_kp_out); h->lb_desc = d_desc_out; h->lb_counts = d_counts_out; h->lb_cap = cap; h->lb_B = B;
        return run_orient(h, 0, B, true, reinterpret_cast<KpOut*>(d_kp_out), d_desc_out, cap, d_counts_out, nullptr);
    }
    // one pass; the blur is forked next to FAST / quadtree.  (Chunking the batch over the two compute streams, as the host
    // path does, was measured and gains nothing here: every kernel already fills the GPU.)
    if ((rc = run_detect(h, 0, B, true))) return rc;
    CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    h->blur_valid = true;
    h->lb_kp = reinterpret_cast<KpOut*>(d_kp_out); h->lb_desc = d_desc_out; h->lb_counts = d_counts_out; h->lb_cap = cap; h->lb_B = B;
    return run_orient(h, 0, B, true, reinterpret_cast<KpOut*>(d_kp_out), d_desc_out, cap, d_counts_out, nullptr);
}

// Host-pointer batch call = the reference-facing end-to-end path.  The batch is cut into chunks that flow through
// three streams (H2D copy engine -> compute -> D2H copy engine), so the PCIe transfers of chunk i+1 / i-1 overlap the
// kernels of chunk i.  Pinned host buffers give true overlap; pageable ones still work (the copies then serialise).
static int chunk_frames(const orbx_extractor* h, int B) {
    const long long px = (long long)h->rows * h->cols;
    static const int tune = [] { const char* e = std::getenv("ORBX_CHUNK_FRAMES"); return e ? std::atoi(e) : 0; }();   // tuning knob (VGA-equivalent frames per chunk)
    long long c = ((tune > 0 ? tune : 64) * 640LL * 480 + px - 1) / px;          // default ~64 VGA frames (20 MB) per chunk
    if (c < 1) c = 1;
    if (c > B) c = B;
    return (int)c;
}

static int run_masked_range(orbx_extractor* h, int b0, int nb, const uint8_t* d_masks, long long mfs, int mpitch, int rows, int cols, LabelView lv,
                            KpOut* d_kp, uint8_t* d_desc, int cap, int* d_counts, int* d_culled, bool prepacked = false, bool fork = false);
extern "C" void orbx_host_pack_mask(const uint8_t* mask, size_t step, int rows, int cols, uint32_t* bits);   // host_pack.cpp
static int ensure_closing(orbx_extractor* h, int nframes, int rows, int cols);

// masks == nullptr: operator()(image, mask, keypoints, descriptors) per frame; otherwise the two-stage Amos path with culling
static int host_batch_pipeline(orbx_extractor* h, const uint8_t* images, const uint8_t* masks, const orbx_labels* labels, int B, int rows, int cols, size_t step, size_t frame_stride,
                               size_t mask_step, size_t mask_frame_stride, orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* counts_out, int* culled_out) {
    int rc = check_args(h, images, rows, cols, step); if (rc) return rc;
    if (B <= 0 || !kp_out || !desc_out || !counts_out || cap <= 0) FAIL(ORBX_E_INVALID, "bad batch arguments");
    if ((rc = build_plan(h, rows, cols))) return rc;
    if ((rc = ensure_capacity(h, B, cap))) return rc;
    // level 0: mirror the host layout on the device when it is word-aligned and dense enough (one copy per chunk);
    // otherwise copy frame by frame into the pyramid block
    const bool dense = frame_stride >= step * (size_t)rows && frame_stride <= 2 * step * (size_t)rows;
    const bool mirror = dense && (step & 15) == 0 && (frame_stride & 15) == 0;
    // rows that are not word-aligned (KITTI: 1241 px): still ONE dense copy per chunk over the bus -- a strided 2-D copy of 1241-byte rows
    // runs at a fraction of the link rate -- followed by a device-side re-pitch into the pyramid block
    const bool repitch = dense && !mirror;
    if (mirror || repitch) {
        if (h->d_l0.ensure((size_t)B * frame_stride)) return ORBX_E_CUDA;
    }
    if (mirror) {
        h->view.l0 = h->d_l0.p; h->view.l0_fstride = (long long)frame_stride; h->view.l0_pitch = (int)step;
    } else {
        const LevelGeom& g0 = h->levels[0];
        h->view.l0 = h->d_pyr.p + g0.off; h->view.l0_fstride = h->pyr_fstride; h->view.l0_pitch = g0.pitch;
    }
    h->view.pyr = h->d_pyr.p; h->view.pyr_fstride = h->pyr_fstride;
    const int mpitch = align_up(cols, 128);
    const size_t mfs = (size_t)mpitch * rows;
    // The masks only matter as "mask != 0" (src/ORBextractor.cc:1697-1724), and this call is bound by the host -> device link: pack them to
    // 1 bit per pixel on the host (a few worker threads, frame by frame, ahead of the copies) and upload 1/8 of the bytes.  ORBX_HOST_PACK=0
    // uploads the byte masks and packs on the device instead.
    static const int pack_threads = [] { const char* e = std::getenv("ORBX_HOST_PACK"); int v = e ? std::atoi(e) : 6; return v < 0 ? 0 : (v > 16 ? 16 : v); }();
    const bool hostpack = masks && pack_threads > 0;
    const size_t bwords = (size_t)rows * ((cols + 31) / 32);
    struct Packer {                                               // joins its threads on every exit path
        std::vector<std::thread> th; std::atomic<int> next{0}; std::unique_ptr<std::atomic<unsigned char>[]> done; int B = 0;
        ~Packer() { next.store(1 << 30); for (auto& t : th) if (t.joinable()) t.join(); }
        void wait(int b0, int nb) { for (int b = b0; b < b0 + nb; ++b) while (!done[b].load(std::memory_order_acquire)) std::this_thread::yield(); }
    } packer;
    if (masks) {
        if ((rc = ensure_closing(h, B, rows, cols))) return rc;
        if ((!hostpack && h->d_mask.ensure(mfs * B + 64)) || h->d_culled.ensure(B)) return ORBX_E_CUDA;
    }
    if (hostpack) {
        if (h->h_bits_cap < bwords * B) {
            if (h->h_bits) cudaFreeHost(h->h_bits);
            h->h_bits = nullptr; h->h_bits_cap = 0;
            CU_TRY(cudaHostAlloc((void**)&h->h_bits, bwords * B * 4, cudaHostAllocDefault));
            h->h_bits_cap = bwords * B;
        }
        packer.B = B; packer.done.reset(new std::atomic<unsigned char>[B]);
        for (int b = 0; b < B; ++b) packer.done[b].store(0);
        uint32_t* hb = h->h_bits;
        for (int t = 0; t < std::min(pack_threads, B); ++t)
            packer.th.emplace_back([&packer, hb, masks, mask_step, mask_frame_stride, rows, cols, bwords] {
                for (;;) {
                    const int b = packer.next.fetch_add(1);
                    if (b >= packer.B) break;
                    orbx_host_pack_mask(masks + (size_t)b * mask_frame_stride, mask_step, rows, cols, hb + (size_t)b * bwords);
                    packer.done[b].store(1, std::memory_order_release);
                }
            });
    }
    // super-pixel labels ride along with the masks: 16-bit ids in a dense device mirror (cols elements per row) + the per-frame flag tables
    const size_t lfs = (size_t)rows * cols;
    if (labels && (h->d_label16.ensure(lfs * B) || h->d_lflags.ensure((size_t)labels->n_labels * B))) return ORBX_E_CUDA;
    // chunk boundaries: short first chunks (C/4, C/2, then C) let compute start while most of the batch is still on the bus
    const int C = chunk_frames(h, B);
    static const int ramp = [] { const char* e = std::getenv("ORBX_CHUNK_RAMP"); return e ? std::atoi(e) : 1; }();
    std::vector<int> cb(1, 0);
    for (int sz = ramp ? std::max(1, C / 4) : C; cb.back() < B; sz = std::min(C, sz * 2)) cb.push_back(std::min(B, cb.back() + sz));
    const int nchunks = (int)cb.size() - 1;
    while ((int)h->ev_h2d.size() < nchunks) {
        cudaEvent_t a, b;
        CU_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming)); CU_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        h->ev_h2d.push_back(a); h->ev_done.push_back(b);
    }
    // the copy streams must not run ahead of work already queued on the compute stream (e.g. a previous call's kernels)
    CU_TRY(cudaEventRecord(h->ev_done[0], h->stream));
    CU_TRY(cudaStreamWaitEvent(h->s_h2d, h->ev_done[0], 0));
    static const int nstreams = [] { const char* e = std::getenv("ORBX_HOST_STREAMS"); int v = e ? std::atoi(e) : 4; return v < 1 ? 1 : (v > 4 ? 4 : v); }();
    cudaStream_t cs[4] = {h->stream, h->s_alt, h->s_more[0], h->s_more[1]};
    for (int i = 1; i < nstreams; ++i) CU_TRY(cudaStreamWaitEvent(cs[i], h->ev_done[0], 0));
    auto issue_h2d = [&](int c) -> int {
        const int b0 = cb[c], nb = cb[c + 1] - b0;
        if (mirror || repitch) {
            const size_t bytes = (size_t)(nb - 1) * frame_stride + (size_t)(rows - 1) * step + cols;     // never reads past the last row of the last frame
            CU_TRY(cudaMemcpyAsync(h->d_l0.p + (size_t)b0 * frame_stride, images + (size_t)b0 * frame_stride, bytes, cudaMemcpyHostToDevice, h->s_h2d));
            if (repitch) {
                const LevelGeom& g0 = h->levels[0];
                const int rr = repitch_frames(h, h->d_pyr.p + (size_t)b0 * h->pyr_fstride + g0.off, h->pyr_fstride, g0.pitch, h->d_l0.p + (size_t)b0 * frame_stride, (long long)frame_stride, (long long)step,
                                              cols, rows, nb, h->s_h2d);
                if (rr) return rr;
            }
        } else {
            const LevelGeom& g0 = h->levels[0];
            for (int b = b0; b < b0 + nb; ++b)
                CU_TRY(cudaMemcpy2DAsync(h->d_pyr.p + (size_t)b * h->pyr_fstride + g0.off, g0.pitch, images + (size_t)b * frame_stride, step,
                                         cols, rows, cudaMemcpyHostToDevice, h->s_h2d));
        }
        if (hostpack) {                                                                                    // 1 bit per pixel, packed by the worker threads above
            packer.wait(b0, nb);
            CU_TRY(cudaMemcpyAsync(h->d_bits0.p + (size_t)b0 * bwords, h->h_bits + (size_t)b0 * bwords, (size_t)nb * bwords * 4, cudaMemcpyHostToDevice, h->s_h2d));
        } else if (masks) {
            if (mask_step == (size_t)mpitch && mask_frame_stride == mfs)                                  // dense and already pitched: one copy per chunk
                CU_TRY(cudaMemcpyAsync(h->d_mask.p + (size_t)b0 * mfs, masks + (size_t)b0 * mask_frame_stride, (size_t)nb * mfs, cudaMemcpyHostToDevice, h->s_h2d));
            else
                for (int b = b0; b < b0 + nb; ++b)
                    CU_TRY(cudaMemcpy2DAsync(h->d_mask.p + (size_t)b * mfs, mpitch, masks + (size_t)b * mask_frame_stride, mask_step, cols, rows, cudaMemcpyHostToDevice, h->s_h2d));
        }
        if (labels) {
            if (labels->label_step == (size_t)cols && labels->label_frame_stride == lfs)
                CU_TRY(cudaMemcpyAsync(h->d_label16.p + (size_t)b0 * lfs, labels->labels + (size_t)b0 * lfs, (size_t)nb * lfs * 2, cudaMemcpyHostToDevice, h->s_h2d));
            else
                for (int b = b0; b < b0 + nb; ++b)
                    CU_TRY(cudaMemcpy2DAsync(h->d_label16.p + (size_t)b * lfs, (size_t)cols * 2, labels->labels + (size_t)b * labels->label_frame_stride, labels->label_step * 2,
                                             (size_t)cols * 2, rows, cudaMemcpyHostToDevice, h->s_h2d));
            CU_TRY(cudaMemcpyAsync(h->d_lflags.p + (size_t)b0 * labels->n_labels, labels->flagged + (size_t)b0 * labels->n_labels, (size_t)nb * labels->n_labels, cudaMemcpyHostToDevice, h->s_h2d));
        }
        CU_TRY(cudaEventRecord(h->ev_h2d[c], h->s_h2d));
        return ORBX_OK;
    };
    auto issue_compute = [&](int c) -> int {
        const int b0 = cb[c], nb = cb[c + 1] - b0;
        h->cur = cs[c % nstreams];
        CU_TRY(cudaStreamWaitEvent(h->cur, h->ev_h2d[c], 0));
        if (masks) {
            LabelView lv{nullptr, 0, 0, nullptr, 0};
            if (labels) lv = LabelView{h->d_label16.p + (size_t)b0 * lfs, (long long)lfs, cols, h->d_lflags.p + (size_t)b0 * labels->n_labels, labels->n_labels};
            rc = run_masked_range(h, b0, nb, hostpack ? nullptr : h->d_mask.p + (size_t)b0 * mfs, (long long)mfs, mpitch, rows, cols, lv, h->d_kp_out.p, h->d_desc_out.p, cap, h->d_counts.p, h->d_culled.p, hostpack);
        }
        else {
            rc = run_detect(h, b0, nb);
            if (!rc) rc = run_blur_range(h, b0, nb);
            if (!rc) rc = run_orient(h, b0, nb, true, h->d_kp_out.p, h->d_desc_out.p, cap, h->d_counts.p, nullptr);
        }
        cudaStream_t used = h->cur;
        h->cur = h->stream;
        if (rc) return rc;
        CU_TRY(cudaEventRecord(h->ev_done[c], used));
        CU_TRY(cudaStreamWaitEvent(h->s_d2h, h->ev_done[c], 0));
        CU_TRY(cudaMemcpyAsync(kp_out + (size_t)b0 * cap, h->d_kp_out.p + (size_t)b0 * cap, (size_t)nb * cap * sizeof(KpOut), cudaMemcpyDeviceToHost, h->s_d2h));
        CU_TRY(cudaMemcpyAsync(desc_out + (size_t)b0 * cap * 32, h->d_desc_out.p + (size_t)b0 * cap * 32, (size_t)nb * cap * 32, cudaMemcpyDeviceToHost, h->s_d2h));
        CU_TRY(cudaMemcpyAsync(counts_out + b0, h->d_counts.p + b0, (size_t)nb * sizeof(int), cudaMemcpyDeviceToHost, h->s_d2h));
        if (masks && culled_out) CU_TRY(cudaMemcpyAsync(culled_out + b0, h->d_culled.p + b0, (size_t)nb * sizeof(int), cudaMemcpyDeviceToHost, h->s_d2h));
        return ORBX_OK;
    };
    // Uploads run ahead of the kernels: all of them are queued first -- except when the masks are packed on the host, where an upload has to
    // wait for its frames to be packed and the kernels of the chunk before it are queued meanwhile.
    if (!hostpack) {
        for (int c = 0; c < nchunks; ++c) if ((rc = issue_h2d(c))) return rc;
        for (int c = 0; c < nchunks; ++c) if ((rc = issue_compute(c))) return rc;
    } else {
        if ((rc = issue_h2d(0))) return rc;
        for (int c = 0; c < nchunks; ++c) {
            if (c + 1 < nchunks && (rc = issue_h2d(c + 1))) return rc;
            if ((rc = issue_compute(c))) return rc;
        }
    }
    h->lastB = B; h->blur_valid = true;
    int ovf = 0;
    // The calling thread has nothing to do for milliseconds: it can sleep on a blocking-sync event instead of spinning in
    // cudaStreamSynchronize (opt-in, ORBX_BLOCKING_SYNC=1: for hosts with fewer cores than feeding threads; measured 5 % slower at N = 1 and
    // neutral at N = 8 on a 32-vCPU box, so spinning stays the default).
    static const bool blocking = [] { const char* e = std::getenv("ORBX_BLOCKING_SYNC"); return e && std::atoi(e) != 0; }();
    if (blocking && B >= 32) {
        if (!h->ev_block) CU_TRY(cudaEventCreateWithFlags(&h->ev_block, cudaEventBlockingSync | cudaEventDisableTiming));
        CU_TRY(cudaEventRecord(h->ev_block, h->s_d2h));
        CU_TRY(cudaEventSynchronize(h->ev_block));
    }
    CU_TRY(cudaStreamSynchronize(h->s_d2h));            // follows every chunk's kernels (ev_done) and copies
    for (int i = 1; i < nstreams; ++i) CU_TRY(cudaStreamSynchronize(cs[i]));
    CU_TRY(cudaMemcpyAsync(&ovf, h->d_overflow.p, 4, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    if (ovf) { CU_TRY(cudaMemsetAsync(h->d_overflow.p, 0, 16, h->stream)); FAIL(ORBX_E_OVERFLOW, "internal bound exceeded in the quadtree stage"); }
    h->lb_kp = h->d_kp_out.p; h->lb_desc = h->d_desc_out.p; h->lb_counts = h->d_counts.p; h->lb_cap = cap; h->lb_B = B;
    for (int b = 0; b < B; ++b) if (counts_out[b] > cap) FAIL(ORBX_E_CAPACITY, "keypoint buffer too small: a frame holds more than `cap` keypoints (counts_out has the true counts)");
    return ORBX_OK;
}

int orbx_extract_batch(orbx_extractor* h, const uint8_t* images, int B, int rows, int cols, size_t step, size_t frame_stride,
                       orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* counts_out) {
    return host_batch_pipeline(h, images, nullptr, nullptr, B, rows, cols, step, frame_stride, 0, 0, kp_out, desc_out, cap, counts_out, nullptr);
}

int orbx_extract(orbx_extractor* h, const uint8_t* image, int rows, int cols, size_t step,
                 orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* n_out) {
    if (n_out) *n_out = 0;
    if (!h) FAIL(ORBX_E_INVALID, "null handle");
    if (!image || rows <= 0 || cols <= 0) return ORBX_OK;          // empty image: silent return, 0 keypoints (ORBextractor.cc:1553)
    if (!kp_out || !desc_out || !n_out) FAIL(ORBX_E_INVALID, "null output");
    int rc = check_args(h, image, rows, cols, step); if (rc) return rc;
    if ((rc = build_plan(h, rows, cols))) return rc;
    const int icap = h->max_kp;
    if ((rc = ensure_capacity(h, 1, icap))) return rc;
    static const bool use_graph = [] { const char* e = std::getenv("ORBX_GRAPH"); return !e || std::atoi(e) != 0; }();
    if (use_graph && !h->profiling) return extract_graph(h, image, rows, cols, step, kp_out, desc_out, cap, n_out, icap);
    if ((rc = upload_level0(h, image, 1, rows, cols, step, 0))) return rc;
    if ((rc = run_detect(h, 0, 1, true))) return rc;                 // blur forked next to FAST / quadtree
    CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    h->blur_valid = true;
    if ((rc = run_orient(h, 0, 1, true, h->d_kp_out.p, h->d_desc_out.p, icap, h->d_counts.p, nullptr))) return rc;
    // Everything the call returns (count, overflow flag, all icap keypoint / descriptor slots) is gathered into one device block and
    // comes back in ONE copy into a pinned landing buffer: copies straight into the caller's pageable arrays would each be a blocking,
    // staged transfer of its own.  The count is not known before the copy, so all slots travel (61 KB at C1) and n of them are handed out.
    const size_t kb = (size_t)icap * sizeof(KpOut), db = (size_t)icap * 32, blk = 16 + kb + db;
    if (h->d_gather.ensure(blk)) return ORBX_E_CUDA;
    if (h->h_gather_cap < blk) {
        if (h->h_gather) cudaFreeHost(h->h_gather);
        h->h_gather = nullptr; h->h_gather_cap = 0;
        CU_TRY(cudaHostAlloc((void**)&h->h_gather, blk, cudaHostAllocDefault));
        h->h_gather_cap = blk;
    }
    CU_TRY(cudaMemcpyAsync(h->d_gather.p, h->d_counts.p, 4, cudaMemcpyDeviceToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(h->d_gather.p + 4, h->d_overflow.p, 4, cudaMemcpyDeviceToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(h->d_gather.p + 16, h->d_kp_out.p, kb, cudaMemcpyDeviceToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(h->d_gather.p + 16 + kb, h->d_desc_out.p, db, cudaMemcpyDeviceToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(h->h_gather, h->d_gather.p, blk, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    int n, ovf;
    std::memcpy(&n, h->h_gather, 4); std::memcpy(&ovf, h->h_gather + 4, 4);
    if (ovf) { CU_TRY(cudaMemsetAsync(h->d_overflow.p, 0, 16, h->stream)); FAIL(ORBX_E_OVERFLOW, "internal bound exceeded in the quadtree stage"); }
    *n_out = n;
    h->last_kp = h->d_kp_out.p; h->last_desc = h->d_desc_out.p; h->last_n = n;
    if (n > cap) FAIL(ORBX_E_CAPACITY, "keypoint buffer too small");
    std::memcpy(kp_out, h->h_gather + 16, (size_t)n * sizeof(KpOut));
    std::memcpy(desc_out, h->h_gather + 16 + kb, (size_t)n * 32);
    return ORBX_OK;
}

int orbx_detect(orbx_extractor* h, const uint8_t* image, int rows, int cols, size_t step,
                orbx_keypoint* kp_out, int* level_counts, int cap, int* n_out) {
    if (n_out) *n_out = 0;
    if (!h) FAIL(ORBX_E_INVALID, "null handle");
    if (!image || rows <= 0 || cols <= 0) return ORBX_OK;
    if (!kp_out || !level_counts || !n_out) FAIL(ORBX_E_INVALID, "null output");
    int rc = check_args(h, image, rows, cols, step); if (rc) return rc;
    if ((rc = build_plan(h, rows, cols))) return rc;
    const int icap = h->max_kp;
    if ((rc = ensure_capacity(h, 1, icap))) return rc;
    if ((rc = upload_level0(h, image, 1, rows, cols, step, 0))) return rc;
    if ((rc = run_detect(h, 0, 1))) return rc;
    if ((rc = run_orient(h, 0, 1, false, h->d_kp_out.p, nullptr, icap, h->d_counts.p, h->d_level_counts.p))) return rc;
    // one gathered block, one copy into the pinned landing buffer (see orbx_extract): [n | overflow | level counts | all icap keypoint slots]
    const size_t lc = (size_t)4 * h->nlevels, o_kp = (8 + lc + 15) & ~(size_t)15, kb = (size_t)icap * sizeof(KpOut), blk = o_kp + kb;
    if (h->d_gather.ensure(blk)) return ORBX_E_CUDA;
    if (h->h_gather_cap < blk) {
        if (h->h_gather) cudaFreeHost(h->h_gather);
        h->h_gather = nullptr; h->h_gather_cap = 0;
        CU_TRY(cudaHostAlloc((void**)&h->h_gather, blk, cudaHostAllocDefault));
        h->h_gather_cap = blk;
    }
    CU_TRY(cudaMemcpyAsync(h->d_gather.p, h->d_counts.p, 4, cudaMemcpyDeviceToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(h->d_gather.p + 4, h->d_overflow.p, 4, cudaMemcpyDeviceToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(h->d_gather.p + 8, h->d_level_counts.p, lc, cudaMemcpyDeviceToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(h->d_gather.p + o_kp, h->d_kp_out.p, kb, cudaMemcpyDeviceToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(h->h_gather, h->d_gather.p, blk, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    int n, ovf;
    std::memcpy(&n, h->h_gather, 4); std::memcpy(&ovf, h->h_gather + 4, 4);
    if (ovf) { CU_TRY(cudaMemsetAsync(h->d_overflow.p, 0, 16, h->stream)); FAIL(ORBX_E_OVERFLOW, "internal bound exceeded in the quadtree stage"); }
    std::memcpy(level_counts, h->h_gather + 8, lc);
    *n_out = n;
    if (n > cap) FAIL(ORBX_E_CAPACITY, "keypoint buffer too small");
    std::memcpy(kp_out, h->h_gather + o_kp, (size_t)n * sizeof(KpOut));
    return ORBX_OK;
}

int orbx_describe(orbx_extractor* h, const orbx_keypoint* kp_in, const int* level_counts, orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* n_out) {
    if (!h || !level_counts || !n_out) FAIL(ORBX_E_INVALID, "null argument");
    if (!h->have_pyramid) FAIL(ORBX_E_STATE, "orbx_describe needs the pyramid of a previous orbx_detect / orbx_extract");
    CU_TRY(cudaSetDevice(h->device));
    int n = 0;
    for (int l = 0; l < h->nlevels; ++l) { if (level_counts[l] < 0) FAIL(ORBX_E_INVALID, "negative level count"); n += level_counts[l]; }
    *n_out = n;
    if (n == 0) { h->last_kp = nullptr; h->last_desc = nullptr; h->last_n = 0; return ORBX_OK; }
    if (!kp_in || !kp_out || !desc_out) FAIL(ORBX_E_INVALID, "null buffer");
    if (n > cap) FAIL(ORBX_E_CAPACITY, "keypoint buffer too small");
    // the reference takes the level from the position in allKeypoints, not from kp.octave: stamp it
    std::vector<KpOut> tmp(n);
    std::memcpy(tmp.data(), kp_in, sizeof(KpOut) * n);
    { int o = 0; for (int l = 0; l < h->nlevels; ++l) for (int i = 0; i < level_counts[l]; ++i, ++o) {
        const LevelGeom& g = h->levels[l];
        const int x = (int)lrintf(tmp[o].x), y = (int)lrintf(tmp[o].y);
        if (x < ORBX_EDGE || y < ORBX_EDGE || x >= g.w - ORBX_EDGE || y >= g.h - ORBX_EDGE) FAIL(ORBX_E_INVALID, "keypoint closer than 19 px to the level border");
        if (tmp[o].octave != l) FAIL(ORBX_E_INVALID, "keypoint octave does not match its level bucket");
    } }
    if (h->d_kp_tmp.ensure((size_t)n * 2) || h->d_desc_tmp.ensure((size_t)n * 32)) return ORBX_E_CUDA;
    int rc;
    if ((rc = run_blur(h, 1))) return rc;
    CU_TRY(cudaMemcpyAsync(h->d_kp_tmp.p, tmp.data(), sizeof(KpOut) * n, cudaMemcpyHostToDevice, h->stream));
    k_describe_given<<<(n + 3) / 4, 128, 0, h->stream>>>(h->d_levels.p, h->nlevels, h->d_kp_tmp.p, n, h->d_blur.p, h->d_kp_tmp.p + n, h->d_desc_tmp.p);
    LAUNCH_CHECK();
    CU_TRY(cudaMemcpyAsync(kp_out, h->d_kp_tmp.p + n, sizeof(KpOut) * n, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaMemcpyAsync(desc_out, h->d_desc_tmp.p, (size_t)n * 32, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    h->last_kp = h->d_kp_tmp.p + n; h->last_desc = h->d_desc_tmp.p; h->last_n = n;
    return ORBX_OK;
}

static int copy_level(orbx_extractor* h, const uint8_t* base, int pitch, int level, int border, uint8_t* dst, size_t dst_step) {
    const LevelGeom& g = h->levels[level];
    if (border == 0) {
        CU_TRY(cudaMemcpy2DAsync(dst, dst_step, base, pitch, g.w, g.h, cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        return ORBX_OK;
    }
    // padded export: BORDER_REFLECT_101 frame of `border` px, materialised by a small kernel
    const int W = g.w + 2 * border, H = g.h + 2 * border;
    if (h->d_mask.ensure((size_t)W * H)) return ORBX_E_CUDA;
    k_border101<<<dim3((W + 31) / 32, (H + 7) / 8), dim3(32, 8), 0, h->stream>>>(base, pitch, g.w, g.h, border, h->d_mask.p, W, H);
    LAUNCH_CHECK();
    CU_TRY(cudaMemcpy2DAsync(dst, dst_step, h->d_mask.p, W, W, H, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

int orbx_pyramid_level(orbx_extractor* h, int level, int border, uint8_t* dst, size_t dst_step, int* rows_out, int* cols_out) {
    if (!h) FAIL(ORBX_E_INVALID, "null handle");
    if (!h->have_pyramid) FAIL(ORBX_E_STATE, "no resident pyramid");
    if (level < 0 || level >= h->nlevels || border < 0 || border > 64) FAIL(ORBX_E_INVALID, "bad level / border");
    CU_TRY(cudaSetDevice(h->device));
    const LevelGeom& g = h->levels[level];
    if (rows_out) *rows_out = g.h + 2 * border;
    if (cols_out) *cols_out = g.w + 2 * border;
    if (!dst) return ORBX_OK;
    if (dst_step < (size_t)(g.w + 2 * border)) FAIL(ORBX_E_INVALID, "dst_step too small");
    const uint8_t* base; int pitch;
    if (level == 0) { base = h->view.l0; pitch = h->view.l0_pitch; } else { base = h->d_pyr.p + g.off; pitch = g.pitch; }
    return copy_level(h, base, pitch, level, border, dst, dst_step);
}

}  // extern "C"

#include "orbx_extractor_amos.inl"
#include "orbx_extractor_taps.inl"

int orbx_internal_pyramid(orbx_extractor* h, OrbxPyramidInfo* out) {
    if (!h || !out) FAIL(ORBX_E_INVALID, "null handle");
    if (!h->have_pyramid) FAIL(ORBX_E_STATE, "extractor holds no pyramid (call extract / detect first)");
    out->device = h->device; out->nlevels = h->nlevels; out->stream = h->stream;
    for (int l = 0; l < h->nlevels; ++l) {
        const LevelGeom& g = h->levels[l];
        if (l == 0) { out->ptr[0] = h->view.l0; out->pitch[0] = h->view.l0_pitch; }
        else { out->ptr[l] = h->d_pyr.p + g.off; out->pitch[l] = g.pitch; }
        out->w[l] = g.w; out->h[l] = g.h;
    }
    out->scale = h->mvScaleFactor.data(); out->inv_scale = h->mvInvScaleFactor.data();
    return ORBX_OK;
}


int orbx_internal_last_result(orbx_extractor* h, OrbxLastResult* out) {
    if (!h || !out) FAIL(ORBX_E_INVALID, "null handle");
    if (h->last_n < 0) FAIL(ORBX_E_STATE, "extractor holds no result (call orbx_extract or orbx_describe first)");
    out->device = h->device; out->n = h->last_n; out->nlevels = h->nlevels; out->stream = h->stream;
    out->keys = h->last_kp; out->desc = h->last_desc; out->scale = h->mvScaleFactor.data();
    return ORBX_OK;
}

int orbx_internal_last_batch(orbx_extractor* h, OrbxBatchInfo* out) {
    if (!h || !out) FAIL(ORBX_E_INVALID, "null handle");
    if (!h->have_pyramid || h->lb_B <= 0) FAIL(ORBX_E_STATE, "extractor holds no batched result (call orbx_extract_batch[_device] first)");
    out->device = h->device; out->nlevels = h->nlevels; out->B = h->lb_B; out->cap = h->lb_cap; out->stream = h->stream;
    for (int l = 0; l < h->nlevels; ++l) {
        const LevelGeom& g = h->levels[l];
        if (l == 0) { out->ptr[0] = h->view.l0; out->pitch[0] = h->view.l0_pitch; out->fstride[0] = h->view.l0_fstride; }
        else { out->ptr[l] = h->d_pyr.p + g.off; out->pitch[l] = g.pitch; out->fstride[l] = h->pyr_fstride; }
        out->w[l] = g.w; out->h[l] = g.h;
    }
    out->keys = h->lb_kp; out->desc = h->lb_desc; out->counts = h->lb_counts;
    out->scale = h->mvScaleFactor.data(); out->inv_scale = h->mvInvScaleFactor.data();
    return ORBX_OK;
}

#ifdef ORBX_QT_STAMPS
// probe build only (tools/qt_stamps_probe.py): clock64 stamps of the quadtree kernels' phases, (level 0, frame 0) instance
extern "C" int orbx_probe_qt_stamps(long long* out, int n) {
    if (!out || n <= 0 || n > 64) FAIL(ORBX_E_INVALID, "bad arguments");
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpyFromSymbol(out, g_qt_stamps, sizeof(long long) * (size_t)n));
    return ORBX_OK;
}
#endif
