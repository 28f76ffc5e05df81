// orbx_matcher.cu -- host side of the B200-native ORBmatcher / ComputeStereoMatches (C ABI of include/orbx_b200.h).
//
// Mirrors ORB_SLAM2::ORBmatcher (/root/reference/include/ORBmatcher.h:57-215, src/ORBmatcher.cc) and
// Frame::ComputeStereoMatches (/root/reference/src/Frame.cc:1179-1573).  The host only marshals the caller's
// arrays to the device and launches the kernels of k_match.cuh; there is no CPU matching path.
#include "../../include/orbx_b200.h"
#include "orbx_internal.h"
#include "k_match.cuh"
#include "k_frame.cuh"
#include "k_bow.cuh"

#include <climits>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <thread>

void orbx_set_error(const std::string& s);
#define CU_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { \
    orbx_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); return ORBX_E_CUDA; } } while (0)
#define FAIL(code, msg) do { orbx_set_error(msg); return (code); } while (0)
#define LAUNCH_CHECK() do { cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) { \
    orbx_set_error(std::string("kernel launch: ") + cudaGetErrorString(_e)); return ORBX_E_CUDA; } ++m->launches; } while (0)

static_assert(sizeof(KpM) == 28, "KpM must match cv::KeyPoint");

namespace {
// bump allocator over one device arena, reset at the start of every call
struct Arena {
    uint8_t* base = nullptr; size_t cap = 0, off = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return ORBX_OK;
        if (base) cudaFree(base);
        base = nullptr; cap = 0;
        size_t want = bytes + (bytes >> 2) + (1 << 20);
        if (cudaMalloc((void**)&base, want) != cudaSuccess) { orbx_set_error("cudaMalloc (matcher arena)"); return ORBX_E_CUDA; }
        cap = want; return ORBX_OK;
    }
    void reset() { off = 0; }
    template <typename T> T* get(size_t count) {
        size_t o = (off + 255) & ~(size_t)255;
        off = o + count * sizeof(T);
        return off <= cap ? reinterpret_cast<T*>(base + o) : nullptr;
    }
};
inline size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }
}

// host -> device uploads of one call are packed into a pinned mirror of a device arena and sent with ONE cudaMemcpyAsync
struct UploadArena {
    uint8_t* dbase = nullptr; uint8_t* hbase = nullptr; size_t cap = 0, off = 0;
    // batched calls stage hundreds of frames per call: their copies into the pinned mirror are deferred and done by a few threads at
    // flush time (a single thread's memcpy, ~10 GB/s, was the limit of the end-to-end frame-pair figure: 65 k pairs/s against 241 k on the device)
    struct Job { size_t dst; const void* src; size_t bytes; };
    std::vector<Job> jobs; bool defer = false; size_t deferred = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return ORBX_OK;
        if (dbase) cudaFree(dbase);
        if (hbase) cudaFreeHost(hbase);
        dbase = hbase = nullptr; cap = 0;
        size_t want = bytes + (bytes >> 2) + (1 << 20);
        if (cudaMalloc((void**)&dbase, want) != cudaSuccess || cudaHostAlloc((void**)&hbase, want, cudaHostAllocDefault) != cudaSuccess) {
            orbx_set_error("cudaMalloc / cudaHostAlloc (matcher upload arena)"); return ORBX_E_CUDA;
        }
        cap = want; return ORBX_OK;
    }
    void reset() { off = 0; jobs.clear(); deferred = 0; }
    // returns the device address; the bytes are copied into the pinned mirror now (or at stage()) and travel at flush()
    void* put(const void* host, size_t bytes) {
        size_t o = (off + 255) & ~(size_t)255;
        if (o + bytes > cap) return nullptr;
        off = o + bytes;
        if (host && bytes) {
            if (defer && bytes >= 4096) { jobs.push_back(Job{o, host, bytes}); deferred += bytes; }
            else std::memcpy(hbase + o, host, bytes);
        }
        return dbase + o;
    }
    void stage() {                                                   // run the deferred copies
        if (jobs.empty()) return;
        static const int tmax = [] { const char* e = std::getenv("ORBX_STAGE_THREADS"); int v = e ? std::atoi(e) : 8; return v < 1 ? 1 : (v > 16 ? 16 : v); }();
        const int T = deferred >= (1u << 20) ? std::min<int>(tmax, (int)jobs.size()) : 1;
        auto work = [this](int t, int T_) { for (size_t j = (size_t)t; j < jobs.size(); j += (size_t)T_) std::memcpy(hbase + jobs[j].dst, jobs[j].src, jobs[j].bytes); };
        if (T <= 1) work(0, 1);
        else { std::vector<std::thread> th; for (int t = 1; t < T; ++t) th.emplace_back(work, t, T); work(0, T); for (auto& x : th) x.join(); }
        jobs.clear(); deferred = 0;
    }
    void release() { if (dbase) cudaFree(dbase); if (hbase) cudaFreeHost(hbase); dbase = hbase = nullptr; cap = 0; }
};

struct orbx_matcher {
    float nnratio; int checkOri; int device; cudaStream_t stream = nullptr; long long launches = 0;
    Arena arena; Arena cand_arena; UploadArena uparena;
    size_t cand_cap = 0;           // candidate entries the cand arena holds
    int cand_tries = 0;            // consecutive grow-and-repeat rounds of the current call (bounded)
    cudaEvent_t ev_left = nullptr, ev_right = nullptr;     // stereo batch: the matcher stream waits for both extractors' streams
    cudaEvent_t ev_stereo = nullptr;                       // ... and the extractors' streams wait for the stereo kernels before they touch their pyramids again
    float* stereo_out = nullptr; size_t stereo_out_n = 0;  // device results of the host-pointer stereo batch call
    float* stereo_scale = nullptr; std::vector<float> stereo_scale_host;   // mvScaleFactor | mvInvScaleFactor of the extractors, persistent
    bool profiling = false; std::vector<cudaEvent_t> prof_events;          // per-stage events of the batched calls (bench.py's roofline)
    void mark() { if (!profiling) return; cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return; cudaEventRecord(e, stream); prof_events.push_back(e); }
    uint8_t* dl_host = nullptr; size_t dl_cap = 0;     // pinned landing buffer for results that come back in one copy
    int ensure_download(size_t bytes) {
        if (bytes <= dl_cap) return ORBX_OK;
        if (dl_host) cudaFreeHost(dl_host);
        dl_host = nullptr; dl_cap = 0;
        const size_t want = bytes + (bytes >> 2) + 4096;
        if (cudaHostAlloc((void**)&dl_host, want, cudaHostAllocDefault) != cudaSuccess) { orbx_set_error("cudaHostAlloc (matcher download buffer)"); return ORBX_E_CUDA; }
        dl_cap = want; return ORBX_OK;
    }
};

static int flush_uploads(orbx_matcher* m) {
    m->uparena.stage();
    if (m->uparena.off) CU_TRY(cudaMemcpyAsync(m->uparena.dbase, m->uparena.hbase, m->uparena.off, cudaMemcpyHostToDevice, m->stream));
    return ORBX_OK;
}

struct FrameUpload { FrameDev dev; };

// bytes a frame needs in the arena
static size_t frame_bytes(const orbx_frame_view* f) {
    return pad((size_t)f->n * 28) + pad((size_t)f->n * 32) + pad((size_t)f->n * 4) + pad((size_t)f->nlevels * 4) +
           pad((size_t)(GRID_CELLS + 1) * 4) + 2 * pad((size_t)f->n * 4) + 4096;
}

static int check_frame(const orbx_frame_view* f) {
    if (!f || f->n < 0 || f->n >= (1 << 20) || (f->n && (!f->keys_un || !f->descriptors)) || f->nlevels <= 0 || f->nlevels > ORBX_MAX_LEVELS || !f->scale_factors)
        FAIL(ORBX_E_INVALID, "bad frame view");
    return ORBX_OK;
}

// stage a frame view for upload (uparena) and reserve its grid arrays; build_grid() launches the 64x48 grid build after the flush
static int upload_frame(orbx_matcher* m, const orbx_frame_view* f, FrameDev& d, uint32_t*& sort_keys) {
    const int n = f->n;
    KpM* keys = reinterpret_cast<KpM*>(m->uparena.put(n ? f->keys_un : nullptr, n ? (size_t)n * 28 : 28));
    uint8_t* desc = reinterpret_cast<uint8_t*>(m->uparena.put(n ? f->descriptors : nullptr, n ? (size_t)n * 32 : 32));
    float* ur = f->u_right ? reinterpret_cast<float*>(m->uparena.put(n ? f->u_right : nullptr, n ? (size_t)n * 4 : 4)) : nullptr;
    float* sc = reinterpret_cast<float*>(m->uparena.put(f->scale_factors, (size_t)f->nlevels * 4));
    int* cs = m->arena.get<int>(GRID_CELLS + 1); int* en = m->arena.get<int>(n ? n : 1); uint32_t* sk = m->arena.get<uint32_t>(n ? n : 1);
    if (!keys || !desc || !sc || !cs || !en || !sk || (f->u_right && !ur)) FAIL(ORBX_E_CUDA, "matcher arena exhausted");
    d.n = n; d.keys = keys; d.desc = desc; d.u_right = ur; d.min_x = f->min_x; d.min_y = f->min_y; d.max_x = f->max_x; d.max_y = f->max_y;
    d.gw_inv = f->grid_element_width_inv; d.gh_inv = f->grid_element_height_inv; d.scale = sc; d.nlevels = f->nlevels; d.cell_start = cs; d.entries = en;
    sort_keys = sk;
    return ORBX_OK;
}
static int build_grid(orbx_matcher* m, const FrameDev& d, uint32_t* sort_keys) {
    k_grid_build<<<1, 1024, 0, m->stream>>>(d.keys, d.n, d.min_x, d.min_y, d.gw_inv, d.gh_inv, sort_keys, const_cast<int*>(d.entries), const_cast<int*>(d.cell_start));
    LAUNCH_CHECK();
    return ORBX_OK;
}

// ---- device-resident Frame (include/orbx_b200.h, "Device-resident Frame") ----
template <typename T> struct FBuf {
    T* p = nullptr; size_t n = 0;
    int ensure(size_t count) {
        if (count <= n) return ORBX_OK;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        count += count >> 2;
        if (cudaMalloc((void**)&p, count * sizeof(T)) != cudaSuccess) { orbx_set_error("cudaMalloc (frame)"); return ORBX_E_CUDA; }
        n = count; return ORBX_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};
struct orbx_frame {
    int device; cudaStream_t stream = nullptr; long long launches = 0;
    int n = -1, nlevels = 0;       // n >= 0 once the grid is built (the frame is then usable by the matchers)
    int n_pending = 0, state = 0;  // keypoints taken / undistorted / gridded (FRAME_HAS_*)
    FBuf<KpM> keys_in, keys_un; FBuf<uint8_t> desc; FBuf<float> u_right, depth, scale, depth_img, bounds4;
    FBuf<int> cell_start, entries; FBuf<uint32_t> sort_keys;
    FrameDev dev;
    // ComputeImageBounds cache (the reference computes them once, Frame::mbInitialComputations)
    bool bounds_valid = false; orbx_camera bcam; int brows = 0, bcols = 0;
    float cmin_x = 0, cmax_x = 0, cmin_y = 0, cmax_y = 0;
    float min_x = 0, max_x = 0, min_y = 0, max_y = 0, gw_inv = 0, gh_inv = 0;          // bounds the current grid was built with
};
#define FRAME_HAS_KEYS 1
#define FRAME_HAS_UN 2
#define FRAME_HAS_GRID 4

// a matcher argument that is either a host frame view (uploaded + gridded per call) or a device-resident frame
struct FrameArg {
    const orbx_frame_view* v; const orbx_frame* f;
    int n() const { return v ? v->n : f->n; }
    int nlevels() const { return v ? v->nlevels : f->nlevels; }
};
static int check_frame_arg(const orbx_matcher* m, const FrameArg& a) {
    if (a.v) return check_frame(a.v);
    if (!a.f || a.f->n < 0) FAIL(ORBX_E_STATE, "device frame is empty (call orbx_frame_assign first)");
    if (m && a.f->device != m->device) FAIL(ORBX_E_INVALID, "frame and matcher must live on the same device");
    return ORBX_OK;
}
static size_t frame_arg_bytes(const FrameArg& a) { return a.v ? frame_bytes(a.v) : 4096; }
static int stage_frame(orbx_matcher* m, const FrameArg& a, FrameDev& d, uint32_t*& sort_keys) {
    if (a.v) return upload_frame(m, a.v, d, sort_keys);
    d = a.f->dev; sort_keys = nullptr;
    return ORBX_OK;
}
static int grid_frame(orbx_matcher* m, const FrameArg& a, const FrameDev& d, uint32_t* sort_keys) {
    return a.v ? build_grid(m, d, sort_keys) : ORBX_OK;        // a device frame built its grid in orbx_frame_assign
}

// host array -> staged upload (host != nullptr) or plain device scratch (host == nullptr)
template <typename T> static int up(orbx_matcher* m, const T* host, size_t count, T*& dev) {
    dev = host ? reinterpret_cast<T*>(m->uparena.put(count ? host : nullptr, (count ? count : 1) * sizeof(T))) : m->arena.get<T>(count ? count : 1);
    if (!dev) FAIL(ORBX_E_CUDA, "matcher arena exhausted");
    return ORBX_OK;
}

// COUNT pass, scan, FILL pass -- no host round trip in between: the candidate arena keeps the capacity of earlier calls (at least
// 64 per query); the FILL pass and the resolve kernels refuse to touch it when the total exceeds that capacity, and the caller, which
// reads the total back together with its results, grows the arena and repeats the call (rare).
static int window_search(orbx_matcher* m, QueryParams& P, const FrameDev& F, int*& counts, int*& offsets, uint32_t*& cand, uint2*& pre_best) {
    cudaStream_t s = m->stream;
    const int nq = P.nq;
    counts = m->arena.get<int>(nq + 1); offsets = m->arena.get<int>(nq + 2); pre_best = m->arena.get<uint2>(nq + 1);
    if (!counts || !offsets || !pre_best) FAIL(ORBX_E_CUDA, "matcher arena exhausted");
    const size_t want = std::max<size_t>((size_t)nq * 64, 1 << 16);
    if (m->cand_cap < want) {
        const int rc = m->cand_arena.reserve(want * 4);
        m->cand_cap = m->cand_arena.cap / 4;                     // 0 after a failed grow: nothing stale survives
        if (rc) return rc;
    }
    cand = reinterpret_cast<uint32_t*>(m->cand_arena.base);
    if (nq == 0) { CU_TRY(cudaMemsetAsync(offsets, 0, 8, s)); return ORBX_OK; }
    k_window_search<false><<<(nq + 3) / 4, 128, 0, s>>>(P, F, counts, nullptr, nullptr, 0, nullptr);
    LAUNCH_CHECK();
    k_scan_counts<<<1, 1024, 0, s>>>(counts, nq, offsets);
    LAUNCH_CHECK();
    k_window_search<true><<<(nq + 3) / 4, 128, 0, s>>>(P, F, counts, offsets, cand, (int)std::min<size_t>(m->cand_cap, 0x7FFFFFFF), pre_best);
    LAUNCH_CHECK();
    return ORBX_OK;
}
// after the call's final synchronisation: did the candidate lists fit?  If not, grow and tell the caller to run again.
// Returns 0 = they fit, 1 = the arena has been grown (run the call again), < 0 = error.  A failed grow leaves the arena empty
// (Arena::reserve frees first), so the capacity is re-read from it either way and the error is returned instead of a retry.
static int cand_overflow(orbx_matcher* m, long long total) {
    if (total >= 0 && (size_t)total <= m->cand_cap) { m->cand_tries = 0; return 0; }
    if (total < 0 || ++m->cand_tries > 2) { m->cand_tries = 0; FAIL(ORBX_E_OVERFLOW, "candidate lists do not fit after growing the arena"); }
    const int rc = m->cand_arena.reserve(((size_t)total + (size_t)total / 2 + 1024) * 4);
    m->cand_cap = m->cand_arena.cap / 4;
    if (rc) { m->cand_tries = 0; return rc; }
    return 1;
}
#define CAND_RETRY(total) do { const int _ov = cand_overflow(m, (total)); if (_ov < 0) return _ov; if (_ov) goto retry; } while (0)

// Results come back in ONE copy into a pinned landing buffer (pageable destinations make every cudaMemcpyAsync a blocking, staged
// transfer of its own): the pieces are gathered into a contiguous device block first (device-to-device copies are cheap to enqueue).
struct Gather {
    orbx_matcher* m; uint8_t* dblock = nullptr; size_t bytes = 0, off = 0;
    int begin(size_t total) {
        bytes = (total + 15) & ~(size_t)15; off = 0;
        dblock = m->arena.get<uint8_t>(bytes);
        if (!dblock) { orbx_set_error("matcher arena exhausted"); return ORBX_E_CUDA; }
        return m->ensure_download(bytes);
    }
    // returns the offset of the piece inside the block
    size_t add(const void* dsrc, size_t n) {
        const size_t o = off; off += (n + 3) & ~(size_t)3;
        if (n && cudaMemcpyAsync(dblock + o, dsrc, n, cudaMemcpyDeviceToDevice, m->stream) != cudaSuccess) failed = true;
        return o;
    }
    bool failed = false;
    int finish() {
        if (failed || cudaMemcpyAsync(m->dl_host, dblock, off, cudaMemcpyDeviceToHost, m->stream) != cudaSuccess || cudaStreamSynchronize(m->stream) != cudaSuccess) {
            orbx_set_error(std::string("result download: ") + cudaGetErrorString(cudaGetLastError())); return ORBX_E_CUDA;
        }
        return ORBX_OK;
    }
    const uint8_t* host(size_t o) const { return m->dl_host + o; }
};

#define RESOLVE_SMEM_BYTES (160 * 1024)

extern "C" {

int orbx_matcher_create(float nnratio, int check_orientation, int device, orbx_matcher** out) {
    if (!out) FAIL(ORBX_E_INVALID, "null out");
    *out = nullptr;
    int ndev = 0;
    CU_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) FAIL(ORBX_E_CUDA, "no such CUDA device (this library has no CPU fallback)");
    CU_TRY(cudaSetDevice(device));
    orbx_matcher* m = new orbx_matcher();
    m->nnratio = nnratio; m->checkOri = check_orientation != 0; m->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete m; FAIL(ORBX_E_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); }
    cudaFuncSetAttribute(k_resolve_init, cudaFuncAttributeMaxDynamicSharedMemorySize, RESOLVE_SMEM_BYTES);   // per device
    cudaFuncSetAttribute(k_resolve_init_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, RESOLVE_SMEM_BYTES);
    cudaFuncSetAttribute(k_resolve_proj_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, RESOLVE_SMEM_BYTES);
    cudaFuncSetAttribute(k_resolve_proj_points, cudaFuncAttributeMaxDynamicSharedMemorySize, RESOLVE_SMEM_BYTES);
    *out = m;
    return ORBX_OK;
}
void orbx_matcher_destroy(orbx_matcher* m) {
    if (!m) return;
    cudaSetDevice(m->device); cudaStreamSynchronize(m->stream);
    if (m->arena.base) cudaFree(m->arena.base);
    if (m->cand_arena.base) cudaFree(m->cand_arena.base);
    m->uparena.release();
    if (m->dl_host) cudaFreeHost(m->dl_host);
    if (m->stereo_out) cudaFree(m->stereo_out);
    if (m->stereo_scale) cudaFree(m->stereo_scale);
    if (m->ev_left) cudaEventDestroy(m->ev_left);
    if (m->ev_stereo) cudaEventDestroy(m->ev_stereo);
    if (m->ev_right) cudaEventDestroy(m->ev_right);
    cudaStreamDestroy(m->stream);
    delete m;
}
void* orbx_matcher_stream(orbx_matcher* m) { return m ? (void*)m->stream : nullptr; }
// Per-stage device timing of the batched calls (bench.py): while enabled, orbx_search_for_initialization[_frames]_batch records an event
// before each of its 5 launches and after the last (6 marks: grid build, window count, scan, window fill, resolve) and
// orbx_compute_stereo_matches_batch[_device] 3 marks (row-band Hamming + SAD, median cut).  collect synchronises, sums the elapsed ms of
// the `nstages` consecutive intervals of every profiled call, and clears the events.
int orbx_matcher_profile_enable(orbx_matcher* m, int on) { if (!m) return ORBX_E_INVALID; m->profiling = on != 0; return ORBX_OK; }
int orbx_matcher_profile_collect(orbx_matcher* m, int nstages, double* stage_ms, int* ncalls) {
    if (!m || nstages <= 0 || !stage_ms || !ncalls) return ORBX_E_INVALID;
    if (cudaSetDevice(m->device) != cudaSuccess) return ORBX_E_CUDA;
    CU_TRY(cudaStreamSynchronize(m->stream));
    for (int i = 0; i < nstages; ++i) stage_ms[i] = 0.0;
    const size_t per = (size_t)nstages + 1, calls = m->prof_events.size() / per;
    for (size_t c = 0; c < calls; ++c)
        for (int i = 0; i < nstages; ++i) { float ms = 0.f; cudaEventElapsedTime(&ms, m->prof_events[c * per + i], m->prof_events[c * per + i + 1]); stage_ms[i] += ms; }
    for (cudaEvent_t e : m->prof_events) cudaEventDestroy(e);
    m->prof_events.clear();
    *ncalls = (int)calls;
    return ORBX_OK;
}
long long orbx_matcher_launch_count(const orbx_matcher* m) { return m ? m->launches : 0; }

int orbx_descriptor_distance(orbx_matcher* m, const uint8_t* a, const uint8_t* b, int n, int* out) {
    if (!m || n < 0 || (n && (!a || !b || !out))) FAIL(ORBX_E_INVALID, "bad arguments");
    if (n == 0) return ORBX_OK;
    CU_TRY(cudaSetDevice(m->device));
    int rc = m->arena.reserve(pad((size_t)n * 4) + 4096); if (rc) return rc;
    if ((rc = m->uparena.reserve(2 * pad((size_t)n * 32) + 4096))) return rc;
    m->arena.reset(); m->uparena.reset();
    uint8_t *da, *db; int* dout;
    if ((rc = up(m, a, (size_t)n * 32, da)) || (rc = up(m, b, (size_t)n * 32, db)) || (rc = up<int>(m, nullptr, n, dout))) return rc;
    if ((rc = flush_uploads(m))) return rc;
    k_descriptor_distance<<<(n + 127) / 128, 128, 0, m->stream>>>(reinterpret_cast<const uint4*>(da), reinterpret_cast<const uint4*>(db), n, dout);
    LAUNCH_CHECK();
    CU_TRY(cudaMemcpyAsync(out, dout, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(cudaStreamSynchronize(m->stream));
    return ORBX_OK;
}

static int search_init_impl(orbx_matcher* m, const FrameArg A1, const FrameArg A2, float* prev_matched_xy, int* matches12, int window_size, int* nmatches) {
    if (!m || !nmatches) FAIL(ORBX_E_INVALID, "null argument");
    int rc;
    if ((rc = check_frame_arg(m, A1)) || (rc = check_frame_arg(m, A2))) return rc;
    const int n1 = A1.n(), n2 = A2.n();
    if (n1 && (!prev_matched_xy || !matches12)) FAIL(ORBX_E_INVALID, "null buffer");
    *nmatches = 0;
    if (n1 == 0) return ORBX_OK;
    CU_TRY(cudaSetDevice(m->device));
    const size_t need = frame_arg_bytes(A1) + frame_arg_bytes(A2) + pad((size_t)n1 * 8) + 6 * pad((size_t)(n1 + n2 + 2) * 4) + pad((size_t)n1 * 12 + 256) + 8192;
    if ((rc = m->arena.reserve(need)) || (rc = m->uparena.reserve(need))) return rc;
retry:
    m->arena.reset(); m->uparena.reset();
    FrameDev d2; uint32_t* sk2;
    if ((rc = stage_frame(m, A2, d2, sk2))) return rc;
    const KpM* k1; const uint8_t* desc1; float* prev;
    if (A1.v) {
        KpM* k1u; uint8_t* d1u;
        if ((rc = up(m, reinterpret_cast<const KpM*>(A1.v->keys_un), (size_t)n1, k1u)) || (rc = up(m, A1.v->descriptors, (size_t)n1 * 32, d1u))) return rc;
        k1 = k1u; desc1 = d1u;
    } else { k1 = A1.f->dev.keys; desc1 = A1.f->dev.desc; }
    if ((rc = up(m, prev_matched_xy, (size_t)n1 * 2, prev))) return rc;
    if ((rc = flush_uploads(m)) || (rc = grid_frame(m, A2, d2, sk2))) return rc;
    QueryParams P; std::memset(&P, 0, sizeof(P));
    P.mode = MODE_INIT; P.nq = n1; P.q_keys = k1; P.q_desc = desc1; P.q_xy = prev; P.window = (float)window_size;
    int *counts, *offsets; uint32_t* cand; uint2* pre;
    if ((rc = window_search(m, P, d2, counts, offsets, cand, pre))) return rc;
    int* md = m->arena.get<int>(n2 + 1); int* m21 = m->arena.get<int>(n2 + 1); int* m12 = m->arena.get<int>(n1); int* binof = m->arena.get<int>(n1); int* dn = m->arena.get<int>(1);
    if (!md || !m21 || !m12 || !binof || !dn) FAIL(ORBX_E_CUDA, "matcher arena exhausted");
    const int resolve_smem = RESOLVE_SMEM_BYTES;
    k_resolve_init<<<1, RESOLVE_THREADS, resolve_smem, m->stream>>>(n1, n2, k1, d2.keys, counts, offsets, cand, pre, (int)std::min<size_t>(m->cand_cap, 0x7FFFFFFF), m->nnratio, m->checkOri, resolve_smem / 4,
                                                                     md, m21, m12, binof, prev, dn);
    LAUNCH_CHECK();
    Gather g{m};
    if ((rc = g.begin(8 + (size_t)n1 * 12 + 64))) return rc;
    const size_t o_total = g.add(offsets + n1, 4), o_n = g.add(dn, 4), o_m12 = g.add(m12, (size_t)n1 * 4), o_prev = g.add(prev, (size_t)n1 * 8);
    if ((rc = g.finish())) return rc;
    int total; std::memcpy(&total, g.host(o_total), 4);
    CAND_RETRY(total);                       // the caller's arrays are only written below, so the inputs are still intact
    std::memcpy(nmatches, g.host(o_n), 4);
    std::memcpy(matches12, g.host(o_m12), (size_t)n1 * 4);
    std::memcpy(prev_matched_xy, g.host(o_prev), (size_t)n1 * 8);
    return ORBX_OK;
}

// SearchForInitialization for n_pairs independent frame pairs: one upload, five launches, one download (see k_match.cuh, InitPair)
static int search_init_batch_impl(orbx_matcher* m, int n_pairs, const FrameArg* A1, const FrameArg* A2, float* const* prev_matched_xy, int* const* matches12,
                                  int window_size, int* nmatches) {
    if (!m || !nmatches || n_pairs < 0 || (n_pairs && (!A1 || !A2 || !prev_matched_xy || !matches12))) FAIL(ORBX_E_INVALID, "null argument");
    if (n_pairs == 0) return ORBX_OK;
    int rc;
    size_t need_up = 8192 + pad(sizeof(InitPair) * (size_t)n_pairs), need_dev = 8192, n1sum = 0;
    int n1max = 0, n2max = 0;
    for (int p = 0; p < n_pairs; ++p) {
        if ((rc = check_frame_arg(m, A1[p])) || (rc = check_frame_arg(m, A2[p]))) return rc;
        const int n1 = A1[p].n(), n2 = A2[p].n();
        if (n1 && (!prev_matched_xy[p] || !matches12[p])) FAIL(ORBX_E_INVALID, "null buffer");
        need_up += frame_arg_bytes(A1[p]) + frame_arg_bytes(A2[p]) + pad((size_t)n1 * 8) + 1024;
        need_dev += 4 * pad((size_t)(n1 + 2) * 4) + pad((size_t)(n1 + 1) * 8) + 2 * pad((size_t)(n2 + 1) * 4) + frame_arg_bytes(A2[p]) + 2048;
        n1max = std::max(n1max, n1); n2max = std::max(n2max, n2); n1sum += (size_t)n1;
    }
    need_dev += pad((size_t)n_pairs * 8) + pad(n1sum * 4);
    CU_TRY(cudaSetDevice(m->device));
    if ((rc = m->arena.reserve(need_dev)) || (rc = m->uparena.reserve(need_up))) return rc;
    // every pair owns a fixed slice of the candidate arena (the lists of SearchForInitialization hold the level-0 features inside the window only)
    size_t slice = std::max<size_t>((size_t)16 * (size_t)std::max(n1max, 1), 4096);
    cudaStream_t s = m->stream;
    std::vector<InitPair> hp((size_t)n_pairs);
retry:
    if (m->cand_cap < slice * (size_t)n_pairs) {
        rc = m->cand_arena.reserve(slice * (size_t)n_pairs * 4);
        m->cand_cap = m->cand_arena.cap / 4;
        if (rc) return rc;
    }
    m->arena.reset(); m->uparena.reset();
    int* d_nm = m->arena.get<int>(n_pairs); int* d_tot = m->arena.get<int>(n_pairs); int* d_m12 = m->arena.get<int>(n1sum ? n1sum : 1);
    if (!d_nm || !d_tot || !d_m12) FAIL(ORBX_E_CUDA, "matcher arena exhausted");
    std::vector<size_t> o1((size_t)n_pairs + 1, 0);
    std::vector<float*> dprev((size_t)n_pairs);
    for (int p = 0; p < n_pairs; ++p) {
        InitPair& ip = hp[p]; std::memset(&ip, 0, sizeof(ip));
        const int n1 = A1[p].n(), n2 = A2[p].n();
        uint32_t* sk2 = nullptr;
        if ((rc = stage_frame(m, A2[p], ip.F2, sk2))) return rc;
        ip.sort_keys = sk2;
        if (A1[p].v) {
            KpM* k1u; uint8_t* d1u;
            if ((rc = up(m, reinterpret_cast<const KpM*>(A1[p].v->keys_un), (size_t)n1, k1u)) || (rc = up(m, A1[p].v->descriptors, (size_t)n1 * 32, d1u))) return rc;
            ip.k1 = k1u; ip.d1 = d1u;
        } else { ip.k1 = A1[p].f->dev.keys; ip.d1 = A1[p].f->dev.desc; }
        ip.n1 = n1;
        ip.counts = m->arena.get<int>(n1 + 1); ip.offsets = m->arena.get<int>(n1 + 2); ip.pre = m->arena.get<uint2>(n1 + 1);
        ip.md = m->arena.get<int>(n2 + 1); ip.m21 = m->arena.get<int>(n2 + 1); ip.binof = m->arena.get<int>(n1 + 1);
        if (!ip.counts || !ip.offsets || !ip.pre || !ip.md || !ip.m21 || !ip.binof) FAIL(ORBX_E_CUDA, "matcher arena exhausted");
        ip.m12 = d_m12 + o1[p]; o1[p + 1] = o1[p] + (size_t)n1;
        ip.cand = reinterpret_cast<uint32_t*>(m->cand_arena.base) + (size_t)p * slice;
        ip.nmatches = d_nm + p; ip.total = d_tot + p;
    }
    for (int p = 0; p < n_pairs; ++p) {                                // vbPrevMatched of all pairs back to back: they come back in one copy
        if ((rc = up(m, prev_matched_xy[p], (size_t)A1[p].n() * 2, hp[p].prev))) return rc;
        dprev[p] = hp[p].prev;
    }
    InitPair* d_pairs;
    if ((rc = up(m, hp.data(), (size_t)n_pairs, d_pairs))) return rc;
    if ((rc = flush_uploads(m))) return rc;
    CU_TRY(cudaMemsetAsync(d_nm, 0, (size_t)n_pairs * 4, s));
    const int icap = (int)std::min<size_t>(slice, 0x7FFFFFFF);
    m->mark();
    k_grid_build_pairs<<<n_pairs, 1024, 0, s>>>(d_pairs);
    LAUNCH_CHECK();
    m->mark();
    if (n1max > 0) {
        const dim3 grid((n1max + 3) / 4, n_pairs);
        k_window_search_pairs<false><<<grid, 128, 0, s>>>(d_pairs, (float)window_size, icap);
        LAUNCH_CHECK();
        m->mark();
        k_scan_counts_pairs<<<n_pairs, 1024, 0, s>>>(d_pairs);
        LAUNCH_CHECK();
        m->mark();
        k_window_search_pairs<true><<<grid, 128, 0, s>>>(d_pairs, (float)window_size, icap);
        LAUNCH_CHECK();
        m->mark();
        // shared memory of one resolve CTA: the lists of a typical pair (window 100 on a VGA frame: ~7 candidates per feature of F1) staged
        // with the state arrays; larger pairs read their lists from global memory (same result)
        const size_t want_words = (size_t)5 * n1max + (size_t)3 * n2max + (size_t)10 * n1max + 64;
        const int resolve_smem = (int)std::min<size_t>(RESOLVE_SMEM_BYTES, (want_words * 4 + 1023) & ~(size_t)1023);
        k_resolve_init_pairs<<<n_pairs, RESOLVE_THREADS, resolve_smem, s>>>(d_pairs, icap, m->nnratio, m->checkOri, resolve_smem / 4);
        LAUNCH_CHECK();
        m->mark();
    }
    // results: [nmatches | totals | all m12] are contiguous in the arena; the prev arrays live in the upload arena (contiguous per pair)
    if ((rc = m->ensure_download(2 * pad((size_t)n_pairs * 4) + pad(n1sum * 4) + n1sum * 8 + (size_t)n_pairs * 256 + 4096))) return rc;
    uint8_t* hd = m->dl_host;
    const size_t o_tot = pad((size_t)n_pairs * 4), o_m12 = 2 * o_tot, o_prev = o_m12 + pad(n1sum * 4);
    CU_TRY(cudaMemcpyAsync(hd, d_nm, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(hd + o_tot, d_tot, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, s));
    if (n1sum) CU_TRY(cudaMemcpyAsync(hd + o_m12, d_m12, n1sum * 4, cudaMemcpyDeviceToHost, s));
    // the prev arrays were staged one after the other (256-byte aligned): fetch the span that covers them in one copy
    const uint8_t* pv_lo = reinterpret_cast<const uint8_t*>(dprev[0]);
    const uint8_t* pv_hi = reinterpret_cast<const uint8_t*>(dprev[n_pairs - 1]) + (size_t)A1[n_pairs - 1].n() * 8;
    if (n1sum) {
        CU_TRY(cudaMemcpyAsync(hd + o_prev, pv_lo, (size_t)(pv_hi - pv_lo), cudaMemcpyDeviceToHost, s));
    }
    CU_TRY(cudaStreamSynchronize(s));
    if (n1max > 0) {
        const int* tot = reinterpret_cast<const int*>(hd + o_tot);
        long long worst = 0;
        for (int p = 0; p < n_pairs; ++p) worst = std::max<long long>(worst, tot[p]);
        if ((size_t)worst > slice) {                                   // some pair's lists did not fit its slice: grow every slice and run again
            if (++m->cand_tries > 2) { m->cand_tries = 0; FAIL(ORBX_E_OVERFLOW, "candidate lists do not fit after growing the arena"); }
            slice = (size_t)worst + (size_t)worst / 2 + 1024;
            goto retry;
        }
        m->cand_tries = 0;
    }
    std::memcpy(nmatches, hd, (size_t)n_pairs * 4);
    for (int p = 0; p < n_pairs; ++p) {
        const int n1 = A1[p].n();
        if (!n1) continue;
        std::memcpy(matches12[p], hd + o_m12 + o1[p] * 4, (size_t)n1 * 4);
        std::memcpy(prev_matched_xy[p], hd + o_prev + (reinterpret_cast<const uint8_t*>(dprev[p]) - pv_lo), (size_t)n1 * 8);
    }
    return ORBX_OK;
}

// x3Dc = Rcw * x3Dw + tcw, invzc = 1.0 / z, (u, v) = (fx xc invzc + cx, fy yc invzc + cy)  (ORBmatcher.cc:1608-1623) for every feature of the
// last frame that has a usable map point.  The product is evaluated as cv::gemm does for a 3 x 3 by 3 x 1 float matrix -- hand-unrolled
// FLOAT arithmetic ((r0 x + r1 y) + r2 z), then + t (pinned against cv2.gemm of OpenCV 4.13 by the gemm_3x3_3x1_c fixtures under tests/golden) -- with every
// operation individually rounded; 1.0 / z is the double division the reference's literal asks for, rounded to float.
struct ProjPose { float R[9], t[3], fx, fy, cx, cy; };
__global__ void k_project_frame(int n, const float* __restrict__ xyz, const uint8_t* __restrict__ has, ProjPose P, float min_x, float max_x, float min_y, float max_y,
                                float* __restrict__ uv, float* __restrict__ iz, uint8_t* __restrict__ valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float u = 0.f, v = 0.f, invzc = 0.f; uint8_t ok = 0;
    if (has[i]) {
        const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P.R[0], x), __fmul_rn(P.R[1], y)), __fmul_rn(P.R[2], z)), P.t[0]);
        const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P.R[3], x), __fmul_rn(P.R[4], y)), __fmul_rn(P.R[5], z)), P.t[1]);
        const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P.R[6], x), __fmul_rn(P.R[7], y)), __fmul_rn(P.R[8], z)), P.t[2]);
        invzc = __double2float_rn(__ddiv_rn(1.0, (double)zc));
        if (!(invzc < 0)) {
            u = __fadd_rn(__fmul_rn(__fmul_rn(P.fx, xc), invzc), P.cx);
            v = __fadd_rn(__fmul_rn(__fmul_rn(P.fy, yc), invzc), P.cy);
            ok = !(u < min_x || u > max_x || v < min_y || v > max_y);
        }
        if (!ok) { u = 0.f; v = 0.f; invzc = 0.f; }
    }
    uv[2 * i] = u; uv[2 * i + 1] = v; iz[i] = invzc; valid[i] = ok;
}
struct ProjPoseArgs { const float* world; const uint8_t* has; ProjPose P; float* uv_out; float* invz_out; uint8_t* valid_out; };

static int search_proj_frame_impl(orbx_matcher* m, const FrameArg cur, int n_last, const float* proj_uv, const float* proj_invz,
                                    const int* last_octave, const float* last_angle, const uint8_t* mp_desc, const uint8_t* valid, const uint8_t* mp_observed,
                                    const uint8_t* cur_occupied, float th, int forward, int backward, float mbf, int* cur_match, int* nmatches,
                                  int max_dist = M_TH_HIGH, int no_ur = 0, int check_ori = -1 /* -1: the matcher's own setting */, const ProjPoseArgs* pose = nullptr) {
    if (!m || !nmatches) FAIL(ORBX_E_INVALID, "null argument");
    int rc;
    if ((rc = check_frame_arg(m, cur))) return rc;
    if (pose) { if (n_last && (!pose->world || !pose->has)) FAIL(ORBX_E_INVALID, "bad arguments"); proj_uv = pose->world; proj_invz = pose->world; valid = pose->has; }   // (validated below as "present")
    if (n_last < 0 || n_last >= (1 << 20) || (n_last && (!proj_uv || !proj_invz || !last_octave || !last_angle || !mp_desc || !valid)) || (cur.n() && !cur_match))
        FAIL(ORBX_E_INVALID, "bad arguments");
    *nmatches = 0;
    for (int i = 0; i < n_last; ++i) if (valid[i] && (last_octave[i] < 0 || last_octave[i] >= cur.nlevels())) FAIL(ORBX_E_INVALID, "octave out of range");
    CU_TRY(cudaSetDevice(m->device));
    const int nc = cur.n();
    const size_t need = frame_arg_bytes(cur) + pad((size_t)n_last * 32) + 8 * pad((size_t)(n_last + nc + 2) * 8) + pad((size_t)nc * 4 + 256) + 8192;
    if ((rc = m->arena.reserve(need)) || (rc = m->uparena.reserve(need))) return rc;
retry:
    m->arena.reset(); m->uparena.reset();
    FrameDev dc; uint32_t* skc;
    if ((rc = stage_frame(m, cur, dc, skc))) return rc;
    float *uv, *iz, *la; int* lo; uint8_t *dd, *va, *ob = nullptr, *oc = nullptr;
    float* wxyz = nullptr; uint8_t* whas = nullptr;
    if (pose) {                                                       // the projection itself runs on the device: upload world points, not (u, v)
        uv = m->arena.get<float>((size_t)n_last * 2 + 2); iz = m->arena.get<float>((size_t)n_last + 1); va = m->arena.get<uint8_t>((size_t)n_last + 16);
        if (!uv || !iz || !va) FAIL(ORBX_E_CUDA, "matcher arena exhausted");
        if ((rc = up(m, pose->world, (size_t)n_last * 3, wxyz)) || (rc = up(m, pose->has, (size_t)n_last, whas))) return rc;
    } else if ((rc = up(m, proj_uv, (size_t)n_last * 2, uv)) || (rc = up(m, proj_invz, (size_t)n_last, iz)) || (rc = up(m, valid, (size_t)n_last, va))) return rc;
    if ((rc = up(m, last_angle, (size_t)n_last, la)) || (rc = up(m, last_octave, (size_t)n_last, lo)) || (rc = up(m, mp_desc, (size_t)n_last * 32, dd))) return rc;
    if (mp_observed && (rc = up(m, mp_observed, (size_t)n_last, ob))) return rc;
    if (cur_occupied && (rc = up(m, cur_occupied, (size_t)nc, oc))) return rc;
    if ((rc = flush_uploads(m)) || (rc = grid_frame(m, cur, dc, skc))) return rc;
    if (pose && n_last) {
        k_project_frame<<<(n_last + 127) / 128, 128, 0, m->stream>>>(n_last, wxyz, whas, pose->P, dc.min_x, dc.max_x, dc.min_y, dc.max_y, uv, iz, va);
        LAUNCH_CHECK();
    }
    QueryParams P; std::memset(&P, 0, sizeof(P));
    P.mode = MODE_PROJ_FRAME; P.nq = n_last; P.q_desc = dd; P.q_xy = uv; P.q_invz = iz; P.q_octave = lo; P.q_valid = va; P.th = th; P.forward = forward; P.backward = backward; P.mbf = mbf; P.no_ur = no_ur;
    int *counts, *offsets; uint32_t* cand; uint2* pre;
    if ((rc = window_search(m, P, dc, counts, offsets, cand, pre))) return rc;
    int* occ = m->arena.get<int>(nc + 1); int* cm = m->arena.get<int>(nc + 1); int* pushes = m->arena.get<int>(2 * (size_t)n_last + 2); int* dn = m->arena.get<int>(1);
    if (!occ || !cm || !pushes || !dn) FAIL(ORBX_E_CUDA, "matcher arena exhausted");
    k_resolve_proj_frame<<<1, RESOLVE_THREADS, RESOLVE_SMEM_BYTES, m->stream>>>(n_last, nc, dc.keys, la, ob, oc, counts, offsets, cand, pre, (int)std::min<size_t>(m->cand_cap, 0x7FFFFFFF), check_ori < 0 ? m->checkOri : check_ori, max_dist, RESOLVE_SMEM_BYTES / 4, occ, cm, pushes, dn);
    LAUNCH_CHECK();
    Gather g{m};
    if ((rc = g.begin(8 + (size_t)nc * 4 + 64))) return rc;
    const size_t o_total = g.add(offsets + n_last, 4), o_n = g.add(dn, 4), o_cm = g.add(cm, (size_t)nc * 4);
    if ((rc = g.finish())) return rc;
    int total; std::memcpy(&total, g.host(o_total), 4);
    CAND_RETRY(total);
    std::memcpy(nmatches, g.host(o_n), 4);
    if (nc) std::memcpy(cur_match, g.host(o_cm), (size_t)nc * 4);
    if (pose && n_last && (pose->uv_out || pose->invz_out || pose->valid_out)) {       // test / inspection taps of the device projection
        if (pose->uv_out) CU_TRY(cudaMemcpyAsync(pose->uv_out, uv, (size_t)n_last * 8, cudaMemcpyDeviceToHost, m->stream));
        if (pose->invz_out) CU_TRY(cudaMemcpyAsync(pose->invz_out, iz, (size_t)n_last * 4, cudaMemcpyDeviceToHost, m->stream));
        if (pose->valid_out) CU_TRY(cudaMemcpyAsync(pose->valid_out, va, (size_t)n_last, cudaMemcpyDeviceToHost, m->stream));
        CU_TRY(cudaStreamSynchronize(m->stream));
    }
    return ORBX_OK;
}

static int search_proj_points_impl(orbx_matcher* m, const FrameArg F, int n_points, const float* track_uv, const float* track_ur, const int* track_level,
                                     const float* track_view_cos, const uint8_t* mp_desc, const uint8_t* mp_observed, const uint8_t* f_occupied, float th,
                                     int* f_match, int* nmatches) {
    if (!m || !nmatches) FAIL(ORBX_E_INVALID, "null argument");
    int rc;
    if ((rc = check_frame_arg(m, F))) return rc;
    if (n_points < 0 || n_points >= (1 << 20) || (n_points && (!track_uv || !track_ur || !track_level || !track_view_cos || !mp_desc)) || (F.n() && !f_match))
        FAIL(ORBX_E_INVALID, "bad arguments");
    *nmatches = 0;
    for (int i = 0; i < n_points; ++i) if (track_level[i] < 0 || track_level[i] >= F.nlevels()) FAIL(ORBX_E_INVALID, "predicted level out of range");
    CU_TRY(cudaSetDevice(m->device));
    const int nf = F.n();
    const size_t need = frame_arg_bytes(F) + pad((size_t)n_points * 32) + 8 * pad((size_t)(n_points + nf + 2) * 8) + pad((size_t)nf * 4 + 256) + 8192;
    if ((rc = m->arena.reserve(need)) || (rc = m->uparena.reserve(need))) return rc;
retry:
    m->arena.reset(); m->uparena.reset();
    FrameDev df; uint32_t* skf;
    if ((rc = stage_frame(m, F, df, skf))) return rc;
    float *uv, *ur, *vc; int* lv; uint8_t *dd, *ob = nullptr, *oc = nullptr;
    if ((rc = up(m, track_uv, (size_t)n_points * 2, uv)) || (rc = up(m, track_ur, (size_t)n_points, ur)) || (rc = up(m, track_view_cos, (size_t)n_points, vc)) ||
        (rc = up(m, track_level, (size_t)n_points, lv)) || (rc = up(m, mp_desc, (size_t)n_points * 32, dd))) return rc;
    if (mp_observed && (rc = up(m, mp_observed, (size_t)n_points, ob))) return rc;
    if (f_occupied && (rc = up(m, f_occupied, (size_t)nf, oc))) return rc;
    if ((rc = flush_uploads(m)) || (rc = grid_frame(m, F, df, skf))) return rc;
    QueryParams P; std::memset(&P, 0, sizeof(P));
    P.mode = MODE_PROJ_POINTS; P.nq = n_points; P.q_desc = dd; P.q_xy = uv; P.q_octave = lv; P.q_ur = ur; P.q_viewcos = vc; P.th = th;
    int *counts, *offsets; uint32_t* cand; uint2* pre;
    if ((rc = window_search(m, P, df, counts, offsets, cand, pre))) return rc;
    int* occ = m->arena.get<int>(nf + 1); int* fm = m->arena.get<int>(nf + 1); int* dn = m->arena.get<int>(1);
    if (!occ || !fm || !dn) FAIL(ORBX_E_CUDA, "matcher arena exhausted");
    k_resolve_proj_points<<<1, RESOLVE_THREADS, RESOLVE_SMEM_BYTES, m->stream>>>(n_points, nf, df.keys, ob, oc, counts, offsets, cand, pre, (int)std::min<size_t>(m->cand_cap, 0x7FFFFFFF), m->nnratio, RESOLVE_SMEM_BYTES / 4, occ, fm, dn);
    LAUNCH_CHECK();
    Gather g{m};
    if ((rc = g.begin(8 + (size_t)nf * 4 + 64))) return rc;
    const size_t o_total = g.add(offsets + n_points, 4), o_n = g.add(dn, 4), o_fm = g.add(fm, (size_t)nf * 4);
    if ((rc = g.finish())) return rc;
    int total; std::memcpy(&total, g.host(o_total), 4);
    CAND_RETRY(total);
    std::memcpy(nmatches, g.host(o_n), 4);
    if (nf) std::memcpy(f_match, g.host(o_fm), (size_t)nf * 4);
    return ORBX_OK;
}

// capacity of the per-row candidate lists of one stereo pair (k_stereo_rows): a right keypoint of level l enters at most 2 * (2 * scale[l]) + 3 rows
static size_t stereo_row_entries(const float* scale, int nlevels, int nr_max, int nRows) {
    float smax = 1.f;
    for (int l = 0; l < nlevels; ++l) smax = std::max(smax, scale[l]);
    const double band = std::min<double>((double)nRows, 4.0 * (double)smax + 3.0);
    return (size_t)std::max(nr_max, 1) * (size_t)std::ceil(band) + 32;
}

int orbx_compute_stereo_matches(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, const orbx_keypoint* keys_left, const uint8_t* desc_left, int nl,
                                const orbx_keypoint* keys_right, const uint8_t* desc_right, int nr, float mb, float mbf, float* u_right, float* depth) {
    if (!m || !left || !right || nl < 0 || nr < 0 || nr >= (1 << 20) || (nl && (!keys_left || !desc_left || !u_right || !depth)) || (nr && (!keys_right || !desc_right)))
        FAIL(ORBX_E_INVALID, "bad arguments");
    if (nl == 0) return ORBX_OK;
    CU_TRY(cudaSetDevice(m->device));
    OrbxPyramidInfo L, R;
    int rc;
    if ((rc = orbx_internal_pyramid(left, &L)) || (rc = orbx_internal_pyramid(right, &R))) return rc;
    if (L.device != m->device || R.device != m->device) FAIL(ORBX_E_INVALID, "extractors and matcher must live on the same device");
    if (L.nlevels != R.nlevels) FAIL(ORBX_E_INVALID, "left / right extractors differ in nlevels");
    for (int i = 0; i < nl; ++i) if (keys_left[i].octave < 0 || keys_left[i].octave >= L.nlevels) FAIL(ORBX_E_INVALID, "octave out of range");
    for (int i = 0; i < nr; ++i) if (keys_right[i].octave < 0 || keys_right[i].octave >= L.nlevels) FAIL(ORBX_E_INVALID, "octave out of range");
    // both extractors' streams must have finished writing their pyramids
    CU_TRY(cudaStreamSynchronize(L.stream)); CU_TRY(cudaStreamSynchronize(R.stream));
    const int nRows = L.h[0];
    if (nRows > STEREO_MAX_ROWS) FAIL(ORBX_E_INVALID, "image too high");
    const size_t ent_cap = stereo_row_entries(L.scale, L.nlevels, nr, nRows);
    const size_t need = pad((size_t)nl * 60) + pad((size_t)nr * 60) + 3 * pad((size_t)nl * 4) + 2 * pad((size_t)L.nlevels * 4) + pad((size_t)nl * 8 + 256) + 8192 + 64 * 256 +
                        pad((size_t)(nRows + 1) * 4) + pad(ent_cap * 4);
    if ((rc = m->arena.reserve(need)) || (rc = m->uparena.reserve(need))) return rc;
    m->arena.reset(); m->uparena.reset();
    KpM *kl, *kr; uint8_t *dl, *dr; float *sc, *isc;
    if ((rc = up(m, reinterpret_cast<const KpM*>(keys_left), (size_t)nl, kl)) || (rc = up(m, reinterpret_cast<const KpM*>(keys_right), (size_t)nr, kr)) ||
        (rc = up(m, desc_left, (size_t)nl * 32, dl)) || (rc = up(m, desc_right, (size_t)nr * 32, dr)) ||
        (rc = up(m, L.scale, (size_t)L.nlevels, sc)) || (rc = up(m, L.inv_scale, (size_t)L.nlevels, isc))) return rc;
    float* dur = m->arena.get<float>(nl); float* ddep = m->arena.get<float>(nl); int* sad = m->arena.get<int>(nl);
    int* row_off = m->arena.get<int>((size_t)nRows + 1); int* entries = m->arena.get<int>(ent_cap);
    if (!dur || !ddep || !sad || !row_off || !entries) FAIL(ORBX_E_CUDA, "matcher arena exhausted");
    if ((rc = flush_uploads(m))) return rc;
    StereoPyr PL, PR;
    for (int l = 0; l < L.nlevels; ++l) { PL.lv[l] = {L.ptr[l], L.pitch[l], L.w[l], L.h[l], 0}; PR.lv[l] = {R.ptr[l], R.pitch[l], R.w[l], R.h[l], 0}; }
    const StereoBatch single{nullptr, nullptr, 0};
    k_stereo_rows<<<1, 1024, 0, m->stream>>>(kr, nr, nRows, sc, row_off, entries, (int)ent_cap, single);
    LAUNCH_CHECK();
    k_stereo_match<<<(nl + 3) / 4, 128, 0, m->stream>>>(kl, dl, nl, kr, dr, nr, PL, PR, sc, isc, mb, mbf, dur, ddep, sad, single, row_off, entries, (int)ent_cap);
    LAUNCH_CHECK();
    k_stereo_median_cut<<<1, 1024, 0, m->stream>>>(nl, sad, dur, ddep, single);
    LAUNCH_CHECK();
    Gather g{m};
    if ((rc = g.begin((size_t)nl * 8 + 64))) return rc;
    const size_t o_ur = g.add(dur, (size_t)nl * 4), o_dep = g.add(ddep, (size_t)nl * 4);
    if ((rc = g.finish())) return rc;
    std::memcpy(u_right, g.host(o_ur), (size_t)nl * 4);
    std::memcpy(depth, g.host(o_dep), (size_t)nl * 4);
    return ORBX_OK;
}

// Frame::ComputeStereoMatches for the B stereo pairs of the two handles' last batched extract call: everything it reads (both pyramids,
// keypoints, descriptors, counts) is still on the device, so the call is two launches on the matcher's stream behind the extractors' work.
static int stereo_batch_launch(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int B, int cap, float mb, float mbf, float* d_ur, float* d_dep) {
    if (!m || !left || !right || B <= 0 || cap <= 0 || !d_ur || !d_dep) FAIL(ORBX_E_INVALID, "bad arguments");
    CU_TRY(cudaSetDevice(m->device));
    OrbxBatchInfo L, R; int rc;
    if ((rc = orbx_internal_last_batch(left, &L)) || (rc = orbx_internal_last_batch(right, &R))) return rc;
    if (L.device != m->device || R.device != m->device) FAIL(ORBX_E_INVALID, "extractors and matcher must live on the same device");
    if (L.nlevels != R.nlevels || L.B != B || R.B != B || L.cap != cap || R.cap != cap) FAIL(ORBX_E_INVALID, "left / right batches differ (levels, frames or keypoint capacity)");
    if (!m->ev_left) { CU_TRY(cudaEventCreateWithFlags(&m->ev_left, cudaEventDisableTiming)); CU_TRY(cudaEventCreateWithFlags(&m->ev_right, cudaEventDisableTiming)); }
    if (!m->ev_stereo) CU_TRY(cudaEventCreateWithFlags(&m->ev_stereo, cudaEventDisableTiming));
    CU_TRY(cudaEventRecord(m->ev_left, L.stream)); CU_TRY(cudaEventRecord(m->ev_right, R.stream));
    CU_TRY(cudaStreamWaitEvent(m->stream, m->ev_left, 0)); CU_TRY(cudaStreamWaitEvent(m->stream, m->ev_right, 0));
    const int nRows = L.h[0];
    if (nRows > STEREO_MAX_ROWS) FAIL(ORBX_E_INVALID, "image too high");
    const size_t ent_cap = stereo_row_entries(L.scale, L.nlevels, cap, nRows);
    const size_t need = pad((size_t)B * cap * 4) + pad((size_t)B * (nRows + 1) * 4) + pad((size_t)B * ent_cap * 4) + 8192;
    if ((rc = m->arena.reserve(need))) return rc;
    m->arena.reset();
    // scale tables: a persistent device copy (the call is asynchronous, so nothing may travel through the per-call pinned mirror)
    if (!m->stereo_scale) CU_TRY(cudaMalloc((void**)&m->stereo_scale, 2 * ORBX_MAX_LEVELS * sizeof(float)));
    if (m->stereo_scale_host.size() != (size_t)2 * L.nlevels || std::memcmp(m->stereo_scale_host.data(), L.scale, (size_t)L.nlevels * 4) != 0) {
        m->stereo_scale_host.assign(L.scale, L.scale + L.nlevels); m->stereo_scale_host.insert(m->stereo_scale_host.end(), L.inv_scale, L.inv_scale + L.nlevels);
        CU_TRY(cudaMemcpyAsync(m->stereo_scale, m->stereo_scale_host.data(), (size_t)2 * L.nlevels * 4, cudaMemcpyHostToDevice, m->stream));
        CU_TRY(cudaStreamSynchronize(m->stream));
    }
    const float* sc = m->stereo_scale; const float* isc = m->stereo_scale + L.nlevels;
    int* sad = m->arena.get<int>((size_t)B * cap);
    int* row_off = m->arena.get<int>((size_t)B * (nRows + 1)); int* entries = m->arena.get<int>((size_t)B * ent_cap);
    if (!sad || !row_off || !entries) FAIL(ORBX_E_CUDA, "matcher arena exhausted");
    StereoPyr PL, PR;
    for (int l = 0; l < L.nlevels; ++l) { PL.lv[l] = {L.ptr[l], L.pitch[l], L.w[l], L.h[l], L.fstride[l]}; PR.lv[l] = {R.ptr[l], R.pitch[l], R.w[l], R.h[l], R.fstride[l]}; }
    const StereoBatch SB{L.counts, R.counts, cap};
    m->mark();
    k_stereo_rows<<<B, 1024, 0, m->stream>>>(reinterpret_cast<const KpM*>(R.keys), 0, nRows, sc, row_off, entries, (int)ent_cap, SB);
    LAUNCH_CHECK();
    k_stereo_match<<<dim3((cap + 3) / 4, B), 128, 0, m->stream>>>(reinterpret_cast<const KpM*>(L.keys), L.desc, 0, reinterpret_cast<const KpM*>(R.keys), R.desc, 0, PL, PR, sc, isc, mb, mbf, d_ur, d_dep, sad, SB,
                                                                   row_off, entries, (int)ent_cap);
    LAUNCH_CHECK();
    m->mark();
    k_stereo_median_cut<<<B, 1024, 0, m->stream>>>(0, sad, d_ur, d_dep, SB);
    LAUNCH_CHECK();
    m->mark();
    // the kernels above read both extractors' pyramids, keypoints and descriptors: whatever is queued on the extractors next (their following batch) must not
    // overwrite them before these kernels are done
    CU_TRY(cudaEventRecord(m->ev_stereo, m->stream));
    CU_TRY(cudaStreamWaitEvent(L.stream, m->ev_stereo, 0)); CU_TRY(cudaStreamWaitEvent(R.stream, m->ev_stereo, 0));
    return ORBX_OK;
}
int orbx_compute_stereo_matches_batch_device(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int B, int cap, float mb, float mbf, float* d_u_right, float* d_depth) {
    return stereo_batch_launch(m, left, right, B, cap, mb, mbf, d_u_right, d_depth);
}
int orbx_compute_stereo_matches_batch(orbx_matcher* m, orbx_extractor* left, orbx_extractor* right, int B, int cap, float mb, float mbf, float* u_right, float* depth) {
    if (!m || !u_right || !depth || B <= 0 || cap <= 0) FAIL(ORBX_E_INVALID, "bad arguments");
    CU_TRY(cudaSetDevice(m->device));
    const size_t n = (size_t)B * cap;
    if (m->stereo_out_n < 2 * n) {
        if (m->stereo_out) cudaFree(m->stereo_out);
        m->stereo_out = nullptr; m->stereo_out_n = 0;
        CU_TRY(cudaMalloc((void**)&m->stereo_out, 2 * n * sizeof(float)));
        m->stereo_out_n = 2 * n;
    }
    int rc = stereo_batch_launch(m, left, right, B, cap, mb, mbf, m->stereo_out, m->stereo_out + n); if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(u_right, m->stereo_out, n * 4, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(cudaMemcpyAsync(depth, m->stereo_out + n, n * 4, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(cudaStreamSynchronize(m->stream));
    return ORBX_OK;
}

// ---- public windowed matchers: host frame views and device-resident frames share one implementation ----
int orbx_search_for_initialization(orbx_matcher* m, const orbx_frame_view* F1, const orbx_frame_view* F2, float* prev_matched_xy, int* matches12,
                                   int window_size, int* nmatches) {
    if (!F1 || !F2) FAIL(ORBX_E_INVALID, "bad frame view");
    return search_init_impl(m, FrameArg{F1, nullptr}, FrameArg{F2, nullptr}, prev_matched_xy, matches12, window_size, nmatches);
}
int orbx_search_for_initialization_frames(orbx_matcher* m, const orbx_frame* F1, const orbx_frame* F2, float* prev_matched_xy, int* matches12,
                                          int window_size, int* nmatches) {
    if (!F1 || !F2) FAIL(ORBX_E_INVALID, "null frame");
    return search_init_impl(m, FrameArg{nullptr, F1}, FrameArg{nullptr, F2}, prev_matched_xy, matches12, window_size, nmatches);
}
int orbx_search_for_initialization_batch(orbx_matcher* m, int n_pairs, const orbx_frame_view* F1, const orbx_frame_view* F2, float* const* prev_matched_xy,
                                         int* const* matches12, int window_size, int* nmatches) {
    if (n_pairs < 0 || (n_pairs && (!F1 || !F2))) FAIL(ORBX_E_INVALID, "bad frame views");
    std::vector<FrameArg> a1((size_t)n_pairs), a2((size_t)n_pairs);
    for (int p = 0; p < n_pairs; ++p) { a1[p] = FrameArg{F1 + p, nullptr}; a2[p] = FrameArg{F2 + p, nullptr}; }
    if (m) m->uparena.defer = n_pairs >= 8;                          // many frames per call: stage them with several threads at flush time
    const int rc = search_init_batch_impl(m, n_pairs, a1.data(), a2.data(), prev_matched_xy, matches12, window_size, nmatches);
    if (m) m->uparena.defer = false;
    return rc;
}
int orbx_search_for_initialization_frames_batch(orbx_matcher* m, int n_pairs, const orbx_frame* const* F1, const orbx_frame* const* F2, float* const* prev_matched_xy,
                                                int* const* matches12, int window_size, int* nmatches) {
    if (n_pairs < 0 || (n_pairs && (!F1 || !F2))) FAIL(ORBX_E_INVALID, "null frames");
    std::vector<FrameArg> a1((size_t)n_pairs), a2((size_t)n_pairs);
    for (int p = 0; p < n_pairs; ++p) { if (!F1[p] || !F2[p]) FAIL(ORBX_E_INVALID, "null frame"); a1[p] = FrameArg{nullptr, F1[p]}; a2[p] = FrameArg{nullptr, F2[p]}; }
    return search_init_batch_impl(m, n_pairs, a1.data(), a2.data(), prev_matched_xy, matches12, window_size, nmatches);
}
int orbx_search_by_projection_frame(orbx_matcher* m, const orbx_frame_view* cur, int n_last, const float* proj_uv, const float* proj_invz,
                                    const int* last_octave, const float* last_angle, const uint8_t* mp_desc, const uint8_t* valid, const uint8_t* mp_observed,
                                    const uint8_t* cur_occupied, float th, int forward, int backward, float mbf, int* cur_match, int* nmatches) {
    if (!cur) FAIL(ORBX_E_INVALID, "bad frame view");
    return search_proj_frame_impl(m, FrameArg{cur, nullptr}, n_last, proj_uv, proj_invz, last_octave, last_angle, mp_desc, valid, mp_observed, cur_occupied, th, forward, backward, mbf, cur_match, nmatches);
}
int orbx_search_by_projection_frame_dev(orbx_matcher* m, const orbx_frame* cur, int n_last, const float* proj_uv, const float* proj_invz,
                                        const int* last_octave, const float* last_angle, const uint8_t* mp_desc, const uint8_t* valid, const uint8_t* mp_observed,
                                        const uint8_t* cur_occupied, float th, int forward, int backward, float mbf, int* cur_match, int* nmatches) {
    if (!cur) FAIL(ORBX_E_INVALID, "null frame");
    return search_proj_frame_impl(m, FrameArg{nullptr, cur}, n_last, proj_uv, proj_invz, last_octave, last_angle, mp_desc, valid, mp_observed, cur_occupied, th, forward, backward, mbf, cur_match, nmatches);
}
int orbx_search_by_projection_frame_pose(orbx_matcher* m, const orbx_frame_view* cur, const orbx_frame* cur_dev, int n_last, const float* world_xyz, const uint8_t* has_point,
                                         const float* Rcw, const float* tcw, float fx, float fy, float cx, float cy, const int* last_octave, const float* last_angle,
                                         const uint8_t* mp_desc, const uint8_t* mp_observed, const uint8_t* cur_occupied, float th, int forward, int backward, float mbf,
                                         int* cur_match, int* nmatches, float* proj_uv_out, float* proj_invz_out, uint8_t* valid_out) {
    if ((cur == nullptr) == (cur_dev == nullptr)) FAIL(ORBX_E_INVALID, "exactly one of the frame view and the device frame must be given");
    if (!Rcw || !tcw) FAIL(ORBX_E_INVALID, "null pose");
    ProjPoseArgs a; a.world = world_xyz; a.has = has_point; a.uv_out = proj_uv_out; a.invz_out = proj_invz_out; a.valid_out = valid_out;
    for (int i = 0; i < 9; ++i) a.P.R[i] = Rcw[i];
    for (int i = 0; i < 3; ++i) a.P.t[i] = tcw[i];
    a.P.fx = fx; a.P.fy = fy; a.P.cx = cx; a.P.cy = cy;
    return search_proj_frame_impl(m, FrameArg{cur, cur_dev}, n_last, nullptr, nullptr, last_octave, last_angle, mp_desc, nullptr, mp_observed, cur_occupied, th, forward, backward, mbf,
                                  cur_match, nmatches, M_TH_HIGH, 0, -1, &a);
}
// SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist)  (ORBmatcher.cc:1731-1863) is the Frame x Frame search with the
// level predicted from the distance, every claim blocking (mvpMapPoints[i2] != NULL), no uRight test, and ORBdist as the acceptance bound
static int search_proj_keyframe(orbx_matcher* m, const FrameArg cur, int n_kf, const float* proj_uv, const int* predicted_level, const float* kf_angle, const uint8_t* mp_desc,
                                const uint8_t* valid, const uint8_t* cur_occupied, float th, int orb_dist, int* cur_match, int* nmatches) {
    if (n_kf < 0 || n_kf >= (1 << 20)) FAIL(ORBX_E_INVALID, "bad arguments");
    std::vector<float> invz((size_t)std::max(n_kf, 1), 0.f);                     // only the sign test of the Frame x Frame form reads it
    std::vector<uint8_t> blocks((size_t)std::max(n_kf, 1), 1);                  // a feature claimed in this call is never re-assigned (:1808)
    return search_proj_frame_impl(m, cur, n_kf, proj_uv, invz.data(), predicted_level, kf_angle, mp_desc, valid, blocks.data(), cur_occupied, th, 0, 0, 0.f, cur_match, nmatches, orb_dist, 1);
}
// SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)  (ORBmatcher.cc:388-512, loop closing): the same search on a KeyFrame's
// features (KeyFrame::GetFeaturesInArea walks the same grid in the same order, src/KeyFrame.cc:752-797) with levels
// [predicted - 1, predicted], TH_LOW as acceptance bound, no rotation histogram
static int search_proj_kf_points(orbx_matcher* m, const FrameArg kf, int n_points, const float* proj_uv, const int* predicted_level, const uint8_t* mp_desc, const uint8_t* valid,
                                 const uint8_t* kf_matched, float th, int* kf_match, int* nmatches) {
    if (n_points < 0 || n_points >= (1 << 20)) FAIL(ORBX_E_INVALID, "bad arguments");
    const size_t nn = (size_t)std::max(n_points, 1);
    std::vector<float> zeros(nn, 0.f); std::vector<uint8_t> blocks(nn, 1);
    return search_proj_frame_impl(m, kf, n_points, proj_uv, zeros.data(), predicted_level, zeros.data(), mp_desc, valid, blocks.data(), kf_matched, th, 2, 0, 0.f, kf_match, nmatches,
                                  M_TH_LOW, 1, 0);
}
int orbx_search_by_projection_keyframe_points(orbx_matcher* m, const orbx_frame_view* kf, int n_points, const float* proj_uv, const int* predicted_level, const uint8_t* mp_desc,
                                              const uint8_t* valid, const uint8_t* kf_matched, float th, int* kf_match, int* nmatches) {
    if (!kf) FAIL(ORBX_E_INVALID, "bad frame view");
    return search_proj_kf_points(m, FrameArg{kf, nullptr}, n_points, proj_uv, predicted_level, mp_desc, valid, kf_matched, th, kf_match, nmatches);
}
int orbx_search_by_projection_keyframe_points_dev(orbx_matcher* m, const orbx_frame* kf, int n_points, const float* proj_uv, const int* predicted_level, const uint8_t* mp_desc,
                                                  const uint8_t* valid, const uint8_t* kf_matched, float th, int* kf_match, int* nmatches) {
    if (!kf) FAIL(ORBX_E_INVALID, "null frame");
    return search_proj_kf_points(m, FrameArg{nullptr, kf}, n_points, proj_uv, predicted_level, mp_desc, valid, kf_matched, th, kf_match, nmatches);
}
// SearchBySim3 (ORBmatcher.cc:1290-1555): two order-independent passes (map points of one KeyFrame against the features of the other, levels
// [predicted - 1, predicted], best distance <= TH_HIGH, no exclusions) and the mutual-consistency check, all on the device
struct Sim3Side { int n; const float* proj_uv; const int* predicted_level; const uint8_t* mp_desc; const uint8_t* valid; };
// independent best match of every query on level predicted - 1 or predicted within max_dist; proj_ur / inv_sigma2 != NULL adds Fuse's chi-square gate
static int sim3_pass(orbx_matcher* m, const FrameDev& target, const Sim3Side& q, float th, int* d_best, int*& d_total, int max_dist = M_TH_HIGH,
                     const float* proj_ur = nullptr, const float* inv_sigma2 = nullptr, int nlevels = 0) {
    int rc;
    float *uv, *iz, *dur = nullptr, *dis = nullptr; int* lv; uint8_t *dd, *va;
    std::vector<float> zeros((size_t)std::max(q.n, 1), 0.f);
    if ((rc = up(m, q.proj_uv, (size_t)q.n * 2, uv)) || (rc = up(m, zeros.data(), (size_t)q.n, iz)) || (rc = up(m, q.predicted_level, (size_t)q.n, lv)) ||
        (rc = up(m, q.mp_desc, (size_t)q.n * 32, dd)) || (rc = up(m, q.valid, (size_t)q.n, va))) return rc;
    if (proj_ur && ((rc = up(m, proj_ur, (size_t)q.n, dur)) || (rc = up(m, inv_sigma2, (size_t)nlevels, dis)))) return rc;
    if ((rc = flush_uploads(m))) return rc;
    QueryParams P; std::memset(&P, 0, sizeof(P));
    P.mode = MODE_PROJ_FRAME; P.nq = q.n; P.q_desc = dd; P.q_xy = uv; P.q_invz = iz; P.q_octave = lv; P.q_valid = va; P.th = th; P.forward = 2; P.no_ur = 1;
    if (proj_ur) { P.chi2 = 1; P.q_ur = dur; P.q_invsigma2 = dis; }
    int *counts, *offsets; uint32_t* cand; uint2* pre;
    if ((rc = window_search(m, P, target, counts, offsets, cand, pre))) return rc;
    if (q.n) { k_best_extract<<<(q.n + 127) / 128, 128, 0, m->stream>>>(q.n, offsets, cand, pre, (int)std::min<size_t>(m->cand_cap, 0x7FFFFFFF), max_dist, d_best); LAUNCH_CHECK(); }
    d_total = offsets + q.n;
    return ORBX_OK;
}

// The search of both ORBmatcher::Fuse forms (ORBmatcher.cc:1020-1175, 1179-1310): per map point, the best KeyFrame feature on level
// predicted - 1 or predicted within TH_LOW; proj_ur != NULL: the pose form with its chi-square reprojection gate (stereo 7.8, mono 5.99)
int orbx_fuse_search(orbx_matcher* m, const orbx_frame_view* kf, int n_points, const float* proj_uv, const float* proj_ur, const int* predicted_level,
                     const uint8_t* mp_desc, const uint8_t* valid, const float* inv_level_sigma2, float th, int* best_idx) {
    if (!m) FAIL(ORBX_E_INVALID, "null argument");
    int rc;
    if ((rc = check_frame(kf))) return rc;
    if (n_points < 0 || n_points >= (1 << 20) || (n_points && (!proj_uv || !predicted_level || !mp_desc || !valid || !best_idx)) || (proj_ur && !inv_level_sigma2))
        FAIL(ORBX_E_INVALID, "bad arguments");
    for (int i = 0; i < n_points; ++i) { best_idx[i] = -1; if (valid[i] && (predicted_level[i] < 0 || predicted_level[i] >= kf->nlevels)) FAIL(ORBX_E_INVALID, "predicted level out of range"); }
    for (int j = 0; j < kf->n; ++j) if (kf->keys_un[j].octave < 0 || kf->keys_un[j].octave >= kf->nlevels) FAIL(ORBX_E_INVALID, "octave out of range");
    if (n_points == 0 || kf->n == 0) return ORBX_OK;
    CU_TRY(cudaSetDevice(m->device));
    const size_t need = frame_bytes(kf) + pad((size_t)n_points * 52) + 8 * pad((size_t)(n_points + kf->n + 2) * 8) + 16384;
    if ((rc = m->arena.reserve(need)) || (rc = m->uparena.reserve(need))) return rc;
retry:
    m->arena.reset(); m->uparena.reset();
    FrameDev d; uint32_t* sk;
    if ((rc = upload_frame(m, kf, d, sk))) return rc;
    if ((rc = flush_uploads(m)) || (rc = build_grid(m, d, sk))) return rc;
    int* res = m->arena.get<int>((size_t)n_points + 2);
    if (!res) FAIL(ORBX_E_CUDA, "matcher arena exhausted");
    int* tot;
    Sim3Side q{n_points, proj_uv, predicted_level, mp_desc, valid};
    if ((rc = sim3_pass(m, d, q, th, res + 1, tot, M_TH_LOW, proj_ur, inv_level_sigma2, kf->nlevels))) return rc;
    CU_TRY(cudaMemcpyAsync(res, tot, 4, cudaMemcpyDeviceToDevice, m->stream));
    const size_t res_bytes = ((size_t)n_points + 1) * 4;
    if ((rc = m->ensure_download(res_bytes))) return rc;
    CU_TRY(cudaMemcpyAsync(m->dl_host, res, res_bytes, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(cudaStreamSynchronize(m->stream));
    const int* hr = reinterpret_cast<const int*>(m->dl_host);
    CAND_RETRY(hr[0]);
    std::memcpy(best_idx, hr + 1, (size_t)n_points * 4);
    return ORBX_OK;
}
int orbx_search_by_sim3(orbx_matcher* m, const orbx_frame_view* kf1, const orbx_frame_view* kf2, const float* proj_uv1, const int* predicted_level1, const uint8_t* mp_desc1,
                        const uint8_t* valid1, const float* proj_uv2, const int* predicted_level2, const uint8_t* mp_desc2, const uint8_t* valid2, float th,
                        int* match12, int* nfound) {
    if (!m || !nfound) FAIL(ORBX_E_INVALID, "null argument");
    int rc;
    if ((rc = check_frame(kf1)) || (rc = check_frame(kf2))) return rc;
    const int n1 = kf1->n, n2 = kf2->n;
    if ((n1 && (!proj_uv1 || !predicted_level1 || !mp_desc1 || !valid1 || !match12)) || (n2 && (!proj_uv2 || !predicted_level2 || !mp_desc2 || !valid2))) FAIL(ORBX_E_INVALID, "null buffer");
    for (int i = 0; i < n1; ++i) if (valid1[i] && (predicted_level1[i] < 0 || predicted_level1[i] >= kf2->nlevels)) FAIL(ORBX_E_INVALID, "predicted level out of range");
    for (int i = 0; i < n2; ++i) if (valid2[i] && (predicted_level2[i] < 0 || predicted_level2[i] >= kf1->nlevels)) FAIL(ORBX_E_INVALID, "predicted level out of range");
    *nfound = 0;
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    if (n1 == 0 || n2 == 0) return ORBX_OK;
    CU_TRY(cudaSetDevice(m->device));
    const size_t need = frame_bytes(kf1) + frame_bytes(kf2) + 2 * (pad((size_t)(n1 + n2) * 44) + 8 * pad((size_t)(n1 + n2 + 2) * 8)) + 16384;
    if ((rc = m->arena.reserve(need)) || (rc = m->uparena.reserve(need))) return rc;
retry:
    m->arena.reset(); m->uparena.reset();
    FrameDev d1, d2; uint32_t *sk1, *sk2;
    if ((rc = upload_frame(m, kf1, d1, sk1)) || (rc = upload_frame(m, kf2, d2, sk2))) return rc;
    if ((rc = flush_uploads(m)) || (rc = build_grid(m, d1, sk1)) || (rc = build_grid(m, d2, sk2))) return rc;
    // the upload arena keeps growing inside one call: later flushes resend the earlier bytes unchanged, which is harmless
    int* best1 = m->arena.get<int>(n1 + 1); int* best2 = m->arena.get<int>(n2 + 1); int* res = m->arena.get<int>((size_t)n1 + 4);
    if (!best1 || !best2 || !res) FAIL(ORBX_E_CUDA, "matcher arena exhausted");
    int *tot1, *tot2;
    Sim3Side q1{n1, proj_uv1, predicted_level1, mp_desc1, valid1}, q2{n2, proj_uv2, predicted_level2, mp_desc2, valid2};
    if ((rc = sim3_pass(m, d2, q1, th, best1, tot1))) return rc;                 // map points of KF1 into KF2 (:1367-1440)
    CU_TRY(cudaMemcpyAsync(res + 1, tot1, 4, cudaMemcpyDeviceToDevice, m->stream));
    if ((rc = sim3_pass(m, d1, q2, th, best2, tot2))) return rc;                 // map points of KF2 into KF1 (:1443-1530)
    CU_TRY(cudaMemcpyAsync(res + 2, tot2, 4, cudaMemcpyDeviceToDevice, m->stream));
    CU_TRY(cudaMemsetAsync(res, 0, 4, m->stream));
    k_mutual_check<<<(n1 + 127) / 128, 128, 0, m->stream>>>(n1, best1, best2, res + 3, res);
    LAUNCH_CHECK();
    const size_t res_bytes = ((size_t)n1 + 3) * 4;
    if ((rc = m->ensure_download(res_bytes))) return rc;
    CU_TRY(cudaMemcpyAsync(m->dl_host, res, res_bytes, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(cudaStreamSynchronize(m->stream));
    const int* hr = reinterpret_cast<const int*>(m->dl_host);
    CAND_RETRY(std::max(hr[1], hr[2]));
    *nfound = hr[0];
    std::memcpy(match12, hr + 3, (size_t)n1 * 4);
    return ORBX_OK;
}

int orbx_search_by_projection_keyframe(orbx_matcher* m, const orbx_frame_view* cur, int n_kf, const float* proj_uv, const int* predicted_level, const float* kf_angle,
                                       const uint8_t* mp_desc, const uint8_t* valid, const uint8_t* cur_occupied, float th, int orb_dist, int* cur_match, int* nmatches) {
    if (!cur) FAIL(ORBX_E_INVALID, "bad frame view");
    return search_proj_keyframe(m, FrameArg{cur, nullptr}, n_kf, proj_uv, predicted_level, kf_angle, mp_desc, valid, cur_occupied, th, orb_dist, cur_match, nmatches);
}
int orbx_search_by_projection_keyframe_dev(orbx_matcher* m, const orbx_frame* cur, int n_kf, const float* proj_uv, const int* predicted_level, const float* kf_angle,
                                           const uint8_t* mp_desc, const uint8_t* valid, const uint8_t* cur_occupied, float th, int orb_dist, int* cur_match, int* nmatches) {
    if (!cur) FAIL(ORBX_E_INVALID, "null frame");
    return search_proj_keyframe(m, FrameArg{nullptr, cur}, n_kf, proj_uv, predicted_level, kf_angle, mp_desc, valid, cur_occupied, th, orb_dist, cur_match, nmatches);
}
int orbx_search_by_projection_points(orbx_matcher* m, const orbx_frame_view* F, int n_points, const float* track_uv, const float* track_ur, const int* track_level,
                                     const float* track_view_cos, const uint8_t* mp_desc, const uint8_t* mp_observed, const uint8_t* f_occupied, float th,
                                     int* f_match, int* nmatches) {
    if (!F) FAIL(ORBX_E_INVALID, "bad frame view");
    return search_proj_points_impl(m, FrameArg{F, nullptr}, n_points, track_uv, track_ur, track_level, track_view_cos, mp_desc, mp_observed, f_occupied, th, f_match, nmatches);
}
int orbx_search_by_projection_points_dev(orbx_matcher* m, const orbx_frame* F, int n_points, const float* track_uv, const float* track_ur, const int* track_level,
                                         const float* track_view_cos, const uint8_t* mp_desc, const uint8_t* mp_observed, const uint8_t* f_occupied, float th,
                                         int* f_match, int* nmatches) {
    if (!F) FAIL(ORBX_E_INVALID, "null frame");
    return search_proj_points_impl(m, FrameArg{nullptr, F}, n_points, track_uv, track_ur, track_level, track_view_cos, mp_desc, mp_observed, f_occupied, th, f_match, nmatches);
}

// ---- device-resident Frame ----
#define FLAUNCH_CHECK() do { cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) { \
    orbx_set_error(std::string("kernel launch: ") + cudaGetErrorString(_e)); return ORBX_E_CUDA; } ++f->launches; } while (0)

int orbx_frame_create(int device, orbx_frame** out) {
    if (!out) FAIL(ORBX_E_INVALID, "null out");
    *out = nullptr;
    int ndev = 0;
    CU_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) FAIL(ORBX_E_CUDA, "no such CUDA device (this library has no CPU fallback)");
    CU_TRY(cudaSetDevice(device));
    orbx_frame* f = new orbx_frame();
    f->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete f; FAIL(ORBX_E_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); }
    std::memset(&f->dev, 0, sizeof(f->dev));
    *out = f;
    return ORBX_OK;
}
void orbx_frame_destroy(orbx_frame* f) {
    if (!f) return;
    cudaSetDevice(f->device); cudaStreamSynchronize(f->stream);
    f->keys_in.release(); f->keys_un.release(); f->desc.release(); f->u_right.release(); f->depth.release(); f->scale.release();
    f->depth_img.release(); f->bounds4.release(); f->cell_start.release(); f->entries.release(); f->sort_keys.release();
    cudaStreamDestroy(f->stream);
    delete f;
}

static CamDev make_cam(const orbx_camera& c) {
    CamDev d;
    d.fx = (double)c.fx; d.fy = (double)c.fy; d.cx = (double)c.cx; d.cy = (double)c.cy;
    d.ifx = 1.0 / d.fx; d.ify = 1.0 / d.fy;
    d.k1 = (double)c.k1; d.k2 = (double)c.k2; d.p1 = (double)c.p1; d.p2 = (double)c.p2; d.k3 = (double)c.k3;
    d.distorted = (c.k1 != 0.0f); d.bf = c.bf;
    return d;
}
static int check_cam(const orbx_camera* cam) {
    if (!cam) FAIL(ORBX_E_INVALID, "null camera");
    if (!(cam->fx != 0.f) || !(cam->fy != 0.f)) FAIL(ORBX_E_INVALID, "camera focal length is zero");
    return ORBX_OK;
}

// Every step below is asynchronous on the frame's stream; the public calls that hand results to the host (or make the frame
// visible to the matchers) synchronise.

// mvKeys / mDescriptors: device source (an extractor's result) or host arrays
static int frame_take(orbx_frame* f, const KpM* d_keys, const uint8_t* d_desc, const orbx_keypoint* h_keys, const uint8_t* h_desc, int n, int nlevels, const float* scale_host) {
    if (nlevels <= 0 || nlevels > ORBX_MAX_LEVELS || !scale_host) FAIL(ORBX_E_INVALID, "bad pyramid description");
    cudaStream_t s = f->stream;
    f->state = 0; f->n = -1;
    const size_t nn = n ? n : 1;
    int rc;
    if ((rc = f->keys_in.ensure(nn)) || (rc = f->keys_un.ensure(nn)) || (rc = f->desc.ensure(nn * 32)) || (rc = f->u_right.ensure(nn)) || (rc = f->depth.ensure(nn)) ||
        (rc = f->scale.ensure(ORBX_MAX_LEVELS)) || (rc = f->cell_start.ensure(GRID_CELLS + 1)) || (rc = f->entries.ensure(nn)) || (rc = f->sort_keys.ensure(nn))) return rc;
    if (n) {
        const cudaMemcpyKind kind = d_keys ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        CU_TRY(cudaMemcpyAsync(f->keys_in.p, d_keys ? (const void*)d_keys : (const void*)h_keys, (size_t)n * 28, kind, s));
        CU_TRY(cudaMemcpyAsync(f->desc.p, d_keys ? (const void*)d_desc : (const void*)h_desc, (size_t)n * 32, kind, s));
    }
    CU_TRY(cudaMemcpyAsync(f->scale.p, scale_host, (size_t)nlevels * 4, cudaMemcpyHostToDevice, s));
    f->n_pending = n; f->nlevels = nlevels; f->state = FRAME_HAS_KEYS;
    return ORBX_OK;
}

// depth argument of the RGB-D step -> device pointer, mode, pitch
static int frame_stage_depth(orbx_frame* f, const float* depth, size_t depth_stride, int rows, int cols, const float*& d_depth, int& mode, int& pitch) {
    d_depth = nullptr; mode = 0; pitch = 0;
    const int n = f->n_pending;
    if (!depth || !n) return ORBX_OK;
    int rc;
    if (depth_stride) {
        if (rows <= 0 || cols <= 0 || depth_stride % 4 || depth_stride < (size_t)cols * 4) FAIL(ORBX_E_INVALID, "bad depth image geometry");
        if ((rc = f->depth_img.ensure((size_t)rows * cols))) return rc;
        CU_TRY(cudaMemcpy2DAsync(f->depth_img.p, (size_t)cols * 4, depth, depth_stride, (size_t)cols * 4, rows, cudaMemcpyHostToDevice, f->stream));
        mode = 1; pitch = cols;
    } else {
        if ((rc = f->depth_img.ensure(n))) return rc;
        CU_TRY(cudaMemcpyAsync(f->depth_img.p, depth, (size_t)n * 4, cudaMemcpyHostToDevice, f->stream));
        mode = 2;
    }
    d_depth = f->depth_img.p;
    return ORBX_OK;
}

// UndistortKeyPoints (+ ComputeStereoFromRGBD when with_stereo): one kernel
static int frame_undistort(orbx_frame* f, const CamDev& cd, bool with_stereo, const float* d_depth, int mode, int pitch, int rows, int cols) {
    if (!(f->state & FRAME_HAS_KEYS)) FAIL(ORBX_E_STATE, "device frame holds no keypoints (orbx_frame_take first)");
    const int n = f->n_pending;
    if (n) {
        k_frame_undistort_stereo<<<(n + 127) / 128, 128, 0, f->stream>>>(f->keys_in.p, n, cd, d_depth, with_stereo ? mode : 0, pitch, rows, cols, f->keys_un.p, f->u_right.p, f->depth.p);
        FLAUNCH_CHECK();
    }
    f->state |= FRAME_HAS_UN;
    return ORBX_OK;
}

// ComputeStereoFromRGBD on its own (mvKeysUn already on the device)
static int frame_rgbd(orbx_frame* f, float bf, const float* d_depth, int mode, int pitch, int rows, int cols) {
    if (!(f->state & FRAME_HAS_UN)) FAIL(ORBX_E_STATE, "ComputeStereoFromRGBD needs the undistorted keypoints (orbx_frame_undistort_keypoints first)");
    const int n = f->n_pending;
    if (n) {
        k_frame_rgbd<<<(n + 127) / 128, 128, 0, f->stream>>>(f->keys_in.p, f->keys_un.p, n, bf, d_depth, mode, pitch, rows, cols, f->u_right.p, f->depth.p);
        FLAUNCH_CHECK();
    }
    return ORBX_OK;
}

// AssignFeaturesToGrid with the given image bounds; publishes the frame to the matchers
static int frame_grid(orbx_frame* f, float min_x, float max_x, float min_y, float max_y, float gw_inv, float gh_inv) {
    if (!(f->state & FRAME_HAS_UN)) FAIL(ORBX_E_STATE, "AssignFeaturesToGrid needs the undistorted keypoints (orbx_frame_undistort_keypoints first)");
    const int n = f->n_pending;
    f->min_x = min_x; f->max_x = max_x; f->min_y = min_y; f->max_y = max_y; f->gw_inv = gw_inv; f->gh_inv = gh_inv;
    k_grid_build<<<1, 1024, 0, f->stream>>>(f->keys_un.p, n, min_x, min_y, gw_inv, gh_inv, f->sort_keys.p, f->entries.p, f->cell_start.p);
    FLAUNCH_CHECK();
    CU_TRY(cudaStreamSynchronize(f->stream));
    f->n = n; f->state |= FRAME_HAS_GRID;
    FrameDev& d = f->dev;
    d.n = n; d.keys = f->keys_un.p; d.desc = f->desc.p; d.u_right = f->u_right.p; d.min_x = min_x; d.min_y = min_y; d.max_x = max_x; d.max_y = max_y;
    d.gw_inv = gw_inv; d.gh_inv = gh_inv; d.scale = f->scale.p; d.nlevels = f->nlevels; d.cell_start = f->cell_start.p; d.entries = f->entries.p;
    return ORBX_OK;
}

// ComputeImageBounds (Frame.cc:1120-1176) + grid constants (:301-302), cached like Frame::mbInitialComputations
static int frame_bounds(orbx_frame* f, const orbx_camera* cam, const CamDev& cd, int rows, int cols) {
    if (rows <= 0 || cols <= 0) FAIL(ORBX_E_INVALID, "bad image size");
    if (f->bounds_valid && !std::memcmp(&f->bcam, cam, sizeof(orbx_camera)) && f->brows == rows && f->bcols == cols) return ORBX_OK;
    if (cd.distorted) {
        int rc;
        if ((rc = f->bounds4.ensure(4))) return rc;
        k_frame_bounds<<<1, 32, 0, f->stream>>>(cd, rows, cols, f->bounds4.p);
        FLAUNCH_CHECK();
        float b[4];
        CU_TRY(cudaMemcpyAsync(b, f->bounds4.p, 16, cudaMemcpyDeviceToHost, f->stream));
        CU_TRY(cudaStreamSynchronize(f->stream));
        f->cmin_x = b[0]; f->cmax_x = b[1]; f->cmin_y = b[2]; f->cmax_y = b[3];
    } else { f->cmin_x = 0.f; f->cmax_x = (float)cols; f->cmin_y = 0.f; f->cmax_y = (float)rows; }
    f->bcam = *cam; f->brows = rows; f->bcols = cols; f->bounds_valid = true;
    return ORBX_OK;
}

static int frame_build(orbx_frame* f, const orbx_camera* cam, int rows, int cols, const float* depth, size_t depth_stride) {
    int rc;
    if ((rc = check_cam(cam))) return rc;
    const CamDev cd = make_cam(*cam);
    if ((rc = frame_bounds(f, cam, cd, rows, cols))) return rc;
    const float* d_depth; int mode, pitch;
    if ((rc = frame_stage_depth(f, depth, depth_stride, rows, cols, d_depth, mode, pitch))) return rc;
    if ((rc = frame_undistort(f, cd, true, d_depth, mode, pitch, rows, cols))) return rc;
    return frame_grid(f, f->cmin_x, f->cmax_x, f->cmin_y, f->cmax_y, (float)GRID_COLS / (f->cmax_x - f->cmin_x), (float)GRID_ROWS / (f->cmax_y - f->cmin_y));
}

static int take_from_extractor(orbx_frame* f, orbx_extractor* h) {
    if (!f || !h) FAIL(ORBX_E_INVALID, "null handle");
    f->n = -1; f->state = 0;
    CU_TRY(cudaSetDevice(f->device));
    OrbxLastResult r; int rc;
    if ((rc = orbx_internal_last_result(h, &r))) return rc;
    if (r.device != f->device) FAIL(ORBX_E_INVALID, "extractor and frame must live on the same device");
    if (r.n >= (1 << 20)) FAIL(ORBX_E_INVALID, "too many keypoints");
    CU_TRY(cudaStreamSynchronize(r.stream));                    // the extractor call has returned, so this is a no-op guard
    return frame_take(f, reinterpret_cast<const KpM*>(r.keys), r.desc, nullptr, nullptr, r.n, r.nlevels, r.scale);
}
static int take_from_host(orbx_frame* f, const orbx_keypoint* keys, const uint8_t* descriptors, int n, int nlevels, const float* scale_factors) {
    if (!f || n < 0 || n >= (1 << 20) || (n && (!keys || !descriptors))) FAIL(ORBX_E_INVALID, "bad arguments");
    f->n = -1; f->state = 0;
    CU_TRY(cudaSetDevice(f->device));
    return frame_take(f, nullptr, nullptr, keys, descriptors, n, nlevels, scale_factors);
}

int orbx_frame_assign(orbx_frame* f, orbx_extractor* h, const orbx_camera* cam, int img_rows, int img_cols, const float* depth, size_t depth_stride_bytes) {
    int rc = take_from_extractor(f, h);
    return rc ? rc : frame_build(f, cam, img_rows, img_cols, depth, depth_stride_bytes);
}
int orbx_frame_assign_host(orbx_frame* f, const orbx_keypoint* keys, const uint8_t* descriptors, int n, int nlevels, const float* scale_factors,
                           const orbx_camera* cam, int img_rows, int img_cols, const float* depth, size_t depth_stride_bytes) {
    int rc = take_from_host(f, keys, descriptors, n, nlevels, scale_factors);
    return rc ? rc : frame_build(f, cam, img_rows, img_cols, depth, depth_stride_bytes);
}

// ---- the same steps one by one, as the reference's Frame constructors call them ----
int orbx_frame_take(orbx_frame* f, orbx_extractor* h) { return take_from_extractor(f, h); }
int orbx_frame_take_host(orbx_frame* f, const orbx_keypoint* keys, const uint8_t* descriptors, int n, int nlevels, const float* scale_factors) {
    return take_from_host(f, keys, descriptors, n, nlevels, scale_factors);
}
int orbx_frame_undistort_keypoints(orbx_frame* f, const orbx_camera* cam, orbx_keypoint* keys_un_out) {
    if (!f) FAIL(ORBX_E_INVALID, "null handle");
    int rc;
    if ((rc = check_cam(cam))) return rc;
    CU_TRY(cudaSetDevice(f->device));
    if ((rc = frame_undistort(f, make_cam(*cam), false, nullptr, 0, 0, 0, 0))) return rc;
    if (keys_un_out && f->n_pending) CU_TRY(cudaMemcpyAsync(keys_un_out, f->keys_un.p, (size_t)f->n_pending * 28, cudaMemcpyDeviceToHost, f->stream));
    CU_TRY(cudaStreamSynchronize(f->stream));
    return ORBX_OK;
}
int orbx_frame_compute_stereo_from_rgbd(orbx_frame* f, float bf, const float* depth, size_t depth_stride_bytes, int img_rows, int img_cols, float* u_right_out, float* depth_out) {
    if (!f || !depth) FAIL(ORBX_E_INVALID, "null argument");
    if (!(f->state & FRAME_HAS_UN)) FAIL(ORBX_E_STATE, "ComputeStereoFromRGBD needs the undistorted keypoints (orbx_frame_undistort_keypoints first)");
    CU_TRY(cudaSetDevice(f->device));
    int rc; const float* d_depth; int mode, pitch;
    if ((rc = frame_stage_depth(f, depth, depth_stride_bytes, img_rows, img_cols, d_depth, mode, pitch))) return rc;
    if ((rc = frame_rgbd(f, bf, d_depth, mode, pitch, img_rows, img_cols))) return rc;
    const size_t n = (size_t)f->n_pending;
    if (u_right_out && n) CU_TRY(cudaMemcpyAsync(u_right_out, f->u_right.p, n * 4, cudaMemcpyDeviceToHost, f->stream));
    if (depth_out && n) CU_TRY(cudaMemcpyAsync(depth_out, f->depth.p, n * 4, cudaMemcpyDeviceToHost, f->stream));
    CU_TRY(cudaStreamSynchronize(f->stream));
    return ORBX_OK;
}
int orbx_frame_assign_features_to_grid(orbx_frame* f, const float* bounds6, int* cell_start_out, int* entries_out) {
    if (!f || !bounds6) FAIL(ORBX_E_INVALID, "null argument");
    CU_TRY(cudaSetDevice(f->device));
    int rc;
    if ((rc = frame_grid(f, bounds6[0], bounds6[1], bounds6[2], bounds6[3], bounds6[4], bounds6[5]))) return rc;
    if (cell_start_out) return orbx_frame_grid(f, cell_start_out, entries_out);
    return ORBX_OK;
}

int orbx_frame_set_stereo(orbx_frame* f, const float* u_right, const float* depth) {
    if (!f || !(f->state & FRAME_HAS_UN)) FAIL(ORBX_E_STATE, "device frame is empty (call orbx_frame_assign first)");
    const int n = f->n_pending;
    if (n && (!u_right || !depth)) FAIL(ORBX_E_INVALID, "null buffer");
    if (n == 0) return ORBX_OK;
    CU_TRY(cudaSetDevice(f->device));
    CU_TRY(cudaMemcpyAsync(f->u_right.p, u_right, (size_t)n * 4, cudaMemcpyHostToDevice, f->stream));
    CU_TRY(cudaMemcpyAsync(f->depth.p, depth, (size_t)n * 4, cudaMemcpyHostToDevice, f->stream));
    CU_TRY(cudaStreamSynchronize(f->stream));
    return ORBX_OK;
}

int orbx_frame_size(const orbx_frame* f) { return f ? f->n : -1; }
int orbx_frame_taken(const orbx_frame* f) { return (f && (f->state & FRAME_HAS_KEYS)) ? f->n_pending : -1; }

int orbx_frame_read(orbx_frame* f, orbx_keypoint* keys_un, float* u_right, float* depth, float* bounds) {
    if (!f || f->n < 0) FAIL(ORBX_E_STATE, "device frame is empty (call orbx_frame_assign first)");
    CU_TRY(cudaSetDevice(f->device));
    const size_t n = (size_t)f->n;
    if (keys_un && n) CU_TRY(cudaMemcpyAsync(keys_un, f->keys_un.p, n * 28, cudaMemcpyDeviceToHost, f->stream));
    if (u_right && n) CU_TRY(cudaMemcpyAsync(u_right, f->u_right.p, n * 4, cudaMemcpyDeviceToHost, f->stream));
    if (depth && n) CU_TRY(cudaMemcpyAsync(depth, f->depth.p, n * 4, cudaMemcpyDeviceToHost, f->stream));
    CU_TRY(cudaStreamSynchronize(f->stream));
    if (bounds) { bounds[0] = f->min_x; bounds[1] = f->max_x; bounds[2] = f->min_y; bounds[3] = f->max_y; bounds[4] = f->gw_inv; bounds[5] = f->gh_inv; }
    return ORBX_OK;
}

int orbx_frame_grid(orbx_frame* f, int* cell_start, int* entries) {
    if (!f || f->n < 0) FAIL(ORBX_E_STATE, "device frame is empty (call orbx_frame_assign first)");
    if (!cell_start || (f->n && !entries)) FAIL(ORBX_E_INVALID, "null buffer");
    CU_TRY(cudaSetDevice(f->device));
    CU_TRY(cudaMemcpyAsync(cell_start, f->cell_start.p, (size_t)(GRID_CELLS + 1) * 4, cudaMemcpyDeviceToHost, f->stream));
    if (f->n) CU_TRY(cudaMemcpyAsync(entries, f->entries.p, (size_t)f->n * 4, cudaMemcpyDeviceToHost, f->stream));
    CU_TRY(cudaStreamSynchronize(f->stream));
    return ORBX_OK;
}

int orbx_frame_features_in_area(orbx_matcher* m, const orbx_frame* f, int nq, const float* xy, const float* r, const int* min_level, const int* max_level,
                                int* offsets_out, int* indices, int cap, int* total_out) {
    if (!m || !total_out || nq < 0 || nq >= (1 << 20) || (nq && (!xy || !r || !min_level || !max_level)) || !offsets_out || cap < 0 || (cap && !indices))
        FAIL(ORBX_E_INVALID, "bad arguments");
    int rc;
    if ((rc = check_frame_arg(m, FrameArg{nullptr, f}))) return rc;
    *total_out = 0; offsets_out[0] = 0;
    if (nq == 0) return ORBX_OK;
    CU_TRY(cudaSetDevice(m->device));
    const size_t need = 8 * pad((size_t)(nq + 2) * 8) + 8192;
    if ((rc = m->arena.reserve(need)) || (rc = m->uparena.reserve(need))) return rc;
retry:
    m->arena.reset(); m->uparena.reset();
    float *dxy, *dr; int *dmin, *dmax;
    if ((rc = up(m, xy, (size_t)nq * 2, dxy)) || (rc = up(m, r, (size_t)nq, dr)) || (rc = up(m, min_level, (size_t)nq, dmin)) || (rc = up(m, max_level, (size_t)nq, dmax))) return rc;
    if ((rc = flush_uploads(m))) return rc;
    QueryParams P; std::memset(&P, 0, sizeof(P));
    P.mode = MODE_AREA; P.nq = nq; P.q_xy = dxy; P.q_r = dr; P.q_minlevel = dmin; P.q_maxlevel = dmax;
    int *counts, *offsets; uint32_t* cand; uint2* pre;
    if ((rc = window_search(m, P, f->dev, counts, offsets, cand, pre))) return rc;
    CU_TRY(cudaMemcpyAsync(offsets_out, offsets, (size_t)(nq + 1) * 4, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(cudaStreamSynchronize(m->stream));
    const int total = offsets_out[nq];
    CAND_RETRY(total);
    *total_out = total;
    if (total > cap) FAIL(ORBX_E_CAPACITY, "index buffer too small");
    if (total) {
        CU_TRY(cudaMemcpyAsync(indices, cand, (size_t)total * 4, cudaMemcpyDeviceToHost, m->stream));
        CU_TRY(cudaStreamSynchronize(m->stream));
        for (int i = 0; i < total; ++i) indices[i] &= 0xFFFFF;        // entries are dist << 20 | index with dist == 0
    }
    return ORBX_OK;
}

// ---- bag of words ----
struct orbx_vocabulary {
    int device; cudaStream_t stream = nullptr; long long launches = 0;
    int k, L, weighting, scoring, n = 0, nwords = 0;
    FBuf<int> child_off, child_id, word; FBuf<uint8_t> child_desc; FBuf<double> weight;
    // transform scratch
    FBuf<uint8_t> desc, pack; FBuf<int> flags; FBuf<double> weight_of, stage; FBuf<unsigned long long> keys;
    uint8_t* hpack = nullptr; size_t hpack_cap = 0;               // pinned mirror of `pack` (all outputs of one transform)
};
#define VLAUNCH_CHECK() do { cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) { \
    orbx_set_error(std::string("kernel launch: ") + cudaGetErrorString(_e)); return ORBX_E_CUDA; } ++v->launches; } while (0)

int orbx_vocabulary_create(int device, int k, int L, int weighting, int scoring, int n_nodes, const int* parent, const uint8_t* is_leaf,
                           const uint8_t* descriptors, const double* weights, orbx_vocabulary** out) {
    if (!out) FAIL(ORBX_E_INVALID, "null out");
    *out = nullptr;
    if (k < 2 || L < 1 || n_nodes < 1 || n_nodes >= (1 << 28) || weighting < 0 || weighting > 3 || scoring < 0 || scoring > 5 || !parent || !is_leaf || !descriptors || !weights)
        FAIL(ORBX_E_INVALID, "bad vocabulary arguments");
    const int n = n_nodes + 1;
    // children in push_back (= file) order, counted then filled
    std::vector<int> off(n + 1, 0), cid(n_nodes), word(n, -1);
    for (int i = 0; i < n_nodes; ++i) { if (parent[i] < 0 || parent[i] > i) FAIL(ORBX_E_INVALID, "vocabulary node listed before its parent"); ++off[parent[i] + 1]; }
    for (int i = 0; i < n; ++i) off[i + 1] += off[i];
    { std::vector<int> fill(off.begin(), off.end() - 1); for (int i = 0; i < n_nodes; ++i) cid[fill[parent[i]]++] = i + 1; }
    int nw = 0;
    for (int i = 0; i < n_nodes; ++i) {
        const bool childless = off[i + 2] == off[i + 1];
        if ((is_leaf[i] != 0) != childless) FAIL(ORBX_E_INVALID, "vocabulary leaf flag does not agree with the tree");
        if (is_leaf[i]) word[i + 1] = nw++;
    }
    if (off[1] == off[0]) FAIL(ORBX_E_INVALID, "vocabulary root has no children");
    std::vector<uint8_t> cdesc((size_t)n_nodes * 32); std::vector<double> w(n, 0.0);
    for (int c = 0; c < n_nodes; ++c) std::memcpy(&cdesc[(size_t)c * 32], descriptors + (size_t)(cid[c] - 1) * 32, 32);
    for (int i = 0; i < n_nodes; ++i) w[i + 1] = weights[i];
    int ndev = 0;
    CU_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) FAIL(ORBX_E_CUDA, "no such CUDA device (this library has no CPU fallback)");
    CU_TRY(cudaSetDevice(device));
    orbx_vocabulary* v = new orbx_vocabulary();
    v->device = device; v->k = k; v->L = L; v->weighting = weighting; v->scoring = scoring; v->n = n; v->nwords = nw;
    if (cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking) != cudaSuccess) { delete v; FAIL(ORBX_E_CUDA, "cudaStreamCreate"); }
    if (v->child_off.ensure(n + 1) || v->child_id.ensure(n_nodes) || v->word.ensure(n) || v->child_desc.ensure((size_t)n_nodes * 32) || v->weight.ensure(n)) { orbx_vocabulary_destroy(v); return ORBX_E_CUDA; }
    cudaMemcpyAsync(v->child_off.p, off.data(), (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, v->stream);
    cudaMemcpyAsync(v->child_id.p, cid.data(), (size_t)n_nodes * 4, cudaMemcpyHostToDevice, v->stream);
    cudaMemcpyAsync(v->word.p, word.data(), (size_t)n * 4, cudaMemcpyHostToDevice, v->stream);
    cudaMemcpyAsync(v->child_desc.p, cdesc.data(), cdesc.size(), cudaMemcpyHostToDevice, v->stream);
    cudaMemcpyAsync(v->weight.p, w.data(), (size_t)n * 8, cudaMemcpyHostToDevice, v->stream);
    if (cudaStreamSynchronize(v->stream) != cudaSuccess) { orbx_vocabulary_destroy(v); FAIL(ORBX_E_CUDA, "vocabulary upload"); }
    *out = v;
    return ORBX_OK;
}
int orbx_vocabulary_load_text(int device, const char* path, orbx_vocabulary** out) {
    if (!out || !path) FAIL(ORBX_E_INVALID, "null argument");
    *out = nullptr;
    FILE* f = std::fopen(path, "r");
    if (!f) FAIL(ORBX_E_INVALID, std::string("cannot open vocabulary file ") + path);
    int k = 0, L = 0, n1 = -1, n2 = -1;
    if (std::fscanf(f, "%d %d %d %d", &k, &L, &n1, &n2) != 4 || k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) {
        std::fclose(f); FAIL(ORBX_E_INVALID, "not a vocabulary text file");                 // the reference's own header check (TemplatedVocabulary.h:1358)
    }
    std::vector<int> parent; std::vector<uint8_t> leaf, desc; std::vector<double> weight;
    while (true) {
        int pid, isleaf;
        if (std::fscanf(f, "%d %d", &pid, &isleaf) != 2) break;
        uint8_t d[32]; bool ok = true;
        for (int i = 0; i < 32 && ok; ++i) { int b; ok = std::fscanf(f, "%d", &b) == 1; d[i] = (uint8_t)b; }
        double w;
        if (!ok || std::fscanf(f, "%lf", &w) != 1) { std::fclose(f); FAIL(ORBX_E_INVALID, "truncated vocabulary node line"); }
        parent.push_back(pid); leaf.push_back(isleaf > 0 ? 1 : 0); desc.insert(desc.end(), d, d + 32); weight.push_back(w);
    }
    std::fclose(f);
    if (parent.empty()) FAIL(ORBX_E_INVALID, "vocabulary file holds no nodes");
    return orbx_vocabulary_create(device, k, L, n2, n1, (int)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data(), out);
}
void orbx_vocabulary_destroy(orbx_vocabulary* v) {
    if (!v) return;
    cudaSetDevice(v->device); if (v->stream) cudaStreamSynchronize(v->stream);
    v->child_off.release(); v->child_id.release(); v->word.release(); v->child_desc.release(); v->weight.release();
    v->desc.release(); v->pack.release(); v->flags.release(); v->weight_of.release(); v->stage.release(); v->keys.release();
    if (v->hpack) cudaFreeHost(v->hpack);
    if (v->stream) cudaStreamDestroy(v->stream);
    delete v;
}
int orbx_vocabulary_words(const orbx_vocabulary* v) { return v ? v->nwords : -1; }

int orbx_vocabulary_transform(orbx_vocabulary* v, const uint8_t* descriptors, int n, int levelsup, int* word_of, int* node_of,
                              int* bow_ids, double* bow_values, int* n_bow, int* fv_nodes, int* fv_offsets, int* fv_indices, int* n_fv) {
    if (!v || n < 0 || n >= (1 << 20) || !n_bow || !n_fv || !fv_offsets || (n && (!descriptors || !bow_ids || !bow_values || !fv_nodes || !fv_indices)))
        FAIL(ORBX_E_INVALID, "bad arguments");
    *n_bow = 0; *n_fv = 0; fv_offsets[0] = 0;
    if (n == 0) return ORBX_OK;
    CU_TRY(cudaSetDevice(v->device));
    const size_t nn = n;
    int rc;
    // every output lives in ONE device block mirrored by a pinned host block, so the results come back in a single copy:
    // [counts 2 | word_of n | node_of n | bow_ids n | fv_nodes n | fv_offsets n + 1 | fv_idx n] ints, then bow_vals n doubles (8-aligned)
    const size_t n_int = 2 + 6 * nn + 1, off_d = (n_int * 4 + 7) & ~(size_t)7, pack_bytes = off_d + nn * 8;
    if ((rc = v->desc.ensure(nn * 32)) || (rc = v->weight_of.ensure(nn)) || (rc = v->keys.ensure(nn + 1)) || (rc = v->flags.ensure(nn + 1)) || (rc = v->stage.ensure(nn)) ||
        (rc = v->pack.ensure(pack_bytes))) return rc;
    if (v->hpack_cap < pack_bytes) {
        if (v->hpack) cudaFreeHost(v->hpack);
        v->hpack = nullptr; v->hpack_cap = 0;
        CU_TRY(cudaHostAlloc((void**)&v->hpack, pack_bytes + (pack_bytes >> 2), cudaHostAllocDefault));
        v->hpack_cap = pack_bytes + (pack_bytes >> 2);
    }
    int* pi = reinterpret_cast<int*>(v->pack.p);
    int *d_counts = pi, *d_word = pi + 2, *d_node = d_word + nn, *d_bid = d_node + nn, *d_fn = d_bid + nn, *d_fo = d_fn + nn, *d_fi = d_fo + nn + 1;
    double* d_bv = reinterpret_cast<double*>(v->pack.p + off_d);
    cudaStream_t s = v->stream;
    CU_TRY(cudaMemcpyAsync(v->desc.p, descriptors, nn * 32, cudaMemcpyHostToDevice, s));
    VocDev d; d.k = v->k; d.L = v->L; d.weighting = v->weighting; d.scoring = v->scoring; d.n = v->n;
    d.child_off = v->child_off.p; d.child_id = v->child_id.p; d.child_desc = reinterpret_cast<const uint4*>(v->child_desc.p); d.weight = v->weight.p; d.word = v->word.p;
    k_bow_descend<<<(n + 3) / 4, 128, 0, s>>>(d, reinterpret_cast<const uint4*>(v->desc.p), n, levelsup, d_word, d_node, v->weight_of.p);
    VLAUNCH_CHECK();
    k_bow_assemble<<<1, 1024, 0, s>>>(n, v->weighting, v->scoring, d_word, d_node, v->weight_of.p, v->keys.p, v->flags.p, v->stage.p, d_bid, d_bv, d_fn, d_fo, d_fi, d_counts);
    VLAUNCH_CHECK();
    CU_TRY(cudaMemcpyAsync(v->hpack, v->pack.p, pack_bytes, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    const int* hi = reinterpret_cast<const int*>(v->hpack);
    const int nb = hi[0], nf = hi[1];
    if (word_of) std::memcpy(word_of, hi + 2, nn * 4);
    if (node_of) std::memcpy(node_of, hi + 2 + nn, nn * 4);
    std::memcpy(bow_ids, hi + 2 + 2 * nn, (size_t)nb * 4);
    std::memcpy(bow_values, v->hpack + off_d, (size_t)nb * 8);
    std::memcpy(fv_nodes, hi + 2 + 3 * nn, (size_t)nf * 4);
    std::memcpy(fv_offsets, hi + 2 + 4 * nn, (size_t)(nf + 1) * 4);
    std::memcpy(fv_indices, hi + 2 + 5 * nn + 1, (size_t)fv_offsets[nf] * 4);
    *n_bow = nb; *n_fv = nf;
    return ORBX_OK;
}

static int check_bow_side(const orbx_bow_side* s, bool need_valid) {
    if (!s || s->n < 0 || s->n >= (1 << 20) || s->n_fv < 0 || s->n_fv > s->n || (s->n && (!s->keys || !s->descriptors)) || (need_valid && s->n && !s->valid) ||
        (s->n_fv && (!s->fv_nodes || !s->fv_offsets || !s->fv_indices))) FAIL(ORBX_E_INVALID, "bad bag-of-words side");
    if (s->n_fv) {
        if (s->fv_offsets[0] != 0 || s->fv_offsets[s->n_fv] > s->n) FAIL(ORBX_E_INVALID, "bad feature-vector offsets");
        for (int q = 0; q < s->n_fv; ++q) {
            if (s->fv_offsets[q + 1] < s->fv_offsets[q] || (q && s->fv_nodes[q] <= s->fv_nodes[q - 1])) FAIL(ORBX_E_INVALID, "feature vector is not in map order");
        }
        for (int e = 0; e < s->fv_offsets[s->n_fv]; ++e) if (s->fv_indices[e] < 0 || s->fv_indices[e] >= s->n) FAIL(ORBX_E_INVALID, "feature index out of range");
    }
    return ORBX_OK;
}

int orbx_search_by_bow(orbx_matcher* m, int kf_kf, const orbx_bow_side* s1, const orbx_bow_side* s2, int* match12, int* match21, int* nmatches) {
    if (!m || !nmatches) FAIL(ORBX_E_INVALID, "null argument");
    int rc;
    if ((rc = check_bow_side(s1, true)) || (rc = check_bow_side(s2, false))) return rc;
    if ((s1->n && !match12) || (s2->n && !match21)) FAIL(ORBX_E_INVALID, "null output");
    *nmatches = 0;
    for (int i = 0; i < s1->n; ++i) match12[i] = -1;
    for (int j = 0; j < s2->n; ++j) match21[j] = -1;
    // the merge of the two node lists (:250-360): positions of the nodes both feature vectors hold
    std::vector<int2> pairs;
    for (int a = 0, b = 0; a < s1->n_fv && b < s2->n_fv;) {
        if (s1->fv_nodes[a] < s2->fv_nodes[b]) ++a; else if (s2->fv_nodes[b] < s1->fv_nodes[a]) ++b; else { pairs.push_back(make_int2(a, b)); ++a; ++b; }
    }
    if (pairs.empty()) return ORBX_OK;
    CU_TRY(cudaSetDevice(m->device));
    const int n1 = s1->n, n2 = s2->n, np = (int)pairs.size();
    const size_t need = pad((size_t)n1 * 60) + pad((size_t)n2 * 60) + 8 * pad((size_t)(n1 + n2 + 2) * 4) + pad((size_t)np * 8) + 16384;
    if ((rc = m->arena.reserve(need)) || (rc = m->uparena.reserve(need))) return rc;
    m->arena.reset(); m->uparena.reset();
    KpM *k1, *k2; uint8_t *d1, *d2, *v1, *v2 = nullptr; int *o1, *i1, *o2, *i2; int2* dp;
    if ((rc = up(m, reinterpret_cast<const KpM*>(s1->keys), (size_t)n1, k1)) || (rc = up(m, reinterpret_cast<const KpM*>(s2->keys), (size_t)n2, k2)) ||
        (rc = up(m, s1->descriptors, (size_t)n1 * 32, d1)) || (rc = up(m, s2->descriptors, (size_t)n2 * 32, d2)) || (rc = up(m, s1->valid, (size_t)n1, v1)) ||
        (rc = up(m, s1->fv_offsets, (size_t)s1->n_fv + 1, o1)) || (rc = up(m, s1->fv_indices, (size_t)s1->fv_offsets[s1->n_fv], i1)) ||
        (rc = up(m, s2->fv_offsets, (size_t)s2->n_fv + 1, o2)) || (rc = up(m, s2->fv_indices, (size_t)s2->fv_offsets[s2->n_fv], i2)) ||
        (rc = up(m, pairs.data(), (size_t)np, dp))) return rc;
    if (kf_kf && s2->valid && (rc = up(m, s2->valid, (size_t)n2, v2))) return rc;
    int* res = m->arena.get<int>((size_t)n1 + n2 + 3);             // [nmatches | match12 n1 + 1 | match21 n2 + 1]: one copy back
    int* binof = m->arena.get<int>(n1 + 1); int* hist = m->arena.get<int>(32);
    if (!res || !binof || !hist) FAIL(ORBX_E_CUDA, "matcher arena exhausted");
    int *dn = res, *m12 = res + 1, *m21 = res + 2 + n1;
    const size_t res_bytes = ((size_t)n1 + n2 + 3) * 4;
    if ((rc = m->ensure_download(res_bytes))) return rc;
    if ((rc = flush_uploads(m))) return rc;
    cudaStream_t s = m->stream;
    CU_TRY(cudaMemsetAsync(res, 0xFF, res_bytes, s)); CU_TRY(cudaMemsetAsync(hist, 0, 32 * 4, s));
    BowSideDev a = {n1, k1, reinterpret_cast<const uint4*>(d1), v1, o1, i1}, b = {n2, k2, reinterpret_cast<const uint4*>(d2), v2, o2, i2};
    k_bow_match<<<(np + 3) / 4, 128, 0, s>>>(np, dp, a, b, kf_kf ? 1 : 0, m->nnratio, m->checkOri, m12, m21, binof, hist);
    LAUNCH_CHECK();
    k_bow_finish<<<1, 1024, 0, s>>>(n1, m->checkOri, hist, binof, m12, m21, dn);
    LAUNCH_CHECK();
    CU_TRY(cudaMemcpyAsync(m->dl_host, res, res_bytes, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    const int* hr = reinterpret_cast<const int*>(m->dl_host);
    *nmatches = hr[0];
    if (n1) std::memcpy(match12, hr + 1, (size_t)n1 * 4);
    if (n2) std::memcpy(match21, hr + 2 + n1, (size_t)n2 * 4);
    return ORBX_OK;
}

int orbx_search_for_triangulation(orbx_matcher* m, const orbx_bow_side* s1, const orbx_bow_side* s2, const float* u_right1, const float* u_right2,
                                  const float* F12, float ex, float ey, int nlevels2, const float* scale_factors2, const float* level_sigma2_2, int only_stereo,
                                  int* match12, int* nmatches) {
    if (!m || !nmatches || !F12 || nlevels2 <= 0 || nlevels2 > ORBX_MAX_LEVELS || !scale_factors2 || !level_sigma2_2) FAIL(ORBX_E_INVALID, "null argument");
    int rc;
    if ((rc = check_bow_side(s1, true)) || (rc = check_bow_side(s2, true))) return rc;
    if ((s1->n && (!match12 || !u_right1)) || (s2->n && !u_right2)) FAIL(ORBX_E_INVALID, "null buffer");
    for (int j = 0; j < s2->n; ++j) if (s2->keys[j].octave < 0 || s2->keys[j].octave >= nlevels2) FAIL(ORBX_E_INVALID, "octave out of range");
    *nmatches = 0;
    for (int i = 0; i < s1->n; ++i) match12[i] = -1;
    std::vector<int2> pairs;
    for (int a = 0, b = 0; a < s1->n_fv && b < s2->n_fv;) {
        if (s1->fv_nodes[a] < s2->fv_nodes[b]) ++a; else if (s2->fv_nodes[b] < s1->fv_nodes[a]) ++b; else { pairs.push_back(make_int2(a, b)); ++a; ++b; }
    }
    if (pairs.empty()) return ORBX_OK;
    CU_TRY(cudaSetDevice(m->device));
    const int n1 = s1->n, n2 = s2->n, np = (int)pairs.size();
    const size_t need = pad((size_t)n1 * 68) + pad((size_t)n2 * 68) + 8 * pad((size_t)(n1 + n2 + 2) * 4) + pad((size_t)np * 8) + 16384;
    if ((rc = m->arena.reserve(need)) || (rc = m->uparena.reserve(need))) return rc;
    m->arena.reset(); m->uparena.reset();
    KpM *k1, *k2; uint8_t *d1, *d2, *v1, *v2; int *o1, *i1, *o2, *i2; int2* dp; float *ur1, *ur2, *sc2, *sg2;
    if ((rc = up(m, reinterpret_cast<const KpM*>(s1->keys), (size_t)n1, k1)) || (rc = up(m, reinterpret_cast<const KpM*>(s2->keys), (size_t)n2, k2)) ||
        (rc = up(m, s1->descriptors, (size_t)n1 * 32, d1)) || (rc = up(m, s2->descriptors, (size_t)n2 * 32, d2)) || (rc = up(m, s1->valid, (size_t)n1, v1)) ||
        (rc = up(m, s2->valid, (size_t)n2, v2)) || (rc = up(m, u_right1, (size_t)n1, ur1)) || (rc = up(m, u_right2, (size_t)n2, ur2)) ||
        (rc = up(m, scale_factors2, (size_t)nlevels2, sc2)) || (rc = up(m, level_sigma2_2, (size_t)nlevels2, sg2)) ||
        (rc = up(m, s1->fv_offsets, (size_t)s1->n_fv + 1, o1)) || (rc = up(m, s1->fv_indices, (size_t)s1->fv_offsets[s1->n_fv], i1)) ||
        (rc = up(m, s2->fv_offsets, (size_t)s2->n_fv + 1, o2)) || (rc = up(m, s2->fv_indices, (size_t)s2->fv_offsets[s2->n_fv], i2)) ||
        (rc = up(m, pairs.data(), (size_t)np, dp))) return rc;
    int* res = m->arena.get<int>((size_t)n1 + n2 + 3);             // [nmatches | match12 n1 + 1 | matched2 n2 + 1]
    int* binof = m->arena.get<int>(n1 + 1); int* hist = m->arena.get<int>(32);
    if (!res || !binof || !hist) FAIL(ORBX_E_CUDA, "matcher arena exhausted");
    int *dn = res, *m12 = res + 1, *m21 = res + 2 + n1;
    const size_t res_bytes = ((size_t)n1 + 2) * 4;
    if ((rc = m->ensure_download(res_bytes))) return rc;
    if ((rc = flush_uploads(m))) return rc;
    cudaStream_t s = m->stream;
    CU_TRY(cudaMemsetAsync(res, 0xFF, ((size_t)n1 + n2 + 3) * 4, s)); CU_TRY(cudaMemsetAsync(hist, 0, 32 * 4, s));
    BowSideDev a = {n1, k1, reinterpret_cast<const uint4*>(d1), v1, o1, i1}, b = {n2, k2, reinterpret_cast<const uint4*>(d2), v2, o2, i2};
    TriParams T; std::memcpy(T.F, F12, 36); T.ex = ex; T.ey = ey; T.ur1 = ur1; T.ur2 = ur2; T.scale2 = sc2; T.sigma2_2 = sg2; T.only_stereo = only_stereo ? 1 : 0;
    k_tri_match<<<(np + 3) / 4, 128, 0, s>>>(np, dp, a, b, T, m->checkOri, m12, m21, binof, hist);
    LAUNCH_CHECK();
    k_bow_finish<<<1, 1024, 0, s>>>(n1, m->checkOri, hist, binof, m12, m21, dn);
    LAUNCH_CHECK();
    CU_TRY(cudaMemcpyAsync(m->dl_host, res, res_bytes, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    const int* hr = reinterpret_cast<const int*>(m->dl_host);
    *nmatches = hr[0];
    if (n1) std::memcpy(match12, hr + 1, (size_t)n1 * 4);
    return ORBX_OK;
}

int orbx_distinctive_descriptors(orbx_matcher* m, int n_points, const int* offsets, const uint8_t* descriptors, int* best_idx) {
    if (!m || n_points < 0 || (n_points && (!offsets || !best_idx))) FAIL(ORBX_E_INVALID, "bad arguments");
    if (n_points == 0) return ORBX_OK;
    if (offsets[0] != 0) FAIL(ORBX_E_INVALID, "offsets must start at 0");
    for (int p = 0; p < n_points; ++p) if (offsets[p + 1] < offsets[p]) FAIL(ORBX_E_INVALID, "offsets must not decrease");
    const int total = offsets[n_points];
    if (total && !descriptors) FAIL(ORBX_E_INVALID, "null descriptors");
    CU_TRY(cudaSetDevice(m->device));
    int rc;
    const size_t need = pad((size_t)total * 32) + 2 * pad((size_t)(n_points + 1) * 4) + 8192;
    if ((rc = m->arena.reserve(need)) || (rc = m->uparena.reserve(need))) return rc;
    m->arena.reset(); m->uparena.reset();
    uint8_t* dd; int* doff;
    if ((rc = up(m, descriptors, (size_t)total * 32, dd)) || (rc = up(m, offsets, (size_t)n_points + 1, doff))) return rc;
    int* dbest = m->arena.get<int>(n_points);
    if (!dbest) FAIL(ORBX_E_CUDA, "matcher arena exhausted");
    if ((rc = flush_uploads(m)) || (rc = m->ensure_download((size_t)n_points * 4))) return rc;
    k_distinctive<<<(n_points + 3) / 4, 128, 0, m->stream>>>(n_points, doff, reinterpret_cast<const uint4*>(dd), dbest);
    LAUNCH_CHECK();
    CU_TRY(cudaMemcpyAsync(m->dl_host, dbest, (size_t)n_points * 4, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(cudaStreamSynchronize(m->stream));
    std::memcpy(best_idx, m->dl_host, (size_t)n_points * 4);
    return ORBX_OK;
}

int orbx_match_bruteforce_device(orbx_matcher* m, const uint8_t* d_query, int n_query, const uint8_t* d_train, int n_train, int* d_best_idx, int* d_best_dist, int* d_second_dist) {
    return orbx_match_bruteforce_batch_device(m, 1, d_query, n_query, d_train, n_train, d_best_idx, d_best_dist, d_second_dist);
}

int orbx_match_bruteforce_batch_device(orbx_matcher* m, int n_pairs, const uint8_t* d_query, int n_query, const uint8_t* d_train, int n_train, int* d_best_idx, int* d_best_dist, int* d_second_dist) {
    if (!m || n_pairs < 0 || n_pairs > 65535 || n_query < 0 || n_train < 0 || n_train >= (1 << 20) || !d_query || !d_train || !d_best_idx || !d_best_dist || !d_second_dist) FAIL(ORBX_E_INVALID, "bad arguments");
    if (((uintptr_t)d_query & 15) || ((uintptr_t)d_train & 15)) FAIL(ORBX_E_INVALID, "descriptor arrays must be 16-byte aligned");
    if (n_query == 0 || n_pairs == 0) return ORBX_OK;
    CU_TRY(cudaSetDevice(m->device));
    k_bruteforce_best2<<<dim3((n_query + 8 * BF_QPW - 1) / (8 * BF_QPW), n_pairs), 256, 0, m->stream>>>(reinterpret_cast<const uint4*>(d_query), n_query, reinterpret_cast<const uint4*>(d_train), n_train, d_best_idx, d_best_dist, d_second_dist);
    LAUNCH_CHECK();
    return ORBX_OK;
}

}  // extern "C"
