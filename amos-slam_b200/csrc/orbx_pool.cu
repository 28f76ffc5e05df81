// orbx_pool.cu -- multi-sequence / multi-GPU driver of the extractor: sequences (camera streams) are pinned to
// workers by  gpu = seq_id mod G,  stream = (seq_id div G) mod S;  every worker is one host thread that owns one extractor handle (and with it
// one CUDA stream and one resident pyramid) on its GPU and executes its jobs in submission order.  Frames and sequences are independent units:
// there is no collective and no shared device state between workers -- the reference's "one extractor object per camera, objects run in their
// own threads" (/root/reference/src/Frame.cc:165-173) scaled out over the GPUs of one box.  Built only on the C ABI of include/orbx_b200.h.
#include "../../include/orbx_b200.h"
#include <cuda_runtime.h>

#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

void orbx_set_error(const std::string& s);

namespace {
struct Job {
    long long ticket; int kind;        // 0 = single frame, 1 = batch, 2 = masked batch (labels optional)
    const uint8_t* images; const uint8_t* masks; orbx_labels labels; bool has_labels;
    int B, rows, cols; size_t step, frame_stride, mask_step, mask_frame_stride;
    orbx_keypoint* kp_out; uint8_t* desc_out; int cap; int* counts_out; int* culled_out;
};
struct Worker {
    int device = 0, index = 0; orbx_extractor* ext = nullptr; std::thread th;
    std::mutex mu; std::condition_variable cv; std::deque<Job> q; bool stop = false; int create_rc = ORBX_OK; std::string create_err; bool ready = false;
    long long jobs_done = 0, frames_done = 0;
};
}  // namespace

struct orbx_pool {
    int nfeatures; float scale; int nlevels, iniTh, minTh;
    int G = 0, S = 0; std::vector<int> devices;
    std::vector<std::unique_ptr<Worker>> workers;             // worker w = g * S + s
    std::mutex mu; std::condition_variable cv_done;
    long long next_ticket = 1; std::map<long long, int> pending;     // ticket -> status once finished (ORBX_OK, ...); absent = still running
    std::map<long long, std::string> errors; long long open_jobs = 0;
};

static void worker_main(orbx_pool* p, Worker* w) {
    int rc = orbx_create(p->nfeatures, p->scale, p->nlevels, p->iniTh, p->minTh, w->device, &w->ext);
    {
        std::lock_guard<std::mutex> lk(w->mu);
        w->create_rc = rc; if (rc) w->create_err = orbx_last_error();
        w->ready = true;
    }
    w->cv.notify_all();
    if (rc) return;
    for (;;) {
        Job j;
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return w->stop || !w->q.empty(); });
            if (w->q.empty()) break;                         // stop requested and nothing left to do
            j = w->q.front(); w->q.pop_front();
        }
        int st;
        if (j.kind == 0) st = orbx_extract(w->ext, j.images, j.rows, j.cols, j.step, j.kp_out, j.desc_out, j.cap, j.counts_out);
        else if (j.kind == 1) st = orbx_extract_batch(w->ext, j.images, j.B, j.rows, j.cols, j.step, j.frame_stride, j.kp_out, j.desc_out, j.cap, j.counts_out);
        else st = orbx_extract_masked_batch_labels(w->ext, j.images, j.masks, j.has_labels ? &j.labels : nullptr, j.B, j.rows, j.cols, j.step, j.frame_stride, j.mask_step,
                                                   j.mask_frame_stride, j.kp_out, j.desc_out, j.cap, j.counts_out, j.culled_out);
        {
            std::lock_guard<std::mutex> lk(p->mu);
            p->pending[j.ticket] = st;
            if (st) p->errors[j.ticket] = orbx_last_error();
            --p->open_jobs; ++w->jobs_done; w->frames_done += j.kind == 0 ? 1 : j.B;
        }
        p->cv_done.notify_all();
    }
    orbx_destroy(w->ext); w->ext = nullptr;
}

extern "C" {

int orbx_pool_shard_of(int seq_id, int n_gpus, int streams_per_gpu, int* gpu, int* stream) {
    if (seq_id < 0 || n_gpus <= 0 || streams_per_gpu <= 0) return ORBX_E_INVALID;
    if (gpu) *gpu = seq_id % n_gpus;
    if (stream) *stream = (seq_id / n_gpus) % streams_per_gpu;
    return ORBX_OK;
}

int orbx_pool_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int n_gpus, const int* devices, int streams_per_gpu, orbx_pool** out) {
    if (!out) { orbx_set_error("null out"); return ORBX_E_INVALID; }
    *out = nullptr;
    if (n_gpus <= 0 || streams_per_gpu <= 0 || streams_per_gpu > 16) { orbx_set_error("bad pool shape"); return ORBX_E_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { orbx_set_error("no CUDA device (this library has no CPU fallback)"); return ORBX_E_CUDA; }
    std::unique_ptr<orbx_pool> p(new orbx_pool());
    p->nfeatures = nfeatures; p->scale = scaleFactor; p->nlevels = nlevels; p->iniTh = iniThFAST; p->minTh = minThFAST; p->G = n_gpus; p->S = streams_per_gpu;
    for (int g = 0; g < n_gpus; ++g) {
        const int d = devices ? devices[g] : g;
        if (d < 0 || d >= ndev) { orbx_set_error("pool device ordinal out of range"); return ORBX_E_CUDA; }
        p->devices.push_back(d);
    }
    for (int g = 0; g < n_gpus; ++g)
        for (int s = 0; s < streams_per_gpu; ++s) {
            std::unique_ptr<Worker> w(new Worker());
            w->device = p->devices[g]; w->index = g * streams_per_gpu + s;
            p->workers.push_back(std::move(w));
        }
    for (auto& w : p->workers) w->th = std::thread(worker_main, p.get(), w.get());
    int rc = ORBX_OK; std::string err;
    for (auto& w : p->workers) {
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return w->ready; });
        if (w->create_rc && !rc) { rc = w->create_rc; err = w->create_err; }
    }
    if (rc) {
        for (auto& w : p->workers) { { std::lock_guard<std::mutex> lk(w->mu); w->stop = true; } w->cv.notify_all(); }
        for (auto& w : p->workers) if (w->th.joinable()) w->th.join();
        orbx_set_error(err);
        return rc;
    }
    *out = p.release();
    return ORBX_OK;
}

void orbx_pool_destroy(orbx_pool* p) {
    if (!p) return;
    for (auto& w : p->workers) { { std::lock_guard<std::mutex> lk(w->mu); w->stop = true; } w->cv.notify_all(); }      // queued jobs are finished first
    for (auto& w : p->workers) if (w->th.joinable()) w->th.join();
    delete p;
}

int orbx_pool_gpus(const orbx_pool* p) { return p ? p->G : 0; }
int orbx_pool_streams_per_gpu(const orbx_pool* p) { return p ? p->S : 0; }
int orbx_pool_device_of(const orbx_pool* p, int seq_id) {
    int g = 0;
    if (!p || orbx_pool_shard_of(seq_id, p->G, p->S, &g, nullptr)) return ORBX_E_INVALID;
    return p->devices[g];
}
long long orbx_pool_frames_done(const orbx_pool* p, int gpu, int stream) {
    if (!p || gpu < 0 || gpu >= p->G || stream < 0 || stream >= p->S) return -1;
    orbx_pool* q = const_cast<orbx_pool*>(p);
    std::lock_guard<std::mutex> lk(q->mu);
    return p->workers[(size_t)gpu * p->S + stream]->frames_done;
}

static int pool_enqueue(orbx_pool* p, int seq_id, Job j, long long* ticket) {
    int g = 0, s = 0;
    if (!p || orbx_pool_shard_of(seq_id, p->G, p->S, &g, &s)) { orbx_set_error("bad pool / sequence id"); return ORBX_E_INVALID; }
    if (!j.images || j.rows <= 0 || j.cols <= 0 || j.step < (size_t)j.cols || !j.kp_out || !j.desc_out || !j.counts_out || j.cap <= 0 || j.B <= 0) { orbx_set_error("bad job arguments"); return ORBX_E_INVALID; }
    Worker* w = p->workers[(size_t)g * p->S + s].get();
    {
        std::lock_guard<std::mutex> lk(p->mu);
        j.ticket = p->next_ticket++; ++p->open_jobs;
    }
    if (ticket) *ticket = j.ticket;
    { std::lock_guard<std::mutex> lk(w->mu); w->q.push_back(j); }
    w->cv.notify_one();
    return ORBX_OK;
}

int orbx_pool_submit(orbx_pool* p, int seq_id, const uint8_t* image, int rows, int cols, size_t step, orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* n_out, long long* ticket) {
    Job j{}; j.kind = 0; j.images = image; j.B = 1; j.rows = rows; j.cols = cols; j.step = step; j.kp_out = kp_out; j.desc_out = desc_out; j.cap = cap; j.counts_out = n_out;
    return pool_enqueue(p, seq_id, j, ticket);
}
int orbx_pool_submit_batch(orbx_pool* p, int seq_id, const uint8_t* images, const uint8_t* masks, const orbx_labels* labels, int B, int rows, int cols, size_t step, size_t frame_stride,
                           size_t mask_step, size_t mask_frame_stride, orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* counts_out, int* culled_out, long long* ticket) {
    Job j{}; j.kind = masks ? 2 : 1; j.images = images; j.masks = masks; j.has_labels = labels != nullptr; if (labels) j.labels = *labels;
    j.B = B; j.rows = rows; j.cols = cols; j.step = step; j.frame_stride = frame_stride; j.mask_step = mask_step; j.mask_frame_stride = mask_frame_stride;
    j.kp_out = kp_out; j.desc_out = desc_out; j.cap = cap; j.counts_out = counts_out; j.culled_out = culled_out;
    if (labels && !masks) { orbx_set_error("labels need masks (MovingKeyPoints takes both)"); return ORBX_E_INVALID; }
    return pool_enqueue(p, seq_id, j, ticket);
}
int orbx_pool_wait(orbx_pool* p, long long ticket) {
    if (!p || ticket <= 0) return ORBX_E_INVALID;
    std::unique_lock<std::mutex> lk(p->mu);
    if (ticket >= p->next_ticket) { orbx_set_error("unknown ticket"); return ORBX_E_INVALID; }
    p->cv_done.wait(lk, [&] { return p->pending.count(ticket) != 0; });
    const int st = p->pending[ticket];
    p->pending.erase(ticket);
    if (st) { orbx_set_error(p->errors[ticket]); p->errors.erase(ticket); }
    return st;
}
int orbx_pool_wait_all(orbx_pool* p) {
    if (!p) return ORBX_E_INVALID;
    std::unique_lock<std::mutex> lk(p->mu);
    p->cv_done.wait(lk, [&] { return p->open_jobs == 0; });
    int st = ORBX_OK;
    for (auto& kv : p->pending) if (kv.second && !st) { st = kv.second; orbx_set_error(p->errors[kv.first]); }
    p->pending.clear(); p->errors.clear();
    return st;
}

}  // extern "C"
