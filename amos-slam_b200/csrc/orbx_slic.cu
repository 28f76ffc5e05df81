// orbx_slic.cu -- SLIC super-pixel stage of the reference's `cluster` (/root/reference/src/cluster.cc:88-344) on the GPU, behind the C ABI
// (include/orbx_b200.h: orbx_slic_*).  The label map it produces is what MovingKeyPoints reads (src/ORBextractor.cc:1722-1736).
//
// Boundary: cv::cvtColor(image, COLOR_BGR2Lab) stays with the caller's OpenCV (the drop-in host code calls it as the reference does,
// src/cluster.cc:305); everything after it runs here.  The reference walks the centres one after the other and lets each overwrite the
// pixels of its 2len x 2len window it is strictly closer to (:118-143) -- per pixel that is "the covering centre with the smallest
// distance, lowest index on ties; pixels no window covers keep the label of the round before", which is order-free:
//   k_slic_init    initilizeCenters + fituneCenter (:207-283): one thread per centre; the Sobel gradient (CV_64F, 0.5 / 0.5 blend) is a
//                  half-integer, so the squared-gradient comparison is done exactly on integers (2g)
//   k_slic_bin_*   centres bucketed by (y / len, x / len) every round (counting sort): a pixel only looks into the 3 x 3 buckets its
//                  covering centres can lie in.  Buckets have no capacity bound (dead centres collect at (0, 0), :196-201)
//   k_slic_assign  one thread per pixel: dis = sqrt(disc^2 + m diss^2) in double with individually rounded operations (the oracle
//                  contract is -ffp-contract=off), disc / diss themselves correctly rounded square roots of integers
//   k_slic_update  updateCenter (:160-203): one warp per centre sums x, y, L, A, B, D over the pixels of its window that carry its
//                  label (integer sums, exact), centre := truncated double quotients; an empty centre becomes all zeros as in the reference
// 5 rounds = 26 small launches, ~0.1 ms for a 640 x 480 frame; the reference spends ~0.2 s per frame here on one core.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <algorithm>
#include "../../include/orbx_b200.h"

void orbx_set_error(const std::string& s);          // orbx_extractor.cu
#define SL_FAIL(code, msg) do { orbx_set_error(msg); return (code); } while (0)
#define SL_TRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { orbx_set_error(std::string(#x ": ") + cudaGetErrorString(_e)); return ORBX_E_CUDA; } } while (0)

struct SlicGeom { int rows, cols, len, m, ncx, ncy, n, nbx, nby; };
struct SlicCenters { int *x, *y, *L, *A, *B, *D; };

__device__ __forceinline__ int slic_refl(int p, int n) { if (n == 1) return 0; while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p; return p; }

__global__ void k_slic_init(SlicGeom g, const uint8_t* __restrict__ lab, const uint16_t* __restrict__ depth, SlicCenters c) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= g.n) return;
    const int gi = k / g.ncx, gj = k - gi * g.ncx;
    int cy = gi * g.len + g.len / 2, cx = gj * g.len + g.len / 2;
    c.D[k] = depth[(size_t)cy * g.cols + cx];                       // (the reference's "D < 0" test on an unsigned short never fires)
    if (!(cx - 1 < 0 || cx + 1 >= g.cols || cy - 1 < 0 || cy + 1 >= g.rows)) {
        int best = 0x7FFFFFFF, tx = 0, ty = 0;                      // 4 * 9999999 > any reachable value: the first neighbour always wins first
        for (int mm = -1; mm < 2; ++mm) for (int nn = -1; nn < 2; ++nn) {
            int sum = 0;
            for (int ch = 0; ch < 3; ++ch) {
                int gx = 0, gy = 0;
#pragma unroll
                for (int i = -1; i <= 1; ++i) {
                    const uint8_t* row = lab + (size_t)slic_refl(cy + mm + i, g.rows) * g.cols * 3;
                    const int a = row[slic_refl(cx + nn - 1, g.cols) * 3 + ch], b = row[slic_refl(cx + nn, g.cols) * 3 + ch], d = row[slic_refl(cx + nn + 1, g.cols) * 3 + ch];
                    gy += i * (a + 2 * b + d);                       // Sobel(dx = 0, dy = 1)
                    gx += (i == 0 ? 2 : 1) * (d - a);                // Sobel(dx = 1, dy = 0)
                }
                const int G = gy + gx;                               // = 2 * (0.5 gy + 0.5 gx)
                sum += G * G;
            }
            if (sum < best) { best = sum; ty = mm; tx = nn; }
        }
        cx += tx; cy += ty;
    }
    const uint8_t* p = lab + ((size_t)cy * g.cols + cx) * 3;
    c.x[k] = cx; c.y[k] = cy; c.L[k] = p[0]; c.A[k] = p[1]; c.B[k] = p[2];
}

__global__ void k_slic_bin_count(SlicGeom g, SlicCenters c, int* __restrict__ cnt) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= g.n) return;
    atomicAdd(&cnt[(c.y[k] / g.len) * g.nbx + c.x[k] / g.len], 1);
}
// exclusive scan of the bucket counts by one block (buckets ~ centres: a few 10^4); start[nb] = total; cursor := start
__global__ void k_slic_bin_scan(int nb, const int* __restrict__ cnt, int* __restrict__ start, int* __restrict__ cursor) {
    __shared__ int part[1024];
    const int t = threadIdx.x, per = (nb + 1023) / 1024;
    const int a = min(t * per, nb), b = min(a + per, nb);
    int s = 0;
    for (int i = a; i < b; ++i) s += cnt[i];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) { const int v = t >= o ? part[t - o] : 0; __syncthreads(); part[t] += v; __syncthreads(); }
    int run = part[t] - s;
    for (int i = a; i < b; ++i) { start[i] = run; cursor[i] = run; run += cnt[i]; }
    if (t == 1023) start[nb] = part[1023];
}
__global__ void k_slic_bin_fill(SlicGeom g, SlicCenters c, int* __restrict__ cursor, int* __restrict__ items) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= g.n) return;
    items[atomicAdd(&cursor[(c.y[k] / g.len) * g.nbx + c.x[k] / g.len], 1)] = k;
}

__global__ void k_slic_assign(SlicGeom g, const uint8_t* __restrict__ lab, SlicCenters c, const int* __restrict__ start, const int* __restrict__ items, int* __restrict__ labels) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= g.cols) return;
    const uint8_t* p = lab + ((size_t)y * g.cols + x) * 3;
    const int L = p[0], A = p[1], B = p[2];
    // covering centres: x - len < cx <= x + len, same in y
    const int lo_x = x - g.len + 1, lo_y = y - g.len + 1;
    const int bx0 = max(0, (lo_x < 0 ? -((-lo_x + g.len - 1) / g.len) : lo_x / g.len)), bx1 = min(g.nbx - 1, (x + g.len) / g.len);
    const int by0 = max(0, (lo_y < 0 ? -((-lo_y + g.len - 1) / g.len) : lo_y / g.len)), by1 = min(g.nby - 1, (y + g.len) / g.len);
    double best = 999999.0; int who = -1;
    const double md = (double)g.m;
    for (int by = by0; by <= by1; ++by) for (int bx = bx0; bx <= bx1; ++bx) {
        const int b = by * g.nbx + bx;
        for (int t = start[b]; t < start[b + 1]; ++t) {
            const int k = items[t];
            const int cx = c.x[k], cy = c.y[k];
            if (x < cx - g.len || x >= cx + g.len || y < cy - g.len || y >= cy + g.len) continue;
            const int dL = L - c.L[k], dA = A - c.A[k], dB = B - c.B[k], dx = x - cx, dy = y - cy;
            const double disc = __dsqrt_rn((double)(dL * dL + dA * dA + dB * dB));
            const double diss = __dsqrt_rn((double)(dx * dx + dy * dy));
            const double dis = __dsqrt_rn(__dadd_rn(__dmul_rn(disc, disc), __dmul_rn(md, __dmul_rn(diss, diss))));
            if (dis < best || (dis == best && k < who)) { best = dis; who = k; }     // the reference's walk keeps the LOWEST index among equals
        }
    }
    if (who >= 0) labels[(size_t)y * g.cols + x] = who + 1;
}

__global__ void k_slic_update(SlicGeom g, const uint8_t* __restrict__ lab, const uint16_t* __restrict__ depth, const int* __restrict__ labels, SlicCenters c) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= g.n) return;
    const int cx = c.x[k], cy = c.y[k], w = 2 * g.len;
    int sx = 0, sy = 0, sL = 0, sA = 0, sB = 0, sn = 0; unsigned sD = 0;
    for (int q = lane; q < w * w; q += 32) {
        const int i = cy - g.len + q / w, j = cx - g.len + q % w;
        if (i < 0 || i >= g.rows || j < 0 || j >= g.cols) continue;
        const size_t o = (size_t)i * g.cols + j;
        if (labels[o] != k + 1) continue;
        const uint8_t* p = lab + o * 3;
        sL += p[0]; sA += p[1]; sB += p[2]; sx += j; sy += i; sn += 1; sD += depth[o];
    }
    sx = __reduce_add_sync(0xffffffffu, sx); sy = __reduce_add_sync(0xffffffffu, sy); sL = __reduce_add_sync(0xffffffffu, sL);
    sA = __reduce_add_sync(0xffffffffu, sA); sB = __reduce_add_sync(0xffffffffu, sB); sn = __reduce_add_sync(0xffffffffu, sn);
    sD = __reduce_add_sync(0xffffffffu, sD);
    if (lane == 0) {
        const double num = sn == 0 ? 0.000000001 : (double)sn;      // :195
        c.x[k] = __double2int_rz(__ddiv_rn((double)sx, num)); c.y[k] = __double2int_rz(__ddiv_rn((double)sy, num));
        c.L[k] = __double2int_rz(__ddiv_rn((double)sL, num)); c.A[k] = __double2int_rz(__ddiv_rn((double)sA, num));
        c.B[k] = __double2int_rz(__ddiv_rn((double)sB, num)); c.D[k] = __double2int_rz(__ddiv_rn((double)sD, num));
    }
}

__global__ void k_slic_export(int n, const int* __restrict__ labels, double* __restrict__ out64, uint16_t* __restrict__ out16) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int v = labels[i];
    if (out64) out64[i] = (double)v;
    if (out16) out16[i] = (uint16_t)v;
}
__global__ void k_slic_pack_centers(int n, SlicCenters c, int* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int* o = out + (size_t)k * 7;
    o[0] = c.x[k]; o[1] = c.y[k]; o[2] = c.L[k]; o[3] = c.A[k]; o[4] = c.B[k]; o[5] = c.D[k]; o[6] = k + 1;
}

struct orbx_slic {
    int device = 0; cudaStream_t stream = nullptr;
    size_t px_cap = 0, n_cap = 0;
    uint8_t* d_lab = nullptr; uint16_t* d_depth = nullptr; int* d_labels = nullptr; double* d_out64 = nullptr; uint16_t* d_out16 = nullptr;
    int* d_cent = nullptr;      // 6 x n_cap centre fields + 7 x n_cap packed output
    int *d_cnt = nullptr, *d_start = nullptr, *d_cursor = nullptr, *d_items = nullptr;
};

static void slic_free(orbx_slic* h) {
    cudaFree(h->d_lab); cudaFree(h->d_depth); cudaFree(h->d_labels); cudaFree(h->d_out64); cudaFree(h->d_out16); cudaFree(h->d_cent);
    cudaFree(h->d_cnt); cudaFree(h->d_start); cudaFree(h->d_cursor); cudaFree(h->d_items);
    h->d_lab = nullptr; h->d_depth = nullptr; h->d_labels = nullptr; h->d_out64 = nullptr; h->d_out16 = nullptr; h->d_cent = nullptr;
    h->d_cnt = h->d_start = h->d_cursor = h->d_items = nullptr; h->px_cap = h->n_cap = 0;
}

extern "C" {

int orbx_slic_create(int device, orbx_slic** out) {
    if (!out) SL_FAIL(ORBX_E_INVALID, "null out");
    *out = nullptr;
    int ndev = 0;
    SL_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) SL_FAIL(ORBX_E_CUDA, "no such CUDA device (this library has no CPU fallback)");
    SL_TRY(cudaSetDevice(device));
    orbx_slic* h = new orbx_slic(); h->device = device;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; SL_FAIL(ORBX_E_CUDA, "cudaStreamCreate"); }
    *out = h;
    return ORBX_OK;
}

void orbx_slic_destroy(orbx_slic* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    slic_free(h);
    cudaStreamDestroy(h->stream);
    delete h;
}

int orbx_slic_run(orbx_slic* h, const uint8_t* lab, size_t lab_step, const uint16_t* depth, size_t depth_step, int rows, int cols, int len, int m, int rounds,
                  double* labels_out, size_t labels_step, uint16_t* labels16_out, size_t labels16_step, orbx_slic_center* centers_out, int cap, int* n_out) {
    if (n_out) *n_out = 0;
    if (!h || !lab || !depth || !n_out) SL_FAIL(ORBX_E_INVALID, "null argument");
    if (rows <= 0 || cols <= 0 || len < 1 || len > 64 || rounds < 0 || lab_step < (size_t)cols * 3 || depth_step < (size_t)cols * 2) SL_FAIL(ORBX_E_INVALID, "bad SLIC arguments");
    if ((labels_out && labels_step < (size_t)cols * 8) || (labels16_out && labels16_step < (size_t)cols * 2)) SL_FAIL(ORBX_E_INVALID, "bad label output step");
    SL_TRY(cudaSetDevice(h->device));
    SlicGeom g; g.rows = rows; g.cols = cols; g.len = len; g.m = m;
    g.ncx = cols > len / 2 ? (cols - len / 2 + len - 1) / len : 0; g.ncy = rows > len / 2 ? (rows - len / 2 + len - 1) / len : 0;
    g.n = g.ncx * g.ncy; g.nbx = (cols + len - 1) / len; g.nby = (rows + len - 1) / len;
    *n_out = g.n;
    if (labels16_out && g.n > 65535) SL_FAIL(ORBX_E_INVALID, "more than 65535 super-pixels: 16-bit labels cannot hold them");
    if (centers_out && cap < g.n) SL_FAIL(ORBX_E_CAPACITY, "centre buffer too small");
    const size_t px = (size_t)rows * cols, nb = (size_t)g.nbx * g.nby;
    if (px > h->px_cap || (size_t)g.n > h->n_cap || nb + 1 > h->n_cap + 1) {
        slic_free(h);
        const size_t nc = std::max<size_t>((size_t)g.n, nb) + 1;
        SL_TRY(cudaMalloc((void**)&h->d_lab, px * 3)); SL_TRY(cudaMalloc((void**)&h->d_depth, px * 2)); SL_TRY(cudaMalloc((void**)&h->d_labels, px * 4));
        SL_TRY(cudaMalloc((void**)&h->d_out64, px * 8)); SL_TRY(cudaMalloc((void**)&h->d_out16, px * 2));
        SL_TRY(cudaMalloc((void**)&h->d_cent, nc * 13 * sizeof(int)));
        SL_TRY(cudaMalloc((void**)&h->d_cnt, (nc + 1) * 4)); SL_TRY(cudaMalloc((void**)&h->d_start, (nc + 1) * 4)); SL_TRY(cudaMalloc((void**)&h->d_cursor, (nc + 1) * 4)); SL_TRY(cudaMalloc((void**)&h->d_items, nc * 4));
        h->px_cap = px; h->n_cap = nc;
    }
    cudaStream_t s = h->stream;
    SL_TRY(cudaMemcpy2DAsync(h->d_lab, (size_t)cols * 3, lab, lab_step, (size_t)cols * 3, rows, cudaMemcpyHostToDevice, s));
    SL_TRY(cudaMemcpy2DAsync(h->d_depth, (size_t)cols * 2, depth, depth_step, (size_t)cols * 2, rows, cudaMemcpyHostToDevice, s));
    SL_TRY(cudaMemsetAsync(h->d_labels, 0, px * 4, s));
    SlicCenters c; c.x = h->d_cent; c.y = c.x + h->n_cap; c.L = c.y + h->n_cap; c.A = c.L + h->n_cap; c.B = c.A + h->n_cap; c.D = c.B + h->n_cap;
    int* packed = c.D + h->n_cap;
    if (g.n > 0) {
        const int nblk = (g.n + 127) / 128;
        k_slic_init<<<nblk, 128, 0, s>>>(g, h->d_lab, h->d_depth, c);
        for (int r = 0; r < rounds; ++r) {
            SL_TRY(cudaMemsetAsync(h->d_cnt, 0, (nb + 1) * 4, s));
            k_slic_bin_count<<<nblk, 128, 0, s>>>(g, c, h->d_cnt);
            k_slic_bin_scan<<<1, 1024, 0, s>>>((int)nb, h->d_cnt, h->d_start, h->d_cursor);
            k_slic_bin_fill<<<nblk, 128, 0, s>>>(g, c, h->d_cursor, h->d_items);
            k_slic_assign<<<dim3((cols + 127) / 128, rows), 128, 0, s>>>(g, h->d_lab, c, h->d_start, h->d_items, h->d_labels);
            k_slic_update<<<(g.n + 3) / 4, 128, 0, s>>>(g, h->d_lab, h->d_depth, h->d_labels, c);
        }
        k_slic_pack_centers<<<nblk, 128, 0, s>>>(g.n, c, packed);
    }
    k_slic_export<<<(int)((px + 255) / 256), 256, 0, s>>>((int)px, h->d_labels, labels_out ? h->d_out64 : nullptr, labels16_out ? h->d_out16 : nullptr);
    SL_TRY(cudaGetLastError());
    if (labels_out) SL_TRY(cudaMemcpy2DAsync(labels_out, labels_step, h->d_out64, (size_t)cols * 8, (size_t)cols * 8, rows, cudaMemcpyDeviceToHost, s));
    if (labels16_out) SL_TRY(cudaMemcpy2DAsync(labels16_out, labels16_step, h->d_out16, (size_t)cols * 2, (size_t)cols * 2, rows, cudaMemcpyDeviceToHost, s));
    if (centers_out && g.n > 0) SL_TRY(cudaMemcpyAsync(centers_out, packed, (size_t)g.n * 7 * sizeof(int), cudaMemcpyDeviceToHost, s));
    SL_TRY(cudaStreamSynchronize(s));
    return ORBX_OK;
}

}  // extern "C"
