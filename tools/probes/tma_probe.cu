// tma_probe.cu -- stand-alone check of the TMA plumbing in csrc/tma.cuh, one variant per process (a fault kills the context):
//   tma_probe <mode>   0 = 1-D bulk copy (no tensor map), 1 = 3-D map as __grid_constant__ parameter, 2 = 3-D map read from global memory,
//                      3 = 2-D map as parameter, 4 = mbarrier only
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I amos-slam_b200/csrc -o tools/probes/tma_probe tools/probes/tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tma.cuh"
#include <dlfcn.h>

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void k_probe(const __grid_constant__ CUtensorMap pm, const CUtensorMap* gm, const uint8_t* img, int fence, int mode, int bw, int bh, int x0, int y0, int z, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16384);
    if (threadIdx.x == 0) { mbar_init(bar, 1); if (fence & 1) mbar_fence_init(); if (fence & 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
    __syncwarp();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, mode == 4 ? 0u : (uint32_t)(bw * bh));
        if (mode == 0) bulk_load_1d(sm, img, (uint32_t)(bw * bh), bar);
        else if (mode == 1) tma_load_3d(sm, &pm, x0, y0, z, bar);
        else if (mode == 2) tma_load_3d(sm, gm, x0, y0, z, bar);
        else if (mode == 3) tma_load_2d(sm, &pm, x0, y0, bar);
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < bw * bh; i += 32) out[i] = sm[i];
}

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 1;
    const int W = 640, H = 480, F = 3, pitch = 640; const int bw = getenv("PROBE_BW") ? atoi(getenv("PROBE_BW")) : 48, bh = getenv("PROBE_BH") ? atoi(getenv("PROBE_BH")) : 44, fence = getenv("PROBE_FENCE") ? atoi(getenv("PROBE_FENCE")) : 1;
    std::vector<uint8_t> img((size_t)pitch * H * F);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t* d; cudaMalloc(&d, img.size()); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    CUtensorMap m;
    if (mode == 3) {
        orbx_tmap_encode_fn enc = orbx_tmap_encoder();
        const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H * F}; const cuuint64_t strides[1] = {(cuuint64_t)pitch};
        const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es[2] = {1u, 1u};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode2d failed %d\n", (int)r); return 1; }
    } else if (!orbx_tmap_image(&m, d, W, H, F, pitch, (long long)pitch * H, bw, bh)) { printf("encode failed\n"); return 1; }
    { const unsigned* w = reinterpret_cast<const unsigned*>(&m); printf("desc:"); for (int i = 0; i < 32; ++i) printf(" %08x", w[i]); printf("\n"); }
    if (getenv("PROBE_DLSYM")) {
        void* lib = dlopen("libcuda.so.1", RTLD_NOW); orbx_tmap_encode_fn f2 = lib ? (orbx_tmap_encode_fn)dlsym(lib, "cuTensorMapEncodeTiled") : nullptr;
        printf("dlsym encoder %p vs entry point %p\n", (void*)f2, (void*)orbx_tmap_encoder());
        if (f2 && mode != 3) { CUtensorMap m2; const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F}; const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * H};
            const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u}, es[3] = {1u, 1u, 1u};
            CUresult r = f2(&m2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            const unsigned* w = reinterpret_cast<const unsigned*>(&m2); printf("rc %d desc2:", (int)r); for (int i = 0; i < 32; ++i) printf(" %08x", w[i]); printf("\n"); m = m2; }
    }
    CUtensorMap* dm; cudaMalloc(&dm, sizeof(m)); cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
    uint8_t* dout; cudaMalloc(&dout, bw * bh);
    std::vector<uint8_t> out(bw * bh);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    const int xs[3] = {12, 601, 333}, ys[3] = {13, 440, 200}, zs[3] = {0, 2, 1};
    for (int t = 0; t < 3; ++t) {
        cudaMemset(dout, 0xEE, bw * bh);
        k_probe<<<1, 32, 16384 + 64>>>(m, dm, d, fence, mode, bw, bh, xs[t], ys[t], mode == 3 ? 0 : zs[t], dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d case %d: %s\n", mode, t, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(out.data(), dout, bw * bh, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
            const int gx = xs[t] + x, gy = ys[t] + y;
            uint8_t want;
            if (mode == 0) want = img[y * bw + x];
            else if (mode == 4) want = out[y * bw + x];
            else if (mode == 3) want = (gx < W && gy < H * F) ? img[(size_t)gy * pitch + gx] : 0;
            else want = (gx < W && gy < H) ? img[(size_t)zs[t] * pitch * H + (size_t)gy * pitch + gx] : 0;
            bad += out[y * bw + x] != want;
        }
        printf("mode %d case %d: %d mismatches\n", mode, t, bad);
    }
    return 0;
}
