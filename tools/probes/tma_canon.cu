// tma_canon.cu -- the CUDA programming guide's canonical TMA example (libcu++ barrier + cuda::device::experimental wrappers) bent step by step
// towards the ORB use (u8 pixels, small unaligned boxes, 3-D), to tell an environment problem from a bug in csrc/tma.cuh.
// usage: tma_canon <variant>   0 = int32 2-D 64x64 (the guide's), 1 = u8 2-D 64x64, 2 = u8 2-D 48x44 on 640x480, 3 = u8 3-D 48x44x1
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cuda/barrier>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int rank, int bytes, int x, int y, int z, uint8_t* out) {
    __shared__ alignas(128) uint8_t smem_buffer[16384];
    #pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        if (rank == 2) cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        else cde::cp_async_bulk_tensor_3d_global_to_shared(&smem_buffer, &tensor_map, x, y, z, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem_buffer[i];
}
int main(int argc, char** argv) {
    const int v = argc > 1 ? atoi(argv[1]) : 0;
    const int es = v == 0 ? 4 : 1;
    const int GW = v >= 2 ? 640 : 1024, GH = v >= 2 ? 480 : 1024, F = 3, BW = v >= 2 ? 48 : 64, BH = v >= 2 ? 44 : 64, rank = v == 3 ? 3 : 2;
    std::vector<uint8_t> h((size_t)GW * GH * es * F); for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    CUtensorMap m; uint64_t size[3] = {(uint64_t)GW, (uint64_t)GH, (uint64_t)F}; uint64_t stride[2] = {(uint64_t)GW * es, (uint64_t)GW * es * GH}; uint32_t box[3] = {(uint32_t)BW, (uint32_t)BH, 1}; uint32_t est[3] = {1, 1, 1};
    CUresult r = enc(&m, v == 0 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, size, stride, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int bytes = BW * BH * es, x = getenv("PROBE_X") ? atoi(getenv("PROBE_X")) : 64, y = 128, z = rank == 3 ? 2 : 0;
    uint8_t* dout; cudaMalloc(&dout, bytes);
    kernel<<<1, 128>>>(m, rank, bytes, x, y, z, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d: encode rc %d, run: %s\n", v, (int)r, cudaGetErrorString(e));
    if (e == cudaSuccess) { std::vector<uint8_t> o(bytes); cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost); int bad = 0;
        for (int yy = 0; yy < BH; ++yy) for (int xx = 0; xx < BW * es; ++xx) bad += o[yy * BW * es + xx] != h[(size_t)z * GW * es * GH + (size_t)(y + yy) * GW * es + x * es + xx]; printf("mismatches %d\n", bad); }
    return 0;
}
