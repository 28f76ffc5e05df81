// tma.cuh -- Tensor Memory Accelerator plumbing for the stencil kernels (sm_90+ instruction set, compiled here for sm_100a):
// tiled tensor maps over (x, y, frame) views of 8-bit images, bulk tensor loads into shared memory completing on an mbarrier.
// The stencil stages are instruction-issue bound, so the point of TMA here is to take the whole copy loop (address arithmetic +
// 8-cycle LDGSTS per word) out of the issue slots: one elected lane issues ONE instruction per tile.
#pragma once
#include <cuda.h>            // CUtensorMap + enums only; the encoder is fetched through cudaGetDriverEntryPoint (no -lcuda)
#include <cuda_runtime.h>
#include <stdint.h>

// ---- host: tensor map over a batch of pitched 8-bit images -------------------------------------------------------------------
typedef CUresult (*orbx_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline orbx_tmap_encode_fn orbx_tmap_encoder() {
    static orbx_tmap_encode_fn fn = [] {
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (orbx_tmap_encode_fn)p;
    }();
    return fn;
}
// base: first pixel of frame 0 (16-byte aligned); pitch / fstride in bytes (multiples of 16); box = (bw, bh, 1), bw a multiple of 16, bw, bh <= 256.
// Elements of a box that fall outside [0,w) x [0,h) x [0,frames) are written as zeros.
static inline bool orbx_tmap_image(CUtensorMap* m, const void* base, int w, int h, long long frames, long long pitch, long long fstride, int bw, int bh) {
    orbx_tmap_encode_fn enc = orbx_tmap_encoder();
    if (!enc || ((uintptr_t)base & 15) || (pitch & 15) || (fstride & 15) || (bw & 15) || bw > 256 || bh > 256 || bw <= 0 || bh <= 0) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)(frames > 0 ? frames : 1)};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)fstride};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u}, es[3] = {1u, 1u, 1u};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- device -------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// box of the tensor map at element coordinates (x, y, z) -> dst (128-byte aligned shared memory), completion counted in bytes on bar
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}
