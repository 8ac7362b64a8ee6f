// TMA (cp.async.bulk.tensor) + mbarrier helpers for the CUDA-core kernels that stage an input tile in shared memory.
#pragma once
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no libcuda link)
#include <cuda_runtime.h>
#include <cstdint>

namespace ach {

typedef CUresult (*TmaEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TmaEncodeTiledFn tma_encode_fn() {
    static TmaEncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TmaEncodeTiledFn>(ptr);
        else
            cudaGetLastError();
        tried = true;
    }
    return fn;
}

// fp32 planes (W, H, C, B) with element strides (1, W, H*W, bs): box (bw, bh, bc, 1), out-of-range elements read as 0.
// Returns false when the view cannot be described (row pitch not a multiple of 16 bytes, unaligned base, no driver entry point).
inline bool tma_map_planes(CUtensorMap* tm, const float* base, int W, int H, int C, int B, long long bs, int bw, int bh, int bc) {
    TmaEncodeTiledFn enc = tma_encode_fn();
    if (!enc || !base || (reinterpret_cast<uintptr_t>(base) & 15u) || (W % 4) != 0 || (bw % 4) != 0) return false;
    const long long plane = (long long)H * W;
    const long long bstride = B > 1 ? bs : plane * C;
    if ((bstride % 4) != 0) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4ull, (cuuint64_t)plane * 4ull, (cuuint64_t)bstride * 4ull};
    const cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t tma_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "TMA_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TMA_WAIT_DONE;\n\t"
        "bra TMA_WAIT_LOOP;\n\t"
        "TMA_WAIT_DONE:\n\t"
        "}\n" ::"r"(mbar), "r"(parity)
        : "memory");
}
// box of a 4-D tensor map -> shared memory, completion on an mbarrier (coordinates may be negative / past the end: zero fill)
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t mbar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(mbar)
                 : "memory");
}
#endif

}  // namespace ach
