// tcgen05 / TMEM / mbarrier helpers shared by the tensor-core kernels (sm_100a inline PTX).
#pragma once
#include <algorithm>

#include "common.cuh"

namespace ach {

constexpr int TC_M = 128;   // UMMA M: pixels per tile (TMEM lanes)
constexpr int TC_KC = 16;   // K per shared-memory chunk (2 MMA K-steps of 8 for kind::tf32)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)(layout_type & 7) << 61; // 0 = no swizzle
    return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(mbar), "r"(parity)
        : "memory");
}


// K-major, no-swizzle operand tile of R rows (pixels or outputs) x 16 k: [4 k-cores][R/8 row-cores][8 rows][4 k]
// -> leading (k-core) byte offset = R/8 * 128, stride (row-core) byte offset = 128; MMA K-step ks starts at ks*2*LBO.
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t smem_addr, int rows, int kstep) {
    const uint32_t lbo = (uint32_t)(rows / 8) * 128u;
    return make_desc(smem_addr + (uint32_t)kstep * 2u * lbo, lbo, 128u, 0u);
}

// instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t tf32_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// D[tmem] (+)= A[tmem] . B[smem]: the A operand (128 lanes x 8 tf32 columns) is read from tensor memory
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

__device__ __forceinline__ void split_store(float* hi, float* lo, const float (&v)[4]) {
    float4 h, l;
    h.x = to_tf32(v[0]); h.y = to_tf32(v[1]); h.z = to_tf32(v[2]); h.w = to_tf32(v[3]);
    l.x = v[0] - h.x; l.y = v[1] - h.y; l.z = v[2] - h.z; l.w = v[3] - h.w;
    *reinterpret_cast<float4*>(hi) = h;
    *reinterpret_cast<float4*>(lo) = l;
}

// Issue pattern for tcgen05.mma / bulk copies from a warp: run the surrounding loop WARP-UNIFORMLY (every lane; warp index and TMEM base
// passed through tc_uniform so that ptxas can prove uniformity and keep descriptors / addresses in uniform registers) and issue
// under `if (tc_elect_one())`.  With `if (lane == 0)` / `if (tid == 0)` as the guard ptxas cannot tell that one lane is active and
// wraps every UTCHMMA in an ELECT + R2UR + vote + branch "waterfall" (~10 instructions per MMA, on the one warp everything waits for):
// ncu on mlp_tc.cu showed a kernel bound by exactly that with the tensor pipe at 25 %.
__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t tc_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

__device__ __forceinline__ void tc_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

// Resident CTAs per SM for a persistent tensor-core kernel, from the function's own attributes (registers, static
// + dynamic shared memory) and its TMEM column allocation.  (cudaOccupancyMaxActiveBlocksPerMultiprocessor was
// measured to under-report for these kernels, which starves a persistent grid.)
template <typename Kernel>
inline int tc_ctas_per_sm(Kernel kernel, int threads, size_t dyn_smem, int tmem_cols) {
    cudaFuncAttributes fa{};
    if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return 1;
    const int regs_per_cta = ((fa.numRegs * threads + 511) / 512) * 512;            // allocation granularity
    int n = regs_per_cta > 0 ? 65536 / regs_per_cta : 1;
    const size_t smem = fa.sharedSizeBytes + dyn_smem + 1024;                       // + per-CTA reserve
    n = (int)std::min<size_t>(n, (227 * 1024) / smem);
    n = std::min(n, 512 / std::max(tmem_cols, 32));
    n = std::min(n, 2048 / threads);
    return std::max(n, 1);
}

}  // namespace ach
