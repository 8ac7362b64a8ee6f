// Standalone tcgen05 probe: one CTA, one (or 4) MMA(s) M=128 N=32 K=8 tf32, several operand layouts.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstring>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(lt & 7) << 61;
    return d;
}

// mode 0: A K-major no-swizzle, B K-major no-swizzle
// mode 1: A MN-major SW128,     B K-major no-swizzle
// mode 2: A MN-major no-swizzle (interleave), B K-major no-swizzle
__global__ void probe(const float* A /*[128][8] m-major rows: A[m*8+k]*/, const float* Bm /*[32][8]*/, float* D /*[128][32]*/, int mode,
                      uint32_t* dbg) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* sa = reinterpret_cast<float*>(base);          // up to 4 KB
    float* sb = sa + 2048;                               // 8 KB later
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 4096; i += 128) sa[i] = 0.f;
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // fill A (M=128, K=8)
    for (int i = tid; i < 128 * 8; i += 128) {
        const int m = i / 8, k = i % 8;
        const float v = A[i];
        int off;
        if (mode == 0) {
            // K-major no swizzle: [k_core(2)][m_core(16)][8 rows][4]: LBO(k cores) = 16*128 B, SBO (m cores) = 128 B
            off = (((k / 4) * 16 + (m / 8)) * 8 + (m % 8)) * 4 + (k % 4);
        } else if (mode == 1) {
            // MN-major SW128: [m_group(4)][8 k rows][32 m] with 16B chunk XOR k
            const int mg = m / 32, ml = m % 32, ch = ml / 4;
            off = (mg * 8 + k) * 32 + (((ch ^ k) << 2) | (ml & 3));
        } else {
            // MN-major interleave: core = [8 k][4 m] (16 B rows); [m_core(32)][8 k][4 m]: SBO(m cores)=128 B, LBO(k groups)=n/a
            off = ((m / 4) * 8 + k) * 4 + (m % 4);
        }
        sa[off] = v;
    }
    // fill B (N=32, K=8): K-major no swizzle [k_core(2)][n_core(4)][8][4]
    for (int i = tid; i < 32 * 8; i += 128) {
        const int n = i / 8, k = i % 8;
        sb[(((k / 4) * 4 + (n / 8)) * 8 + (n % 8)) * 4 + (k % 4)] = Bm[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    if (tid == 0) dbg[0] = tmem_d;
    const uint32_t a_major = (mode == 0) ? 0u : 1u;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (a_major << 15) | (0u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    if (tid == 0) {
        uint64_t ad;
        if (mode == 0) ad = make_desc(smem_u32(sa), 16 * 128, 128, 0);
        else if (mode == 1) ad = make_desc(smem_u32(sa), 1024, 4096, 2);
        else ad = make_desc(smem_u32(sa), 4096, 128, 0);
        const uint64_t bd = make_desc(smem_u32(sb), 4 * 128, 128, 0);
        dbg[1] = (uint32_t)ad; dbg[2] = (uint32_t)(ad >> 32); dbg[3] = (uint32_t)bd; dbg[4] = (uint32_t)(bd >> 32); dbg[5] = idesc;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(0u)
            : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    // wait
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(smem_u32(&mbar)), "r"(0u)
                : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[tid * 32 + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(32) : "memory");
}

int main() {
    std::vector<float> A(128 * 8), B(32 * 8), D(128 * 32), ref(128 * 32);
    for (int m = 0; m < 128; ++m)
        for (int k = 0; k < 8; ++k) A[m * 8 + k] = (float)((m * 3 + k * 5) % 17 - 8);
    for (int n = 0; n < 32; ++n)
        for (int k = 0; k < 8; ++k) B[n * 8 + k] = (float)((n * 7 + k * 2) % 13 - 6);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
            float s = 0;
            for (int k = 0; k < 8; ++k) s += A[m * 8 + k] * B[n * 8 + k];
            ref[m * 32 + n] = s;
        }
    float *dA, *dB, *dD;
    uint32_t* dbg;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dbg, 64);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    for (int mode = 0; mode < 3; ++mode) {
        cudaMemset(dD, 0xff, D.size() * 4);
        probe<<<1, 128, 32 * 1024>>>(dA, dB, dD, mode, dbg);
        cudaError_t e = cudaDeviceSynchronize();
        uint32_t h[8];
        cudaMemcpy(h, dbg, 32, cudaMemcpyDeviceToHost);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, zeros = 0;
        for (size_t i = 0; i < D.size(); ++i) { bad += D[i] != ref[i]; zeros += D[i] == 0.f; }
        printf("mode %d: err=%s tmem=0x%08x adesc=%08x_%08x bdesc=%08x_%08x idesc=%08x mismatches=%d zeros=%d\n", mode, cudaGetErrorString(e),
               h[0], h[2], h[1], h[4], h[3], h[5], bad, zeros);
        printf("   D[0][0..7]  :"); for (int j = 0; j < 8; ++j) printf(" %g", D[j]); printf("\n   ref[0][0..7]:");
        for (int j = 0; j < 8; ++j) printf(" %g", ref[j]);
        printf("\n   D[1][0..7]  :"); for (int j = 0; j < 8; ++j) printf(" %g", D[32 + j]); printf("\n   ref[1][0..7]:");
        for (int j = 0; j < 8; ++j) printf(" %g", ref[32 + j]);
        printf("\n   D[37][0..7] :"); for (int j = 0; j < 8; ++j) printf(" %g", D[37 * 32 + j]); printf("\n   ref[37][0..7]:");
        for (int j = 0; j < 8; ++j) printf(" %g", ref[37 * 32 + j]);
        printf("\n");
    }
    return 0;
}
