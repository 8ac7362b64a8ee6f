// MobileViT token self-attention on the tensor cores (mobilevit.py:48-73 inside MobileViTBlock :134-165):
// softmax(Q K^T / sqrt(d)) V per (frame, 2x2 patch-position group, head), d = 8, N = H*W/4 tokens - same contract as
// mvit_attn.cu (the CUDA-core kernel, which stays the default: see the measurement there), nothing unfolded, qkv
// channel-major.  Opt-in with ACH_MVIT_TC=1; parity-tested in both modes (tests/test_kernels_gpu.py::test_mvit_attention).
//
// One CTA = 128 queries (thread = query = TMEM lane) of one (frame, group, head); both contractions are 3xTF32
// tcgen05.mma with the A operand in tensor memory:
//   S chunk (128 x <=128 keys) = Q . K^T      A = Q hi/lo (8 + 8 TMEM columns, written once), B = K hi/lo tiles in shared memory
//   O (128 x 8)              += P . V          A = P hi/lo of 16 keys (tcgen05.st, two stages), B = V^T hi/lo tiles (N = 16, 8 used)
// Two passes over the keys so that O never has to be rescaled: pass 1 recomputes nothing but the row maxima (the S MMAs
// cost ~nothing), pass 2 recomputes each S chunk, turns it into P = exp(S - max) 16 keys at a time (registers), sums the
// row, and feeds P straight back to the tensor core - the N x N score / probability tensors exist only as one 128-column
// TMEM chunk and a 32-column P stage.  Padded keys (N is rounded up to 16) are masked out of the max and get P = 0.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace ach {

constexpr int MA_D = 8;          // head dimension
constexpr int MA_SC = 128;       // keys per S chunk (TMEM columns [0, 128))
constexpr int MA_O = 128;        // O accumulator columns [128, 144)
constexpr int MA_Q = 144;        // Q operand: hi [144, 152), lo [152, 160)
constexpr int MA_P = 160;        // P stages: [160, 192), [192, 224): hi 16 | lo 16
constexpr int MA_TMEM = 256;

__global__ void __launch_bounds__(128) mvit_attn_tc_kernel(const float* __restrict__ qkv, long long qkv_bs, float* __restrict__ out,
                                                           long long out_bs, int heads, int H, int W, float scale, int N, int Npad) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* k_hi = reinterpret_cast<float*>(smem_raw);      // [2 k-cores][Npad/8][8][4]
    float* k_lo = k_hi + Npad * MA_D;
    float* v_hi = k_lo + Npad * MA_D;                      // [Npad/16 chunks][4 k-cores][2 row-cores][8][4]  (V^T, rows 8..15 zero)
    float* v_lo = v_hi + Npad * 16;
    __shared__ __align__(8) uint64_t mbar_s, mbar_p[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int hw2 = W / 2, P = H * W;
    const int head = blockIdx.x % heads, g = blockIdx.x / heads;      // patch position: ph = g >> 1, pw = g & 1
    const int ph = g >> 1, pw = g & 1;
    const int b = blockIdx.z, qt = blockIdx.y;
    const int inner = heads * MA_D;
    const float* qb = qkv + (long long)b * qkv_bs + (long long)(head * MA_D) * P;
    const float* kb = qb + (long long)inner * P;
    const float* vb = kb + (long long)inner * P;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(MA_TMEM) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar_s)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar_p[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar_p[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- K and V^T of this (frame, group, head) as tf32 hi/lo UMMA tiles (zero for padded keys / padded V rows)
    constexpr int SU = 5;   // loads of SU elements in flight per thread (the un-batched loop paid one L2 round trip per element)
    for (int i0 = tid; i0 < Npad * MA_D; i0 += 128 * SU) {
        float kv[SU], vv[SU];
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            const int i = i0 + u * 128;
            const int dd = i / Npad, n = i - dd * Npad;     // consecutive threads -> consecutive tokens of one channel
            kv[u] = vv[u] = 0.f;
            if (i < Npad * MA_D && n < N) {
                const int pix = (2 * (n / hw2) + ph) * W + 2 * (n % hw2) + pw;
                kv[u] = __ldg(kb + (long long)dd * P + pix);
                vv[u] = __ldg(vb + (long long)dd * P + pix);
            }
        }
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            const int i = i0 + u * 128;
            if (i >= Npad * MA_D) break;
            const int dd = i / Npad, n = i - dd * Npad;
            const float kh = __uint_as_float(__float_as_uint(kv[u]) & 0xffffe000u), vh = __uint_as_float(__float_as_uint(vv[u]) & 0xffffe000u);
            const int ki = ((dd >> 2) * (Npad / 8) + (n >> 3)) * 32 + (n & 7) * 4 + (dd & 3);             // B rows = keys, k = dim
            k_hi[ki] = kh;
            k_lo[ki] = kv[u] - kh;
            const int vi = (n >> 4) * 256 + (((n & 15) >> 2) * 2 + 0) * 32 + dd * 4 + (n & 3);           // B rows = dims, k = key
            v_hi[vi] = vh;
            v_lo[vi] = vv[u] - vh;
            v_hi[vi + 32] = 0.f;                                                                         // row-core 1: dims 8..15 (padding)
            v_lo[vi + 32] = 0.f;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base_s;
    const uint32_t t_lane = tm + ((uint32_t)(warp * 32) << 16);
    const uint32_t khi_s = smem_u32(k_hi), klo_s = smem_u32(k_lo), vhi_s = smem_u32(v_hi), vlo_s = smem_u32(v_lo);
    const uint32_t mb_s = smem_u32(&mbar_s);

    // ---- Q row of this thread's query -> tensor memory (A operand of every S MMA)
    // scores are kept in log2 units (scale * log2 e folded into Q): P = 2^(S - max) is one MUFU.EX2 per key instead of a
    // ~15-instruction expf - with d = 8 the exponentials, not the dot products, were 2/3 of the instructions
    const float scale2 = scale * 1.4426950408889634f;
    const int q = qt * 128 + tid;
    const bool valid = q < N;
    const int pix = valid ? (2 * (q / hw2) + ph) * W + 2 * (q % hw2) + pw : 0;
    {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int dd = 0; dd < MA_D; ++dd) {
            const float v = valid ? __ldg(qb + (long long)dd * P + pix) * scale2 : 0.f;
            hi[dd] = __float_as_uint(v) & 0xffffe000u;
            lo[dd] = __float_as_uint(v - __uint_as_float(hi[dd]));
        }
        tmem_st8(t_lane + MA_Q, hi);
        tmem_st8(t_lane + MA_Q + 8, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();

    const int n_chunks = (Npad + MA_SC - 1) / MA_SC;
    uint32_t s_commits = 0;
    // S chunk c = Q . K[c*128 ..]^T into TMEM columns [0, width): issued by thread 0, everybody waits for it
    auto s_chunk = [&](int c) {
        const int width = min(MA_SC, Npad - c * MA_SC);
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = tf32_idesc(width);
            const uint32_t off = (uint32_t)(c * MA_SC / 8) * 128u;
            const uint64_t bh = kmajor_desc(khi_s + off, Npad, 0), bl = kmajor_desc(klo_s + off, Npad, 0);
            mma_tf32_ts(tm, tm + MA_Q, bh, idesc, 0u);
            mma_tf32_ts(tm, tm + MA_Q + 8, bh, idesc, 1u);
            mma_tf32_ts(tm, tm + MA_Q, bl, idesc, 1u);
            tc_commit(mb_s);
        }
        mbar_wait(mb_s, s_commits & 1u);
        ++s_commits;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        return width;
    };

    // ---- pass 1: row maximum over the valid keys
    float m = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < n_chunks; ++c) {
        const int width = s_chunk(c);
        // four 16-column loads in flight, one wait (columns past `width` hold stale scores of an earlier chunk: masked by key index)
#pragma unroll 1
        for (int j0 = 0; j0 < width; j0 += 64) {
            uint32_t r[4][16];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]), "=r"(r[u][4]), "=r"(r[u][5]), "=r"(r[u][6]), "=r"(r[u][7]),
                      "=r"(r[u][8]), "=r"(r[u][9]), "=r"(r[u][10]), "=r"(r[u][11]), "=r"(r[u][12]), "=r"(r[u][13]), "=r"(r[u][14]), "=r"(r[u][15])
                    : "r"(t_lane + (uint32_t)(j0 + 16 * u))
                    : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int jj = j0 + 16 * u + j;
                    if (jj < width && c * MA_SC + jj < N) m = fmaxf(m, __uint_as_float(r[u][j]));
                }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                       // every lane has read the chunk before the next MMA overwrites it
    }

    // ---- pass 2: P = exp(S - m) 16 keys at a time -> tensor memory -> O += P . V
    float l = 0.f;
    uint32_t piece = 0;
    const uint32_t idesc_pv = tf32_idesc(16);
#pragma unroll 1
    for (int c = 0; c < n_chunks; ++c) {
        const int width = s_chunk(c);
#pragma unroll 1
        for (int j0 = 0; j0 < width; j0 += 16, ++piece) {
            uint32_t r[16];
            tmem_ld16(t_lane + (uint32_t)j0, r);
            const int key0 = c * MA_SC + j0;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float p = (key0 + j < N) ? ex2_approx(__uint_as_float(r[j]) - m) : 0.f;
                l += p;
                hi[j] = __float_as_uint(p) & 0xffffe000u;
                lo[j] = __float_as_uint(p - __uint_as_float(hi[j]));
            }
            const uint32_t st = piece & 1u;
            if (piece >= 2u) {                 // the PV MMAs that read this P stage two pieces ago are done
                mbar_wait(smem_u32(&mbar_p[st]), ((piece >> 1) - 1u) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t pcol = (uint32_t)MA_P + st * 32u;
            tmem_st16(t_lane + pcol, hi);
            tmem_st16(t_lane + pcol + 16u, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t voff = (uint32_t)(key0 / 16) * 1024u;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint32_t ah = tm + pcol + (uint32_t)ks * 8u, al = ah + 16u;
                    const uint64_t bh = kmajor_desc(vhi_s + voff, 16, ks), bl = kmajor_desc(vlo_s + voff, 16, ks);
                    mma_tf32_ts(tm + MA_O, ah, bh, idesc_pv, (piece > 0u || ks > 0) ? 1u : 0u);
                    mma_tf32_ts(tm + MA_O, al, bh, idesc_pv, 1u);
                    mma_tf32_ts(tm + MA_O, ah, bl, idesc_pv, 1u);
                }
                tc_commit(smem_u32(&mbar_p[st]));
            }
        }
        // the last piece's barrier ordered every read of this S chunk before the next chunk's MMAs
    }
    {   // all PV MMAs done (a commit covers every earlier MMA of the thread)
        const uint32_t last = piece - 1u;
        mbar_wait(smem_u32(&mbar_p[last & 1u]), (last >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    {
        uint32_t r[16];
        tmem_ld16(t_lane + MA_O, r);
        if (valid) {
            const float inv = 1.0f / l;
            float* ob = out + (long long)b * out_bs + (long long)(head * MA_D) * P + pix;
#pragma unroll
            for (int dd = 0; dd < MA_D; ++dd) ob[(long long)dd * P] = __uint_as_float(r[dd]) * inv;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(MA_TMEM) : "memory");
}

}  // namespace ach

// called by ach_mvit_attention (mvit_attn.cu); returns -1 when the shape is outside this kernel's range
int mvit_attention_tc_launch(const float* qkv, long long qkv_bs, float* out, long long out_bs, int B, int heads, int H, int W, float scale,
                             cudaStream_t st) {
    using namespace ach;
    const int N = (H / 2) * (W / 2);
    const int Npad = (N + 15) & ~15;
    const size_t smem = (size_t)(2 * Npad * MA_D + 2 * Npad * 16) * sizeof(float);
    if (smem > 100 * 1024 || B > 65535) return -1;
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(mvit_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    }
    mvit_attn_tc_kernel<<<dim3(4 * heads, cdiv(N, 128), B), 128, smem, st>>>(qkv, qkv_bs, out, out_bs, heads, H, W, scale, N, Npad);
    return check_launch("ach_mvit_attention");
}
