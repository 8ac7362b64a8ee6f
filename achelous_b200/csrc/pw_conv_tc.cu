// Pointwise convolution / Linear on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate.
//
//   OUT[O][P] = W[O][K] * X[K][P]  per frame, same prologue/epilogue contract as ach_pw_conv (AchPwConv).
//
// The SIMT kernel is instruction-issue bound (ncu: fma pipe ~35 %, issue slots ~70 %): one FFMA does 32 MACs
// per issue slot, one tcgen05.mma (M=128, N=128, K=8, kind::tf32) does 131072.  To keep the reference's fp32
// accuracy (north-star tolerance 1e-3, observed ~1e-6) every operand is split into two TF32 terms,
// x = x_hi + x_lo, w = w_hi + w_lo (x_hi = round-to-tf32(x), x_lo = x - x_hi), and three MMAs accumulate
// x_hi w_hi + x_lo w_hi + x_hi w_lo in the fp32 TMEM accumulator ("3xTF32"; the dropped x_lo w_lo term is
// ~2^-22 relative).
//
// Mapping: UMMA M = 128 pixels (TMEM lanes), N = NT <= 128 outputs (TMEM columns), K = 8 per instruction.
// Both operands are K-major, no swizzle ("interleaved" 8 x 16 B core matrices):
//   A = X^T tile [128 px][32 k]: smem [k-core (4 k)][m-core (8 px)][8 px rows][4 k].  Thread t owns pixel t:
//       it gathers 4 consecutive k of its pixel with 4 warp-coalesced 128-byte loads and writes them as ONE
//       16-byte shared store at float offset kcore*512 + t*4 - consecutive threads hit consecutive 16-byte
//       slots, so the transposition costs no bank conflicts (descriptor: LBO = 2048 B between k cores,
//       SBO = 128 B between pixel cores).  An MN-major A descriptor would avoid the transposition but was
//       measured to yield zeros for kind::tf32 on this part (tools/probe/tc_probe.cu, modes 1-2).
//   B = W tile [NT outputs][32 k], pre-packed on the device by ach_pack_pw_tc into exactly the shared-memory
//       image [k-core][n-core][8 rows][4 k] so the kernel copies it linearly
//       (descriptor: LBO = NT/8 * 128 B between k cores, SBO = 128 B between n cores).
// One CTA = 256 threads = one 128-pixel x NT-output tile.  Thread t owns pixel t % 128; the two thread
// halves split the k-cores of every 16-wide K chunk on the way in and the TMEM columns on the way out.
// K is consumed in chunks of 16 through a single shared-memory stage (33 KB): all threads load + split +
// store the chunk, one thread issues the 6 MMAs and commits them to an mbarrier, everybody waits for the
// commit before refilling; 4-6 CTAs are resident per SM so loads of one CTA overlap MMAs/epilogues of others.
// LayerNorm prologue: because thread = pixel, the (shifted) sum / sum of squares of the pixel's channels
// accumulate in registers while the chunks stream by; the MMA runs on the raw x and the epilogue applies
//   LN(x) . w = rstd * (x . w - mean * sum_k w)      (wsum = row sums of the folded weights, from the host)
// so the activations are read exactly once.
// Epilogue: rolled loop of tcgen05.ld 32x32b.x8 (thread = pixel, registers = outputs; kept small on purpose:
// a fully unrolled 128-output epilogue with erf-GELU thrashed the instruction cache - ncu stall_no_instruction
// 6.4 per issue), folded scale/bias, activation, layer-scale + residual, coalesced 128-byte stores.
#include "common.cuh"

namespace ach {

constexpr int TC_M = 128;   // pixels per tile
constexpr int TC_KC = 16;   // K per shared-memory chunk (2 MMA K-steps of 8)

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

template <int NT>
__global__ void __launch_bounds__(256) pw_conv_tc_kernel(const AchPwConv p, const float* __restrict__ w_hi,
                                                         const float* __restrict__ w_lo, const float* __restrict__ wsum,
                                                         int n_kchunks) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* a_hi = reinterpret_cast<float*>(smem_raw);                   // [4 k-cores][128 px][4]  8 KB
    float* a_lo = a_hi + TC_KC * TC_M;                                  // 8 KB
    float* b_hi = a_lo + TC_KC * TC_M;                                  // [4 k-cores][NT][4]
    float* b_lo = b_hi + NT * TC_KC;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_ln[2][TC_M][2];

    const int tid = threadIdx.x, warp = tid >> 5;
    const int px = tid & (TC_M - 1), half = tid >> 7;   // half is warp-uniform
    const int b = blockIdx.z;
    const int p_base = blockIdx.x * TC_M;
    const int o_tile = blockIdx.y;
    const int o_base = o_tile * NT;
    const int K = p.c0 + p.c1;
    const int P = p.P;
    const float* __restrict__ x0 = p.x0 + (long long)b * p.x0_bs;
    const float* __restrict__ x1 = p.x1 ? p.x1 + (long long)b * p.x1_bs : nullptr;

    // ---- one-time setup: TMEM allocation (warp 0), mbarrier init (thread 0)
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(NT < 32 ? 32 : NT) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int pp = p_base + px;
    const bool p_ok = pp < P;
    // LayerNorm running sums, shifted by the pixel's first channel to avoid cancellation
    const float shift = (p.ln && p_ok) ? x0[pp] : 0.f;
    float s1 = 0.f, s2 = 0.f;

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    // instruction descriptor: D=f32, A=B=tf32, both K-major, N=NT, M=128
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(NT >> 3) << 17) |
                               ((uint32_t)(TC_M >> 4) << 24);
    const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo), b_hi_s = smem_u32(b_hi), b_lo_s = smem_u32(b_lo);
    const uint32_t mbar_s = smem_u32(&mbar);
    constexpr uint32_t B_LBO = (NT / 8) * 128, B_SBO = 128;   // K-major, no swizzle
    constexpr uint32_t A_LBO = (TC_M / 8) * 128, A_SBO = 128; // K-major, no swizzle

    uint32_t parity = 0;
    for (int c = 0; c < n_kchunks; ++c) {
        const int k0 = c * TC_KC;
        // ---- A chunk: thread = pixel; this half's 2 of the 4 k-cores (4 channels each)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const int j = half + 2 * jj;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int kk = k0 + j * 4 + e;
                float t = 0.f;
                if (kk < K && p_ok) {
                    t = (kk < p.c0) ? __ldg(x0 + (long long)kk * P + pp) : __ldg(x1 + (long long)(kk - p.c0) * P + pp);
                    const float d = t - shift;
                    s1 += d;
                    s2 = fmaf(d, d, s2);
                }
                v[e] = t;
            }
            float4 h, l;
            h.x = to_tf32(v[0]); h.y = to_tf32(v[1]); h.z = to_tf32(v[2]); h.w = to_tf32(v[3]);
            l.x = v[0] - h.x; l.y = v[1] - h.y; l.z = v[2] - h.z; l.w = v[3] - h.w;
            *reinterpret_cast<float4*>(a_hi + j * (TC_M * 4) + px * 4) = h;
            *reinterpret_cast<float4*>(a_lo + j * (TC_M * 4) + px * 4) = l;
        }
        // ---- B chunk: linear copy of the pre-packed tile (NT*16 floats each for hi and lo)
        {
            const long long blk = ((long long)o_tile * n_kchunks + c) * (NT * TC_KC);
            const float4* gh = reinterpret_cast<const float4*>(w_hi + blk);
            const float4* gl = reinterpret_cast<const float4*>(w_lo + blk);
            constexpr int N4 = NT * TC_KC / 4;   // 128 (NT=32) .. 512 (NT=128) float4 per matrix
#pragma unroll
            for (int i = 0; i < (N4 + 255) / 256; ++i) {
                const int idx = tid + 256 * i;
                if (idx < N4) {
                    reinterpret_cast<float4*>(b_hi)[idx] = __ldg(gh + idx);
                    reinterpret_cast<float4*>(b_lo)[idx] = __ldg(gl + idx);
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA (async proxy)
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < TC_KC / 8; ++ks) {
                const uint64_t ah = make_desc(a_hi_s + ks * 2 * A_LBO, A_LBO, A_SBO, 0);
                const uint64_t al = make_desc(a_lo_s + ks * 2 * A_LBO, A_LBO, A_SBO, 0);
                const uint64_t bh = make_desc(b_hi_s + ks * 2 * B_LBO, B_LBO, B_SBO, 0);
                const uint64_t bl = make_desc(b_lo_s + ks * 2 * B_LBO, B_LBO, B_SBO, 0);
                mma_tf32(tmem_d, ah, bh, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                mma_tf32(tmem_d, al, bh, idesc, 1u);
                mma_tf32(tmem_d, ah, bl, idesc, 1u);
            }
            // arrives on the mbarrier when all MMAs issued so far have completed (implies fence::before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_s) : "memory");
        }
        mbar_wait(mbar_s, parity);
        parity ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- LayerNorm statistics: combine the two halves' partial sums
    float mean = 0.f, rstd = 1.f;
    if (p.ln) {
        s_ln[half][px][0] = s1;
        s_ln[half][px][1] = s2;
        __syncthreads();
        const float t1 = (s_ln[0][px][0] + s_ln[1][px][0]) / (float)K;
        const float t2 = (s_ln[0][px][1] + s_ln[1][px][1]) / (float)K;
        mean = shift + t1;
        rstd = 1.0f / sqrtf(fmaxf(t2 - t1 * t1, 0.f) + p.ln_eps);
    }

    // ---- epilogue: thread = pixel (TMEM lane 32*(warp%4) + lane); this half's NT/2 columns, 8 at a time
    const uint32_t t_lane = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    constexpr int NH = NT / 2;
#pragma unroll 1
    for (int n0 = half * NH; n0 < (half + 1) * NH; n0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(t_lane + (uint32_t)n0)
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int o = o_base + n0 + j;
            if (o < p.O && p_ok) {
                float acc = __uint_as_float(r[j]);
                if (p.ln) acc = rstd * fmaf(-mean, wsum[o], acc);
                const float s = p.scale ? p.scale[o] : 1.f;
                const float bi = p.bias ? p.bias[o] : 0.f;
                const float pb = p.pbias ? p.pbias[(long long)b * p.O + o] : 0.f;
                float y = apply_act(fmaf(s, acc + pb, bi), p.act);
                if (p.res) {
                    const float g = p.gamma ? p.gamma[o] : 1.f;
                    y = fmaf(g, y, p.res[(long long)b * p.res_bs + (long long)o * P + pp]);
                }
                p.out[(long long)b * p.out_bs + (long long)o * P + pp] = y;
            }
        }
    }

    // ---- teardown
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(NT < 32 ? 32 : NT) : "memory");
    }
}

// ---- device-side packing of K-major [K][ldw] weights into hi/lo UMMA tiles
__global__ void __launch_bounds__(256) pack_pw_tc_kernel(const float* __restrict__ wt, int K, int O, int ldw, int NT, int n_kchunks,
                                                         float* __restrict__ hi, float* __restrict__ lo, long long total) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int blk_elems = NT * TC_KC;
    const long long blk = i / blk_elems;
    int r = (int)(i - blk * blk_elems);
    const int e = r & 3;  r >>= 2;          // k within core
    const int row = r & 7; r >>= 3;         // n within core
    const int ncore = r % (NT / 8);
    const int kcore = r / (NT / 8);
    const int o_tile = (int)(blk / n_kchunks), c = (int)(blk % n_kchunks);
    const int o = o_tile * NT + ncore * 8 + row;
    const int k = c * TC_KC + kcore * 4 + e;
    float w = 0.f;
    if (o < O && k < K) w = wt[(long long)k * ldw + o];
    const float h = to_tf32(w);
    hi[i] = h;
    lo[i] = w - h;
}

static int tc_tile_n(int O) { return O <= 32 ? 32 : (O <= 64 ? 64 : 128); }

template <int NT>
static int launch_tc(const AchPwConv& p, const float* w_hi, const float* w_lo, const float* wsum, cudaStream_t st) {
    const int K = p.c0 + p.c1;
    const int n_kchunks = cdiv(K, TC_KC);
    constexpr size_t smem = 2 * TC_KC * TC_M * 4 + 2 * (size_t)NT * TC_KC * 4;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(pw_conv_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    dim3 grid(cdiv(p.P, TC_M), cdiv(p.O, NT), p.B);
    pw_conv_tc_kernel<NT><<<grid, 256, smem, st>>>(p, w_hi, w_lo, wsum, n_kchunks);
    return check_launch("ach_pw_conv_tc");
}

}  // namespace ach

extern "C" long long ach_pack_pw_tc_elems(int K, int O) {
    using namespace ach;
    const int NT = tc_tile_n(O);
    return (long long)cdiv(O, NT) * cdiv(K, TC_KC) * NT * TC_KC;
}

extern "C" int ach_pack_pw_tc(const float* wt, int K, int O, int ldw, float* w_hi, float* w_lo, void* stream) {
    using namespace ach;
    ACH_REQUIRE(wt && w_hi && w_lo && K > 0 && O > 0 && ldw >= O, "ach_pack_pw_tc: bad args");
    const int NT = tc_tile_n(O);
    const long long total = ach_pack_pw_tc_elems(K, O);
    pack_pw_tc_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(wt, K, O, ldw, NT, cdiv(K, TC_KC), w_hi, w_lo, total);
    return check_launch("ach_pack_pw_tc");
}

extern "C" int ach_pw_conv_tc(const AchPwConv* pp, const float* w_hi, const float* w_lo, const float* wsum, void* stream) {
    using namespace ach;
    const AchPwConv& p = *pp;
    ACH_REQUIRE(p.x0 && p.out && w_hi && w_lo, "ach_pw_conv_tc: null x0/out/weights");
    ACH_REQUIRE(p.B > 0 && p.O > 0 && p.P > 0 && p.c0 > 0 && p.c1 >= 0, "ach_pw_conv_tc: bad dims");
    ACH_REQUIRE((p.c1 == 0) == (p.x1 == nullptr), "ach_pw_conv_tc: x1/c1 mismatch");
    ACH_REQUIRE(p.P % 4 == 0, "ach_pw_conv_tc: P=%d must be a multiple of 4", p.P);
    ACH_REQUIRE(aligned16(p.x0) && aligned16(p.x1) && aligned16(w_hi) && aligned16(w_lo), "ach_pw_conv_tc: views must be 16-byte aligned");
    ACH_REQUIRE(p.x0_bs % 4 == 0 && p.x1_bs % 4 == 0, "ach_pw_conv_tc: batch strides must be multiples of 4 elements");
    ACH_REQUIRE(!p.reduce_max && p.wt_bs == 0, "ach_pw_conv_tc: reduce_max / per-frame weights use ach_pw_conv");
    ACH_REQUIRE(p.B <= 65535, "ach_pw_conv_tc: B too large");
    ACH_REQUIRE(!p.ln || wsum, "ach_pw_conv_tc: the LayerNorm prologue needs wsum (row sums of the folded weights)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (tc_tile_n(p.O)) {
        case 32: return launch_tc<32>(p, w_hi, w_lo, wsum, st);
        case 64: return launch_tc<64>(p, w_hi, w_lo, wsum, st);
        default: return launch_tc<128>(p, w_hi, w_lo, wsum, st);
    }
}
