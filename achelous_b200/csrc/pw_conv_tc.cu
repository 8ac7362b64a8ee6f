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
// Persistent CTAs of 256 threads walk a list of (frame, 128-pixel tile, NT-output tile) work items, so TMEM
// allocation, mbarrier setup and descriptor construction are paid once per CTA, not once per tile.  Thread t
// owns pixel t % 128; the two thread halves split the k-cores of every 16-wide K chunk on the way in and the
// TMEM columns on the way out.  K chunks go through a 2-stage shared-memory ring: all threads load + split +
// store chunk c into stage c & 1, one thread issues its 6 MMAs and commits them to that stage's mbarrier;
// the global loads of chunk c+1 therefore overlap the MMAs of chunk c, and a stage is only refilled after its
// previous MMAs have signalled completion.  3 CTAs are resident per SM (64 KB smem, 128 TMEM columns each).
// LayerNorm prologue: because thread = pixel, the (shifted) sum / sum of squares of the pixel's channels
// accumulate in registers while the chunks stream by; the MMA runs on the raw x and the epilogue applies
//   LN(x) . w = rstd * (x . w - mean * sum_k w)      (wsum = row sums of the folded weights, from the host)
// so the activations are read exactly once.
// Epilogue: rolled loop of tcgen05.ld 32x32b.x16 (thread = pixel, registers = outputs; kept small on purpose:
// a fully unrolled 128-output epilogue with erf-GELU thrashed the instruction cache - ncu stall_no_instruction
// 6.4 per issue), folded scale/bias, activation, layer-scale + residual, coalesced 128-byte stores.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace ach {

template <int NT, int STAGES, int ACT, int HALVES>
__global__ void __launch_bounds__(128 * HALVES, 8 / HALVES) pw_conv_tc_kernel(const AchPwConv p, const float* __restrict__ w_hi,
                                                         const float* __restrict__ w_lo, const float* __restrict__ wsum,
                                                         int n_kchunks, int n_pt, int n_ot, int total_items) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int A_ELEMS = TC_KC * TC_M;     // per hi / lo matrix
    constexpr int B_ELEMS = NT * TC_KC;
    constexpr int STAGE = 2 * A_ELEMS + 2 * B_ELEMS;
    float* stage_base = reinterpret_cast<float*>(smem_raw);   // [2 stages][a_hi | a_lo | b_hi | b_lo]
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_ln[HALVES][TC_M][2];
    __shared__ __align__(16) float4 s_ep[NT];   // per output of the current tile: {scale, scale*wsum, bias + scale*pbias, gamma}

    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int THREADS = 128 * HALVES;   // HALVES thread groups share the k-cores on the way in and the columns on the way out
    const int px = tid & (TC_M - 1), half = tid >> 7;   // half is warp-uniform
    const int K = p.c0 + p.c1;
    const int P = p.P;

    // ---- one-time setup: TMEM allocation (warp 0), mbarrier init (thread 0)
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(NT < 32 ? 32 : NT) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    // instruction descriptor: D=f32, A=B=tf32, both K-major, N=NT, M=128
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(NT >> 3) << 17) |
                               ((uint32_t)(TC_M >> 4) << 24);
    constexpr uint32_t B_LBO = (NT / 8) * 128, B_SBO = 128;   // K-major, no swizzle
    constexpr uint32_t A_LBO = (TC_M / 8) * 128, A_SBO = 128; // K-major, no swizzle
    const uint32_t stage_s = smem_u32(stage_base);
    const uint32_t mbar_s0 = smem_u32(&mbar[0]), mbar_s1 = smem_u32(&mbar[1]);
    uint32_t uses0 = 0u, uses1 = 0u;   // commits issued so far on each stage barrier (identical in every thread)
    const uint32_t t_lane = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);

    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int o_tile = item % n_ot;
        const int pt = (item / n_ot) % n_pt;
        const int b = item / (n_ot * n_pt);
        const int p_base = pt * TC_M;
        const int o_base = o_tile * NT;
        const float* __restrict__ x0 = p.x0 + (long long)b * p.x0_bs;
        const float* __restrict__ x1 = p.x1 ? p.x1 + (long long)b * p.x1_bs : nullptr;
        const int pp = p_base + px;
        const bool p_ok = pp < P;
        // LayerNorm running sums, shifted by the pixel's first channel to avoid cancellation
        const float shift = (p.ln && p_ok) ? x0[pp] : 0.f;
        float s1 = 0.f, s2 = 0.f;

        constexpr int JJ = 4 / HALVES;   // k-cores of a chunk handled by this thread
        // activation loads of chunk c (global -> registers only)
        auto load_a = [&](int c, float (&v)[JJ][4]) {
            const int k0 = c * TC_KC;
            if (k0 + TC_KC <= p.c0 || (k0 >= p.c0 && k0 + TC_KC <= K)) {
                // fast path (the common case): the whole chunk lies inside one source -> one base pointer, constant strides,
                // no per-element bounds / source selection
                const float* __restrict__ src = (k0 < p.c0) ? x0 + (long long)k0 * P + pp : x1 + (long long)(k0 - p.c0) * P + pp;
#pragma unroll
                for (int jj = 0; jj < JJ; ++jj) {
                    const int j = half + HALVES * jj;
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[jj][e] = p_ok ? __ldg(src + (long long)(j * 4 + e) * P) : 0.f;
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < JJ; ++jj) {
                    const int j = half + HALVES * jj;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int kk = k0 + j * 4 + e;
                        float t = 0.f;
                        if (kk < K && p_ok) t = (kk < p.c0) ? __ldg(x0 + (long long)kk * P + pp) : __ldg(x1 + (long long)(kk - p.c0) * P + pp);
                        v[jj][e] = t;
                    }
                }
            }
        };
        // software pipeline: the activation loads of chunk c+1 are issued before chunk c is split / stored / multiplied,
        // so their HBM latency overlaps a whole iteration instead of stalling the shared-memory store that consumes them
        float v[JJ][4], vn[JJ][4];
        load_a(0, v);
        for (int c = 0; c < n_kchunks; ++c) {
            const int k0 = c * TC_KC;
            const int st = (STAGES == 2) ? (c & 1) : 0;
            float* a_hi = stage_base + st * STAGE;
            float* a_lo = a_hi + A_ELEMS;
            float* b_hi = a_lo + A_ELEMS;
            float* b_lo = b_hi + B_ELEMS;
            if (c + 1 < n_kchunks) load_a(c + 1, vn);
            constexpr int N4 = B_ELEMS / 4;   // float4 per weight matrix: 128 (NT=32) .. 512 (NT=128)
            constexpr int NB = (N4 + THREADS - 1) / THREADS;
            float4 wh[NB], wl[NB];
            {
                const long long blk = ((long long)o_tile * n_kchunks + c) * B_ELEMS;
                const float4* gh = reinterpret_cast<const float4*>(w_hi + blk);
                const float4* gl = reinterpret_cast<const float4*>(w_lo + blk);
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int idx = tid + THREADS * i;
                    if (idx < N4) {
                        wh[i] = __ldg(gh + idx);
                        wl[i] = __ldg(gl + idx);
                    }
                }
            }
            // the stage is free once the MMAs of its previous use have completed
            const uint32_t mbar_st = st ? mbar_s1 : mbar_s0;
            const uint32_t uses_st = st ? uses1 : uses0;
            if (uses_st > 0) mbar_wait(mbar_st, (uses_st - 1) & 1);
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) {
                const int j = half + HALVES * jj;
                if (p.ln) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        // padded elements are exactly 0 and must not enter the statistics: predicate once per element
                        const float d = (k0 + j * 4 + e < K && p_ok) ? v[jj][e] - shift : 0.f;
                        s1 += d;
                        s2 = fmaf(d, d, s2);
                    }
                }
                float4 h, l;
                h.x = to_tf32(v[jj][0]); h.y = to_tf32(v[jj][1]); h.z = to_tf32(v[jj][2]); h.w = to_tf32(v[jj][3]);
                l.x = v[jj][0] - h.x; l.y = v[jj][1] - h.y; l.z = v[jj][2] - h.z; l.w = v[jj][3] - h.w;
                *reinterpret_cast<float4*>(a_hi + j * (TC_M * 4) + px * 4) = h;
                *reinterpret_cast<float4*>(a_lo + j * (TC_M * 4) + px * 4) = l;
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const int idx = tid + THREADS * i;
                if (idx < N4) {
                    reinterpret_cast<float4*>(b_hi)[idx] = wh[i];
                    reinterpret_cast<float4*>(b_lo)[idx] = wl[i];
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA (async proxy)
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // (first chunk) previous item's TMEM reads are ordered before the new MMAs
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi_s = stage_s + st * STAGE * 4, a_lo_s = a_hi_s + A_ELEMS * 4;
                const uint32_t b_hi_s = a_lo_s + A_ELEMS * 4, b_lo_s = b_hi_s + B_ELEMS * 4;
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
                // arrives on the stage barrier when all MMAs issued so far have completed
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_st) : "memory");
            }
            if (st) uses1 += 1; else uses0 += 1;
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj)
#pragma unroll
                for (int e = 0; e < 4; ++e) v[jj][e] = vn[jj][e];
        }
        // all MMAs of this item are complete once the last commit has arrived
        {
            const int st = (STAGES == 2) ? ((n_kchunks - 1) & 1) : 0;
            mbar_wait(st ? mbar_s1 : mbar_s0, ((st ? uses1 : uses0) - 1) & 1);
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- per-output epilogue constants of this (frame, output tile) + LayerNorm partial sums -> shared memory
        if (tid < NT) {
            const int o = o_base + tid;
            float4 e = make_float4(0.f, 0.f, 0.f, 1.f);
            if (o < p.O) {
                e.x = p.scale ? p.scale[o] : 1.f;
                e.y = p.ln ? e.x * wsum[o] : 0.f;
                e.z = (p.bias ? p.bias[o] : 0.f) + (p.pbias ? e.x * p.pbias[(long long)b * p.O + o] : 0.f);
                e.w = p.gamma ? p.gamma[o] : 1.f;
            }
            s_ep[tid] = e;
        }
        if (p.ln) {
            s_ln[half][px][0] = s1;
            s_ln[half][px][1] = s2;
        }
        __syncthreads();
        // y = act(rs * (scale*acc) - ms * (scale*wsum) + c)  with rs = rstd, ms = mean*rstd   (rs = 1, ms = 0 without LayerNorm)
        float rs = 1.f, ms = 0.f;
        if (p.ln) {
            const float t1 = (s_ln[0][px][0] + s_ln[HALVES - 1][px][0] * (HALVES - 1)) / (float)K;
            const float t2 = (s_ln[0][px][1] + s_ln[HALVES - 1][px][1] * (HALVES - 1)) / (float)K;
            rs = 1.0f / sqrtf(fmaxf(t2 - t1 * t1, 0.f) + p.ln_eps);
            ms = (shift + t1) * rs;
        }

        // ---- epilogue: thread = pixel (TMEM lane 32*(warp%4) + lane); this half's NT/2 columns, 16 at a time
        constexpr int NH = NT / HALVES;
        float* optr = p.out + (long long)b * p.out_bs + (long long)(o_base + half * NH) * P + pp;
        const float* rptr = p.res ? p.res + (long long)b * p.res_bs + (long long)(o_base + half * NH) * P + pp : nullptr;
        const int o_lim = p.O - o_base;   // valid outputs in this tile
#pragma unroll 1
        for (int n0 = half * NH; n0 < (half + 1) * NH; n0 += 16) {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(t_lane + (uint32_t)n0)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (p.reduce_max) {
                // out (B, O) = max over pixels: warp-shuffle max over the warp's 32 pixels, one atomic per (warp, output)
                const int lane = tid & 31;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 e = s_ep[n0 + j];
                    float y = fmaf(rs * e.x, __uint_as_float(r[j]), fmaf(-ms, e.y, e.z));
                    y = p_ok ? apply_act(y, ACT) : -INFINITY;
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) y = fmaxf(y, __shfl_xor_sync(0xffffffffu, y, sft));
                    if (lane == 0 && n0 + j < o_lim) atomic_max_float(p.out + (long long)b * p.out_bs + o_base + n0 + j, y);
                }
            } else if (p_ok) {
                // residual values first, as 16 independent loads: interleaved with the stores below they would each
                // stall for a full memory round trip (the compiler cannot hoist a load above a possibly aliasing store)
                if (n0 + 16 <= o_lim) {
                    // full block (all but the last block of a ragged tile): no per-output predicate
                    float rr[16];
                    if (rptr) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) rr[j] = rptr[(long long)j * P];
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float4 e = s_ep[n0 + j];
                        float y = fmaf(rs * e.x, __uint_as_float(r[j]), fmaf(-ms, e.y, e.z));
                        y = apply_act(y, ACT);
                        if (rptr) y = fmaf(e.w, y, rr[j]);
                        optr[(long long)j * P] = y;
                    }
                } else {
#pragma unroll 1
                    for (int j = 0; j < 16 && n0 + j < o_lim; ++j) {
                        const float4 e = s_ep[n0 + j];
                        // r[] must stay in registers: select with a compile-time unrolled chain instead of dynamic indexing
                        uint32_t rv = r[0];
#pragma unroll
                        for (int q = 1; q < 16; ++q) rv = (j == q) ? r[q] : rv;
                        float y = fmaf(rs * e.x, __uint_as_float(rv), fmaf(-ms, e.y, e.z));
                        y = apply_act(y, ACT);
                        if (rptr) y = fmaf(e.w, y, rptr[(long long)j * P]);
                        optr[(long long)j * P] = y;
                    }
                }
            }
            optr += (long long)16 * P;
            if (rptr) rptr += (long long)16 * P;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();   // s_ep / s_ln are rewritten and the accumulator is overwritten by the next item
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

template <int NT, int STAGES, int ACT>
static int launch_tc(const AchPwConv& p, const float* w_hi, const float* w_lo, const float* wsum, cudaStream_t st) {
    // narrow tiles (<= 64 TMEM columns) run as 128-thread CTAs, 8 per SM: twice as many independent load / MMA /
    // epilogue pipelines per SM to hide each other's barrier and mbarrier waits; 128-column tiles are capped at 4
    // CTAs per SM by TMEM and keep 256 threads
    constexpr int HALVES = (NT <= 64 && STAGES == 1) ? 1 : 2;   // measured: long-K narrow layers prefer 256 threads (half the loads per thread)
    const int K = p.c0 + p.c1;
    const int n_kchunks = cdiv(K, TC_KC);
    constexpr size_t smem = STAGES * (2 * TC_KC * TC_M * 4 + 2 * (size_t)NT * TC_KC * 4);
    static int ctas_per_wave = 0;
    if (!ctas_per_wave) {
        cudaFuncSetAttribute(pw_conv_tc_kernel<NT, STAGES, ACT, HALVES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int per_sm = tc_ctas_per_sm(pw_conv_tc_kernel<NT, STAGES, ACT, HALVES>, 128 * HALVES, smem, NT < 32 ? 32 : NT);
        ctas_per_wave = sms * (per_sm < 1 ? 1 : per_sm);
    }
    const int n_pt = cdiv(p.P, TC_M), n_ot = cdiv(p.O, NT);
    const long long total = (long long)n_pt * n_ot * p.B;
    ACH_REQUIRE(total < (1LL << 31), "ach_pw_conv_tc: too many tiles");
    const int grid = (int)(total < ctas_per_wave ? total : ctas_per_wave);   // persistent: one wave of resident CTAs
    pw_conv_tc_kernel<NT, STAGES, ACT, HALVES><<<grid, 128 * HALVES, smem, st>>>(p, w_hi, w_lo, wsum, n_kchunks, n_pt, n_ot, (int)total);
    return check_launch("ach_pw_conv_tc");
}

template <int NT>
static int launch_tc_nt(const AchPwConv& p, const float* w_hi, const float* w_lo, const float* wsum, cudaStream_t st) {
    // long K: a second shared-memory stage lets chunk c+1 load while chunk c multiplies; short K: keep the
    // footprint small so that more CTAs (each at a different phase of load / MMA / epilogue) share the SM
    const bool two = p.c0 + p.c1 > 3 * TC_KC;
    switch (p.act) {
        case ACT_NONE: return two ? launch_tc<NT, 2, ACT_NONE>(p, w_hi, w_lo, wsum, st) : launch_tc<NT, 1, ACT_NONE>(p, w_hi, w_lo, wsum, st);
        case ACT_RELU: return two ? launch_tc<NT, 2, ACT_RELU>(p, w_hi, w_lo, wsum, st) : launch_tc<NT, 1, ACT_RELU>(p, w_hi, w_lo, wsum, st);
        case ACT_SILU: return two ? launch_tc<NT, 2, ACT_SILU>(p, w_hi, w_lo, wsum, st) : launch_tc<NT, 1, ACT_SILU>(p, w_hi, w_lo, wsum, st);
        case ACT_GELU: return two ? launch_tc<NT, 2, ACT_GELU>(p, w_hi, w_lo, wsum, st) : launch_tc<NT, 1, ACT_GELU>(p, w_hi, w_lo, wsum, st);
        default: break;
    }
    set_error("ach_pw_conv_tc: activation %d not instantiated", p.act);
    return ACH_ERR_INVALID;
}

int pw_conv_tc_ws_launch(const AchPwConv& p, const float* w_hi, const float* w_lo, const float* wsum, cudaStream_t st);   // pw_conv_tc_ws.cu

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
    ACH_REQUIRE(p.wt_bs == 0, "ach_pw_conv_tc: per-frame weights use ach_pw_conv");
    ACH_REQUIRE(!(p.reduce_max && p.res), "ach_pw_conv_tc: reduce_max excludes a residual");
    ACH_REQUIRE(p.B <= 65535, "ach_pw_conv_tc: B too large");
    ACH_REQUIRE(!p.ln || wsum, "ach_pw_conv_tc: the LayerNorm prologue needs wsum (row sums of the folded weights)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // default: the warp-specialised asynchronous kernel (pw_conv_tc_ws.cu) for every layer; ACH_TC_WS_MINCHUNKS=n keeps the
    // synchronous kernel below for layers with fewer than n K-chunks (A/B switch for tools/op_times.py; large n = all sync)
    static const int ws_min_chunks = getenv("ACH_TC_WS_MINCHUNKS") ? atoi(getenv("ACH_TC_WS_MINCHUNKS")) : 0;
    if (cdiv(p.c0 + p.c1, TC_KC) >= ws_min_chunks) return pw_conv_tc_ws_launch(p, w_hi, w_lo, wsum, st);
    switch (tc_tile_n(p.O)) {
        case 32: return launch_tc_nt<32>(p, w_hi, w_lo, wsum, st);
        case 64: return launch_tc_nt<64>(p, w_hi, w_lo, wsum, st);
        default: return launch_tc_nt<128>(p, w_hi, w_lo, wsum, st);
    }
}
