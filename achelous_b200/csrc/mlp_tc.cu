// Fused inverted-bottleneck MLP of the EdgeNeXt encoders on tcgen05 (conv_encoder.py:23-31, sdta_encoder.py:64-73):
//
//   out = res + gamma * ( W2 . gelu( LN(x) . W1 + b1 ) + b2 )              x, res, out: (B, C, P) fp32 planes
//
// as ONE kernel per block.  The two-launch version (ach_pw_conv_tc twice) writes the 4C-wide hidden tensor to HBM and reads it back:
// ncu on bb.s0.0 (C = 32, 80 x 80, B = 64) shows 150 MB written by the first GEMM and 262 MB read by the second for 52 MB of
// input and 52 MB of output - over the 12 blocks of EN-S0 a quarter of the plan's modelled traffic.  Here the hidden tile never
// leaves the SM: it exists 32 columns at a time, in tensor memory.
//
// Per 128-pixel tile (thread = pixel = TMEM lane), 3xTF32 arithmetic exactly as ach_pw_conv_tc (x = hi + lo, w = hi + lo, three MMAs):
//   * GEMM 1 (K = C, N = 32) produces hidden chunk j in a 32-column accumulator; an epilogue group (4 warps) turns it into
//     gelu(rstd * acc - mean * rstd * wsum + b1) hi/lo terms in its 64-column A-operand slot (tcgen05.ld -> registers -> tcgen05.st);
//     GEMM 2 (K = 32, N = C) accumulates the chunk into the output accumulator.  The LayerNorm statistics come from the same
//     shared-memory tile that feeds "XA", the split activations in tensor memory (2C columns, the A operand of every GEMM 1).
//   * The GELU epilogue (~25 instructions and 2 MUFU per hidden value) bounds the kernel, so everything else is arranged to keep
//     the two epilogue groups issuing: a group signals "accumulator drained" right after its tcgen05.ld, which lets the MMA
//     warp issue GEMM 1 of the group's NEXT chunk underneath the GELU arithmetic of the current one; the next tile's XA is staged
//     before the last chunk of a tile, so the chunk stream crosses tile boundaries without a bubble; with two output accumulators
//     (C = 32, 64) the output epilogue (+ b2, * gamma, + residual, coalesced stores) of tile t runs after the first chunk of t + 1.
//     v1 (four groups on one tile, no drain signal): ncu showed 52 % issue utilisation with a quarter of all issued instructions
//     being mbarrier polls - every chunk was a serial round trip group -> MMA warp -> tensor core -> group.
//   * TMEM (512 columns) decides the shape: C <= 48: TWO tiles in flight ("slots"), one group, one MMA warp and one loader warp each,
//     all weight tiles resident in shared memory;  C >= 64: one tile, the two groups alternate chunks, weights streamed through two
//     bulk-copy rings (the source is L2-resident).
//     slot columns: [output accumulator(s) C or 2C | XA 2C | per group 96: accumulator 32, A2 hi 32, A2 lo 32].
//
// The arithmetic (K order of the MMAs, LayerNorm partial sums, epilogue expressions) is the same as in the two-launch path;
// tests/test_kernels_gpu.py::test_mlp_tc checks the kernel against the emulator and against that path.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace ach {

template <int C>
struct MlpCfg {
    static constexpr int NJ = C / 8;                            // hidden chunks of 32 columns (4C / 32)
    static constexpr int KC = C / TC_KC;                        // K chunks of GEMM 1
    static constexpr bool TWO_SLOTS = C <= 48;                  // two tiles in flight with one group each; else one tile, two groups
    static constexpr int SLOTS = TWO_SLOTS ? 2 : 1;
    static constexpr int SG = TWO_SLOTS ? 1 : 2;                // epilogue groups per slot
    static constexpr int R = NJ / SG;                           // chunk rounds per tile
    static constexpr int NBUF = (C == 32 || C == 64) ? 2 : 1;   // output accumulators per slot
    static constexpr int SLOT_COLS = NBUF * C + 2 * C + SG * 96;
    static constexpr bool RESIDENT = TWO_SLOTS;                 // all weight tiles stay in shared memory
    static constexpr int CL = RESIDENT ? 1 : 2;                 // thread-block cluster: streamed weight tiles are multicast to CL CTAs
    static constexpr int S1 = RESIDENT ? NJ : 4;                // W1 stages
    static constexpr int S2 = RESIDENT ? NJ : 3;                // W2 stages
    static constexpr int W_STAGE = C * 64;                      // floats per stage: hi [C x 32] | lo [C x 32]
    static constexpr int THREADS = (8 + 2 + 2) * 32;            // 2 epilogue groups, 2 MMA warps, 2 loader warps
    static constexpr size_t SMEM = (size_t)(SLOTS * C * TC_M + (S1 + S2) * W_STAGE) * sizeof(float);
    static_assert(NJ % SG == 0 && R >= 2, "chunk rounds");
    static_assert(SLOTS * SLOT_COLS <= 512, "TMEM budget");
    static_assert(CL == 1 || CL == 2, "the half-stage multicast protocol is written for pairs");
};

__device__ __forceinline__ void mlp_mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mlp_mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mlp_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(mbar)
                 : "memory");
}
// mbarrier wait with a suspend-time hint: the waiting warp is parked by the hardware until the phase completes (or the hint expires)
// instead of polling - v2's MMA / loader warps spent ~12 % of the SM's issue slots in try_wait / branch / yield loops
__device__ __forceinline__ void mlp_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "MLP_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra MLP_WAIT_DONE;\n\t"
        "bra MLP_WAIT_LOOP;\n\t"
        "MLP_WAIT_DONE:\n\t"
        "}\n" ::"r"(mbar), "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void mlp_wait_tc(uint32_t mbar, uint32_t parity) {   // barrier wait + ordering of the following tcgen05 ops
    mlp_wait(mbar, parity);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// Issue-side helpers for a WARP-UNIFORM MMA loop: every lane runs the loop (so that ptxas keeps descriptors, TMEM addresses and
// barrier addresses in uniform registers - under an `if (lane == 0)` it wraps every UTCHMMA in an ELECT / 4 x R2UR / BRA.ANY
// waterfall, ~10 instructions per MMA on one warp: ncu showed the C = 96 kernel bound by exactly that), one elected lane issues.
__device__ __forceinline__ void mlp_mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t elected) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 e, %5, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(elected)
        : "memory");
}
__device__ __forceinline__ void mlp_commit(uint32_t mbar, uint32_t elected) {
    asm volatile(
        "{\n\t"
        ".reg .pred e;\n\t"
        "setp.ne.b32 e, %1, 0;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}\n" ::"r"(mbar), "r"(elected)
        : "memory");
}

// apply_act(ACT_GELU) (common.cuh: erf by Abramowitz-Stegun 7.1.26 on MUFU.RCP / MUFU.EX2) for two values at once on Blackwell's
// packed fp32 pipe (FFMA2 / FMUL2: two independent IEEE operations per issue slot).  Element-wise the same operations in the same
// order as the scalar form - negations are moved into the constants, which is exact - so the results are bit-identical; the GELU
// epilogue is what bounds this kernel and this halves its FMA / MUL issue slots.
__device__ __forceinline__ float2 mlp_gelu2(float2 v) {
    const float2 h = __fmul2_rn(v, make_float2(0.5f, 0.5f));
    const float2 z = __fmul2_rn(v, make_float2(0.70710678118654752440f, 0.70710678118654752440f));
    const float2 ax = make_float2(fabsf(z.x), fabsf(z.y));
    const float2 d = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), ax, make_float2(1.0f, 1.0f));
    const float2 t = make_float2(rcp_approx(d.x), rcp_approx(d.y));   // arguments >= 1
    // -poly (all coefficient signs flipped): erf = 1 - poly * e = fma(-poly, e, 1)
    float2 np = __ffma2_rn(make_float2(-1.061405429f, -1.061405429f), t, make_float2(1.453152027f, 1.453152027f));
    np = __ffma2_rn(np, t, make_float2(-1.421413741f, -1.421413741f));
    np = __ffma2_rn(np, t, make_float2(0.284496736f, 0.284496736f));
    np = __ffma2_rn(np, t, make_float2(-0.254829592f, -0.254829592f));
    np = __fmul2_rn(np, t);
    float2 q = __fmul2_rn(ax, ax);
    q = __fmul2_rn(q, make_float2(-1.4426950408889634f, -1.4426950408889634f));   // (-ax * ax) * log2(e)
    const float2 e = make_float2(ex2_approx(q.x), ex2_approx(q.y));              // underflow -> 0 -> erf = +-1
    const float2 r = __ffma2_rn(np, e, make_float2(1.0f, 1.0f));
    const float2 erf = make_float2(copysignf(r.x, z.x), copysignf(r.y, z.y));
    return __ffma2_rn(h, erf, h);
}

__device__ __forceinline__ void mlp_commit_mc(uint32_t mbar, uint32_t elected, uint16_t cta_mask) {   // arrive on the barrier at this offset in the masked CTAs
    asm volatile(
        "{\n\t"
        ".reg .pred e;\n\t"
        "setp.ne.b32 e, %1, 0;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %2;\n\t"
        "}\n" ::"r"(mbar), "r"(elected), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void mlp_bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(mbar), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ uint32_t mlp_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void mlp_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct MlpBars {   // per slot
    uint64_t x_full, x_empty, xa_full, xa_free, acc2_full[2], acc2_empty[2];
    uint64_t acc1_full[2], acc1_drained[2], a2_full[2], a2_free[2];   // per group of the slot
};

template <int C>
__global__ void __launch_bounds__(MlpCfg<C>::THREADS, 1)
    mlp_tc_kernel(const AchMlp p, const float* __restrict__ w1_hi, const float* __restrict__ w1_lo, const float* __restrict__ w2_hi,
                  const float* __restrict__ w2_lo, const float* __restrict__ wsum1, int n_pt, int real_items, int total_items,
                  const __grid_constant__ CUtensorMap tmx) {
    // total_items bounds the tile loops; items >= real_items are phantom tiles (zero-filled loads, no stores): with a cluster every CTA
    // runs the same number of tiles so that the multicast weight stream is consumed in lockstep
    using Cfg = MlpCfg<C>;
    constexpr int NJ = Cfg::NJ, KC = Cfg::KC, SLOTS = Cfg::SLOTS, SG = Cfg::SG, R = Cfg::R, NBUF = Cfg::NBUF;
    constexpr int S1 = Cfg::S1, S2 = Cfg::S2, W_STAGE = Cfg::W_STAGE, SLOT_COLS = Cfg::SLOT_COLS;
    constexpr bool RESIDENT = Cfg::RESIDENT;
    constexpr int CL = Cfg::CL;
    constexpr int XA0 = NBUF * C, G0 = NBUF * C + 2 * C;         // column offsets inside a slot
    constexpr uint32_t W_HALF_BYTES = (uint32_t)C * 32u * 4u;    // hi (or lo) part of a weight stage
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* xs = reinterpret_cast<float*>(smem_raw);    // [SLOTS][KC][16 channels][128 pixels]
    float* w1r = xs + SLOTS * C * TC_M;                // [S1][hi | lo]
    float* w2r = w1r + S1 * W_STAGE;                   // [S2][hi | lo]
    __shared__ __align__(8) MlpBars bars[SLOTS];
    __shared__ __align__(8) uint64_t bar_w1_full[S1], bar_w1_free[S1], bar_w2_full[S2], bar_w2_free[S2];
    __shared__ __align__(8) uint64_t bar_w1_pfree[S1], bar_w2_pfree[S2];   // the peer CTA's MMAs have read the stage (this CTA multicasts into it)
    __shared__ uint32_t tmem_base_s;
    __shared__ float4 s_c1[2 * C];   // per PAIR of hidden columns {wsum1[n], wsum1[n+1], b1[n], b1[n+1]}
    __shared__ float2 s_c2[C];       // per output {b2, gamma}

    // the shuffle makes the warp index provably warp-uniform for the compiler (uniform registers in the MMA / loader loops)
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int P = p.P;
    const int tile_stride = SLOTS * (int)gridDim.x;   // distance between consecutive tiles of one slot

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < SLOTS; ++s) {
            MlpBars& b = bars[s];
            mlp_mbar_init(smem_u32(&b.x_full), 1);
            mlp_mbar_init(smem_u32(&b.x_empty), 4 * SG);
            mlp_mbar_init(smem_u32(&b.xa_full), 4 * SG);
            mlp_mbar_init(smem_u32(&b.xa_free), 1);
            for (int i = 0; i < 2; ++i) {
                mlp_mbar_init(smem_u32(&b.acc2_full[i]), 1);
                mlp_mbar_init(smem_u32(&b.acc2_empty[i]), 4 * SG);
                mlp_mbar_init(smem_u32(&b.acc1_full[i]), 1);
                mlp_mbar_init(smem_u32(&b.acc1_drained[i]), 4);
                mlp_mbar_init(smem_u32(&b.a2_full[i]), 4);
                mlp_mbar_init(smem_u32(&b.a2_free[i]), 1);
            }
        }
        for (int i = 0; i < S1; ++i) {
            mlp_mbar_init(smem_u32(&bar_w1_full[i]), 1);
            mlp_mbar_init(smem_u32(&bar_w1_free[i]), 1);
            mlp_mbar_init(smem_u32(&bar_w1_pfree[i]), CL > 1 ? CL - 1 : 1);
        }
        for (int i = 0; i < S2; ++i) {
            mlp_mbar_init(smem_u32(&bar_w2_full[i]), 1);
            mlp_mbar_init(smem_u32(&bar_w2_free[i]), 1);
            mlp_mbar_init(smem_u32(&bar_w2_pfree[i]), CL > 1 ? CL - 1 : 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 2 * C; i += Cfg::THREADS) s_c1[i] = make_float4(wsum1[2 * i], wsum1[2 * i + 1], p.b1[2 * i], p.b1[2 * i + 1]);
    for (int i = tid; i < C; i += Cfg::THREADS) s_c2[i] = make_float2(p.b2[i], p.gamma[i]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    const uint32_t crank = CL > 1 ? mlp_cluster_rank() : 0u;
    if (CL > 1) mlp_cluster_sync();   // every CTA's barriers are initialised before the leader's first multicast lands

    // register re-allocation between the warpgroups: the MMA / loader warps need few registers, the epilogue warps hold the chunk
    // (32 accumulator values + their hi/lo terms) AND the prefetched residual rows: 4 x 40 + 8 x 232 = 12 x 168
    if (warp >= 8) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");   // one instruction for the whole warpgroup (warps 8-11)
      if (warp < 10) {
        // ================================================================== MMA warp of slot (warp - 8); one lane issues
        const int slot = warp - 8;
        if (slot < SLOTS) {
            MlpBars& bs = bars[slot];
            constexpr uint32_t idesc1 = tf32_idesc(32), idesc2 = tf32_idesc(C);
            const uint32_t w1r_s = smem_u32(w1r), w2r_s = smem_u32(w2r);
            // __shfl_sync(.., 0) tells the compiler the TMEM base (read from shared memory) is warp-uniform
            const uint32_t t_slot = __shfl_sync(0xffffffffu, tmem_d, 0) + (uint32_t)(slot * SLOT_COLS);
            uint32_t q1 = 0, q2 = 0;   // weight chunks consumed so far (ring positions)
            int k = 0;                 // tiles of this slot so far
            const bool elected = tc_elect_one();   // the same lane for the whole kernel: the warp is converged here
            // GEMM 1 of hidden chunk j: accumulator of group sg <- XA . W1[:, 32j .. 32j+31]
            auto gemm1 = [&](int j, int sg) {
                const uint32_t s = RESIDENT ? (uint32_t)j : q1 % S1;
                if (!RESIDENT || k == 0) mlp_wait(smem_u32(&bar_w1_full[s]), RESIDENT ? 0u : (q1 / S1) & 1u);
                const uint32_t b_hi = w1r_s + s * (uint32_t)W_STAGE * 4u;
                const uint64_t dh = kmajor_desc(b_hi, 32, 0), dl = kmajor_desc(b_hi + W_HALF_BYTES, 32, 0);
                const uint32_t acc = t_slot + (uint32_t)(G0 + sg * 96);
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) {
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint32_t ah = t_slot + (uint32_t)(XA0 + kc * 32 + ks * 8), al = ah + 16u;
                        const uint64_t off = (uint64_t)((kc * 2048 + ks * 1024) >> 4);   // the descriptor's address field counts 16-byte units
                        if (elected) {
                            mma_tf32_ts(acc, ah, dh + off, idesc1, (kc > 0 || ks > 0) ? 1u : 0u);
                            mma_tf32_ts(acc, al, dh + off, idesc1, 1u);
                            mma_tf32_ts(acc, ah, dl + off, idesc1, 1u);
                        }
                    }
                }
                if (elected) {
                    tc_commit(smem_u32(&bs.acc1_full[sg]));
                    if (!RESIDENT) {
                        tc_commit(smem_u32(&bar_w1_free[s]));
                        if (CL > 1) mlp_commit_mc(smem_u32(&bar_w1_pfree[s]), 1u, (uint16_t)(1u << (crank ^ 1u)));   // tell the peer
                    }
                    if (j == NJ - 1) tc_commit(smem_u32(&bs.xa_free));   // every GEMM 1 of the tile has read XA
                }
                __syncwarp();
                ++q1;
            };
            // GEMM 2 of hidden chunk j: output accumulator `buf` (+)= A2 of group sg . W2[32j .. 32j+31, :]
            auto gemm2 = [&](int j, int sg, int buf) {
                const uint32_t s = RESIDENT ? (uint32_t)j : q2 % S2;
                if (!RESIDENT || k == 0) mlp_wait(smem_u32(&bar_w2_full[s]), RESIDENT ? 0u : (q2 / S2) & 1u);
                const uint32_t b_hi = w2r_s + s * (uint32_t)W_STAGE * 4u;
                const uint64_t dh = kmajor_desc(b_hi, C, 0), dl = kmajor_desc(b_hi + W_HALF_BYTES, C, 0);
                const uint32_t a2 = t_slot + (uint32_t)(G0 + sg * 96 + 32);
                const uint32_t acc = t_slot + (uint32_t)(buf * C);
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint32_t ah = a2 + (uint32_t)(kk * 16 + ks * 8), al = ah + 32u;
                        const uint64_t off = (uint64_t)((kk * C * 64 + ks * 2 * (C / 8) * 128) >> 4);
                        if (elected) {
                            mma_tf32_ts(acc, ah, dh + off, idesc2, (j > 0 || kk > 0 || ks > 0) ? 1u : 0u);
                            mma_tf32_ts(acc, al, dh + off, idesc2, 1u);
                            mma_tf32_ts(acc, ah, dl + off, idesc2, 1u);
                        }
                    }
                }
                if (elected) {
                    tc_commit(smem_u32(&bs.a2_free[sg]));
                    if (!RESIDENT) {
                        tc_commit(smem_u32(&bar_w2_free[s]));
                        if (CL > 1) mlp_commit_mc(smem_u32(&bar_w2_pfree[s]), 1u, (uint16_t)(1u << (crank ^ 1u)));
                    }
                    if (j == NJ - 1) tc_commit(smem_u32(&bs.acc2_full[buf]));
                }
                __syncwarp();
                ++q2;
            };
#pragma unroll 1
            for (int item = blockIdx.x + slot * (int)gridDim.x; item < total_items; item += tile_stride, ++k) {
                const bool next_tile = item + tile_stride < total_items;
                const int buf = k % NBUF;
                if (k == 0) {   // first round of the first tile; later first rounds are issued at the drains of the previous tile's last round
                    mlp_wait_tc(smem_u32(&bs.xa_full), 0u);
#pragma unroll 1
                    for (int sg = 0; sg < SG; ++sg) gemm1(sg, sg);
                }
#pragma unroll 1
                for (int r = 0; r < R; ++r) {
                    const uint32_t use = (uint32_t)(k * R + r);   // chunks each group of the slot has processed before this round
                    // accumulators drained -> GEMM 1 of the next round (underneath the GELU arithmetic of this one)
                    if (r + 1 < R || next_tile) {
                        if (r + 1 == R) mlp_wait_tc(smem_u32(&bs.xa_full), (uint32_t)(k + 1) & 1u);   // the next tile's XA is staged
#pragma unroll 1
                        for (int sg = 0; sg < SG; ++sg) {
                            mlp_wait_tc(smem_u32(&bs.acc1_drained[sg]), use & 1u);
                            gemm1(((r + 1) % R) * SG + sg, sg);
                        }
                    }
                    // A2 slots filled -> GEMM 2, in chunk order (the accumulation order is part of the result)
#pragma unroll 1
                    for (int sg = 0; sg < SG; ++sg) {
                        mlp_wait_tc(smem_u32(&bs.a2_full[sg]), use & 1u);
                        if (r == 0 && sg == 0 && k >= NBUF)   // the output epilogue that last used this accumulator has drained it
                            mlp_wait_tc(smem_u32(&bs.acc2_empty[buf]), (uint32_t)(k / NBUF - 1) & 1u);
                        gemm2(r * SG + sg, sg, buf);
                    }
                }
            }
        }
      } else {
        // ================================================================== loader warps: activation tiles per slot (+ weights: warp 10)
        const int slot = warp - 10;
        if (lane == 0 && slot < SLOTS) {
            MlpBars& bs = bars[slot];
            const uint32_t xs_s = smem_u32(xs + slot * C * TC_M), w1r_s = smem_u32(w1r), w2r_s = smem_u32(w2r);
            uint32_t q1 = 0, q2 = 0;
            int k = 0;
            auto load_x = [&](int item, int t) {
                const int pt = item % n_pt, b = item / n_pt;
                if (t > 0) mlp_wait(smem_u32(&bs.x_empty), (uint32_t)(t - 1) & 1u);   // every group of the slot has read the previous tile
                const uint32_t full = smem_u32(&bs.x_full);
                tma_mbar_expect_tx(full, (uint32_t)C * TC_M * 4u);   // boxes are always complete: out-of-range pixels are zero-filled
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) tma_load_4d(xs_s + (uint32_t)kc * TC_KC * TC_M * 4u, &tmx, pt * TC_M, 0, kc * TC_KC, b, full);
            };
            auto load_w = [&](int j, bool first, uint32_t& q) {
                const int ns = first ? S1 : S2;
                uint64_t* fullb = first ? bar_w1_full : bar_w2_full;
                uint64_t* freeb = first ? bar_w1_free : bar_w2_free;
                uint64_t* pfreeb = first ? bar_w1_pfree : bar_w2_pfree;
                const uint32_t s = RESIDENT ? (uint32_t)j : q % (uint32_t)ns;
                if (!RESIDENT && q >= (uint32_t)ns) {   // the MMAs that read the stage are done (re-arming the barrier needs that in every CTA)
                    mlp_wait(smem_u32(&freeb[s]), (q / (uint32_t)ns - 1u) & 1u);
                    if (CL > 1) mlp_wait(smem_u32(&pfreeb[s]), (q / (uint32_t)ns - 1u) & 1u);   // ... and in the peer, whose stage this CTA writes too
                }
                const uint32_t full = smem_u32(&fullb[s]);
                const uint32_t dst = (first ? w1r_s : w2r_s) + s * (uint32_t)W_STAGE * 4u;
                const float* hi = (first ? w1_hi : w2_hi) + (long long)j * (C * 32);
                const float* lo = (first ? w1_lo : w2_lo) + (long long)j * (C * 32);
                tma_mbar_expect_tx(full, 2u * W_HALF_BYTES);
                if (CL > 1) {
                    // the two CTAs of a cluster each fetch HALF of the stage (rank 0 the hi tiles, rank 1 the lo tiles) and multicast it
                    // into both ring stages, completing both barriers: per SM half the bulk-copy requests for the same weight stream
                    // (the stream - 576 KB per tile at C = 96 - is latency bound: the rings hold 7 x 24 KB and cannot be deeper)
                    if (crank == 0) mlp_bulk_g2s_mc(dst, hi, W_HALF_BYTES, full, (uint16_t)3);
                    else mlp_bulk_g2s_mc(dst + W_HALF_BYTES, lo, W_HALF_BYTES, full, (uint16_t)3);
                } else {
                    mlp_bulk_g2s(dst, hi, W_HALF_BYTES, full);
                    mlp_bulk_g2s(dst + W_HALF_BYTES, lo, W_HALF_BYTES, full);
                }
                ++q;
            };
            const int first_item = blockIdx.x + slot * (int)gridDim.x;
            if (first_item < total_items) load_x(first_item, 0);
            if (RESIDENT) {
                if (slot == 0) {
#pragma unroll 1
                    for (int j = 0; j < NJ; ++j) {
                        load_w(j, true, q1);
                        load_w(j, false, q2);
                    }
                }
#pragma unroll 1
                for (int item = first_item; item + tile_stride < total_items; item += tile_stride, ++k) load_x(item + tile_stride, k + 1);
            } else {
                // streamed weights: this thread feeds the W1 ring (and the activation tiles), warp 11 the W2 ring - each blocks only on its
                // own ring's free barriers, so both run as far ahead as their stages allow.  (A single thread issuing both rings in MMA
                // order stalled the W1 prefetch behind the W2 stage that only frees at the end of a round: ~5 us per 2-chunk round
                // where tensor core and epilogue need ~1.2 us.)
#pragma unroll 1
                for (int item = first_item; item < total_items; item += tile_stride, ++k) {
#pragma unroll 1
                    for (int j = 0; j < NJ; ++j) {
                        load_w(j, true, q1);
                        if (j == SG - 1 && item + tile_stride < total_items) load_x(item + tile_stride, k + 1);
                    }
                }
            }
        } else if (lane == 0 && !RESIDENT && slot == 1) {
            // ---- W2 ring (streamed weights only; with resident weights warp 11 is the second slot's activation loader above)
            const uint32_t w1r_s = smem_u32(w1r), w2r_s = smem_u32(w2r);
            uint32_t q2 = 0;
            auto load_w2 = [&](int j) {
                const uint32_t s = q2 % (uint32_t)S2;
                if (q2 >= (uint32_t)S2) {
                    mlp_wait(smem_u32(&bar_w2_free[s]), (q2 / (uint32_t)S2 - 1u) & 1u);
                    if (CL > 1) mlp_wait(smem_u32(&bar_w2_pfree[s]), (q2 / (uint32_t)S2 - 1u) & 1u);
                }
                const uint32_t full = smem_u32(&bar_w2_full[s]);
                const uint32_t dst = w2r_s + s * (uint32_t)W_STAGE * 4u;
                const float* hi = w2_hi + (long long)j * (C * 32);
                const float* lo = w2_lo + (long long)j * (C * 32);
                tma_mbar_expect_tx(full, 2u * W_HALF_BYTES);
                if (CL > 1) {
                    if (crank == 0) mlp_bulk_g2s_mc(dst, hi, W_HALF_BYTES, full, (uint16_t)3);
                    else mlp_bulk_g2s_mc(dst + W_HALF_BYTES, lo, W_HALF_BYTES, full, (uint16_t)3);
                } else {
                    mlp_bulk_g2s(dst, hi, W_HALF_BYTES, full);
                    mlp_bulk_g2s(dst + W_HALF_BYTES, lo, W_HALF_BYTES, full);
                }
                ++q2;
            };
            (void)w1r_s;
#pragma unroll 1
            for (int item = blockIdx.x; item < total_items; item += tile_stride) {
#pragma unroll 1
                for (int j = 0; j < NJ; ++j) load_w2(j);
            }
        }
        __syncwarp();
      }
    } else {
        // ================================================================== epilogue groups (thread = pixel = TMEM lane)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;" ::: "memory");
        const int g = warp >> 2, wq = warp & 3, px = wq * 32 + lane;
        const int slot = Cfg::TWO_SLOTS ? g : 0, sg = Cfg::TWO_SLOTS ? 0 : g;
        MlpBars& bs = bars[slot];
        const uint32_t t_lane = tmem_d + ((uint32_t)(wq * 32) << 16) + (uint32_t)(slot * SLOT_COLS);
        const uint32_t t_acc1 = t_lane + (uint32_t)(G0 + sg * 96), t_a2 = t_acc1 + 32u;
        const float* xs_slot = xs + slot * C * TC_M;

        struct Item {
            int b, pp;
            bool p_ok;
            float rs, ms;
        };
        // reads the landed activation tile: LayerNorm statistics of this pixel + this group's share of the XA columns
        auto produce_xa = [&](int item, int t) {
            Item it;
            const int pt = item % n_pt;
            it.b = item / n_pt;
            it.pp = pt * TC_M + px;
            it.p_ok = it.pp < P && item < real_items;
            mlp_wait(smem_u32(&bs.x_full), (uint32_t)t & 1u);
            if (t > 0) mlp_wait_tc(smem_u32(&bs.xa_free), (uint32_t)(t - 1) & 1u);   // GEMM 1 of the previous tile no longer reads XA
            const float* gs = xs_slot + px;
            // LayerNorm running sums, shifted by the pixel's first channel to avoid cancellation; two partial sums per statistic
            // (k % 16 < 8 and >= 8) added at the end: the summation order of the two-thread-per-pixel GEMM kernel
            const float shift = it.p_ok ? gs[0] : 0.f;
            float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
                float v[TC_KC];
#pragma unroll
                for (int e = 0; e < TC_KC; ++e) v[e] = it.p_ok ? gs[(kc * TC_KC + e) * TC_M] : 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float d = it.p_ok ? v[e] - shift : 0.f;
                    s1a += d;
                    s2a = fmaf(d, d, s2a);
                }
#pragma unroll
                for (int e = 8; e < 16; ++e) {
                    const float d = it.p_ok ? v[e] - shift : 0.f;
                    s1b += d;
                    s2b = fmaf(d, d, s2b);
                }
                if (kc % SG == sg) {   // warp-uniform
                    uint32_t hi[TC_KC], lo[TC_KC];
#pragma unroll
                    for (int e = 0; e < TC_KC; ++e) {
                        hi[e] = __float_as_uint(v[e]) & 0xffffe000u;
                        lo[e] = __float_as_uint(v[e] - __uint_as_float(hi[e]));
                    }
                    tmem_st16(t_lane + (uint32_t)(XA0 + kc * 32), hi);
                    tmem_st16(t_lane + (uint32_t)(XA0 + kc * 32 + 16), lo);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                mlp_mbar_arrive(smem_u32(&bs.xa_full));
                mlp_mbar_arrive(smem_u32(&bs.x_empty));
            }
            const float t1 = (s1a + s1b) / (float)C;
            const float t2 = (s2a + s2b) / (float)C;
            it.rs = 1.0f / sqrtf(fmaxf(t2 - t1 * t1, 0.f) + p.ln_eps);
            it.ms = (shift + t1) * it.rs;
            return it;
        };
        // residual rows of this group's output units, requested at the top of a tile's last chunk and consumed by its output epilogue
        // (v2 issued them right before the accumulator read: 6 % of all stall samples sat on their first use)
        constexpr int NU = C / 16 / SG;   // 16-column output units per group
        float rr[NU * 16];
        auto load_res = [&](const Item& it) {
            if (it.p_ok) {
                const float* rptr = p.res + (long long)it.b * p.res_bs + it.pp;
#pragma unroll
                for (int n = 0; n < NU; ++n) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) rr[n * 16 + i] = __ldg(rptr + (long long)(16 * (sg + n * SG) + i) * P);
                }
            }
        };
        // output epilogue of tile t_k: this group's 16-column units of accumulator k % NBUF -> + b2, * gamma, + residual -> stores
        auto final_epilogue = [&](const Item& done, int t_k) {
            const int buf = t_k % NBUF;
            float* optr = p.out + (long long)done.b * p.out_bs + done.pp;
            mlp_wait_tc(smem_u32(&bs.acc2_full[buf]), (uint32_t)(t_k / NBUF) & 1u);
#pragma unroll
            for (int n = 0; n < NU; ++n) {
                const int uu = sg + n * SG;
                uint32_t r[16];
                tmem_ld16(t_lane + (uint32_t)(buf * C + 16 * uu), r);
                if (done.p_ok) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float2 c = s_c2[16 * uu + i];
                        // the GEMM kernel's epilogue with rs = scale = 1, ms = 0 (acc + b2), then gamma * y + res
                        optr[(long long)(16 * uu + i) * P] = fmaf(c.y, __uint_as_float(r[i]) + c.x, rr[n * 16 + i]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mlp_mbar_arrive(smem_u32(&bs.acc2_empty[buf]));
        };

        uint32_t u = 0;   // hidden chunks this group has processed (phase bookkeeping of its four barriers)
        int k = 0;
        int item = blockIdx.x + slot * (int)gridDim.x;
        Item cur = {}, nxt = {}, fin = {};
        bool pending = false;
        if (item < total_items) cur = produce_xa(item, 0);
#pragma unroll 1
        for (; item < total_items; ++k) {
            const int next = item + tile_stride;
#pragma unroll 1
            for (int r = 0; r < R; ++r, ++u) {
                const int j = r * SG + sg;
                // the next tile's A operand before the last chunk: the chunk stream then crosses the tile boundary without a gap
                if (r == R - 1) {
                    load_res(cur);
                    if (next < total_items) nxt = produce_xa(next, k + 1);
                }
                // ---- hidden chunk j: accumulator -> LayerNorm fix-up + bias + GELU -> hi/lo A operand of GEMM 2
                mlp_wait_tc(smem_u32(&bs.acc1_full[sg]), u & 1u);
                uint32_t ra[16], rb[16];
                tmem_ld16(t_acc1, ra);
                tmem_ld16(t_acc1 + 16u, rb);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mlp_mbar_arrive(smem_u32(&bs.acc1_drained[sg]));   // GEMM 1 of this group's next chunk may start now
                const float2 rs2 = make_float2(cur.rs, cur.rs), nms2 = make_float2(-cur.ms, -cur.ms);
                // all 16 column pairs as independent chains in ONE block (v3 split them in two halves around the A2-slot wait: ncu showed
                // 31 % of the epilogue warps' samples on fixed-latency dependencies - with two warps per scheduler the chunk needs the ILP),
                // and only then the wait for GEMM 2 of the previous chunk, which by now has long completed
                uint32_t hi0[16], lo0[16], hi1[16], lo1[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float4 c = s_c1[j * 16 + i];
                    const float2 a = i < 8 ? make_float2(__uint_as_float(ra[2 * (i & 7)]), __uint_as_float(ra[2 * (i & 7) + 1]))
                                           : make_float2(__uint_as_float(rb[2 * (i & 7)]), __uint_as_float(rb[2 * (i & 7) + 1]));
                    // same expression as the GEMM kernel's epilogue with scale = 1: act(rs * acc - ms * wsum + bias), two columns at a time
                    float2 y = __ffma2_rn(rs2, a, __ffma2_rn(nms2, make_float2(c.x, c.y), make_float2(c.z, c.w)));
                    y = mlp_gelu2(y);
                    const uint32_t h0 = __float_as_uint(y.x) & 0xffffe000u, h1 = __float_as_uint(y.y) & 0xffffe000u;
                    const float2 l = __ffma2_rn(make_float2(__uint_as_float(h0), __uint_as_float(h1)), make_float2(-1.0f, -1.0f), y);   // y - hi, exact
                    if (i < 8) {
                        hi0[2 * (i & 7)] = h0, hi0[2 * (i & 7) + 1] = h1;
                        lo0[2 * (i & 7)] = __float_as_uint(l.x), lo0[2 * (i & 7) + 1] = __float_as_uint(l.y);
                    } else {
                        hi1[2 * (i & 7)] = h0, hi1[2 * (i & 7) + 1] = h1;
                        lo1[2 * (i & 7)] = __float_as_uint(l.x), lo1[2 * (i & 7) + 1] = __float_as_uint(l.y);
                    }
                }
                if (u > 0) mlp_wait_tc(smem_u32(&bs.a2_free[sg]), (u - 1u) & 1u);   // GEMM 2 of the previous chunk has read the slot
                tmem_st16(t_a2, hi0);
                tmem_st16(t_a2 + 16u, hi1);
                tmem_st16(t_a2 + 32u, lo0);
                tmem_st16(t_a2 + 48u, lo1);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mlp_mbar_arrive(smem_u32(&bs.a2_full[sg]));
                // with two output accumulators the previous tile's outputs go out after this tile's first chunk
                if (r == 0 && pending) {
                    final_epilogue(fin, k - 1);
                    pending = false;
                }
            }
            if (NBUF == 2 && next < total_items) {
                fin = cur;
                pending = true;
            } else {
                final_epilogue(cur, k);
            }
            cur = nxt;
            item = next;
        }
    }

    // ---- teardown
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512) : "memory");
    if (CL > 1) mlp_cluster_sync();   // no CTA leaves while a peer may still signal its barriers
}

template <int C>
static int launch_mlp(const AchMlp& p, const float* w1_hi, const float* w1_lo, const float* w2_hi, const float* w2_lo, const float* wsum1,
                      cudaStream_t st) {
    using Cfg = MlpCfg<C>;
    static PerDeviceOnce once;
    static int sms_dev[ACH_MAX_DEVICES] = {};
    if (once.first()) {
        cudaFuncSetAttribute(mlp_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        sms_dev[current_device()] = sms;
    }
    int sms = sms_dev[current_device()];
    if (sms <= 0) sms = 148;
    alignas(64) CUtensorMap tmx;
    memset(&tmx, 0, sizeof(tmx));
    // (P, 1, C, B) planes view: box = 128 pixels x 16 channels of one frame
    ACH_REQUIRE(tma_map_planes(&tmx, p.x, p.P, 1, C, p.B, p.x_bs, TC_M, 1, TC_KC), "ach_mlp_tc: the activation view cannot be described by a tensor map");
    const int n_pt = cdiv(p.P, TC_M);
    const long long total = (long long)n_pt * p.B;
    ACH_REQUIRE(total < (1LL << 30), "ach_mlp_tc: too many tiles");
    // persistent: one CTA per SM (the kernel owns all 512 TMEM columns); with two slots a CTA wants at least two tiles
    const long long want = Cfg::SLOTS == 2 ? (total + 1) / 2 : total;
    int grid = (int)(want < sms ? want : sms);
    long long bound = total;
    if (Cfg::CL > 1) {   // whole clusters, and the same number of tiles (real or phantom) in every CTA
        grid = (grid + Cfg::CL - 1) / Cfg::CL * Cfg::CL;
        if (grid > sms) grid -= Cfg::CL;
        bound = (total + grid - 1) / grid * grid;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = Cfg::CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = Cfg::CL > 1 ? 1 : 0;
    const int real_i = (int)total, bound_i = (int)bound;
    cudaLaunchKernelEx(&cfg, mlp_tc_kernel<C>, p, w1_hi, w1_lo, w2_hi, w2_lo, wsum1, n_pt, real_i, bound_i, tmx);
    return check_launch("ach_mlp_tc");
}

}  // namespace ach

extern "C" int ach_mlp_tc_supported(int C) { return C == 32 || C == 48 || C == 64 || C == 96; }

extern "C" int ach_mlp_tc(const AchMlp* pp, const float* w1_hi, const float* w1_lo, const float* w2_hi, const float* w2_lo,
                          const float* wsum1, void* stream) {
    using namespace ach;
    const AchMlp& p = *pp;
    ACH_REQUIRE(p.x && p.res && p.out && p.b1 && p.b2 && p.gamma && w1_hi && w1_lo && w2_hi && w2_lo && wsum1, "ach_mlp_tc: null arg");
    ACH_REQUIRE(p.B > 0 && p.P > 0 && p.P % 4 == 0, "ach_mlp_tc: bad dims (P=%d must be a positive multiple of 4)", p.P);
    ACH_REQUIRE(aligned16(p.x) && p.x_bs % 4 == 0, "ach_mlp_tc: x must be a 16-byte aligned view");
    ACH_REQUIRE(aligned16(w1_hi) && aligned16(w1_lo) && aligned16(w2_hi) && aligned16(w2_lo), "ach_mlp_tc: weight tiles must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (p.C) {
        case 32: return launch_mlp<32>(p, w1_hi, w1_lo, w2_hi, w2_lo, wsum1, st);
        case 48: return launch_mlp<48>(p, w1_hi, w1_lo, w2_hi, w2_lo, wsum1, st);
        case 64: return launch_mlp<64>(p, w1_hi, w1_lo, w2_hi, w2_lo, wsum1, st);
        case 96: return launch_mlp<96>(p, w1_hi, w1_lo, w2_hi, w2_lo, wsum1, st);
        default: break;
    }
    set_error("ach_mlp_tc: C=%d not instantiated (32, 48, 64, 96); use two ach_pw_conv_tc launches", p.C);
    return ACH_ERR_INVALID;
}
