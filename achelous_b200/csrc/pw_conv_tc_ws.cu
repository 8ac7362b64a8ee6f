// Warp-specialised, asynchronous version of the tcgen05 pointwise-convolution GEMM (same contract, operand layouts and
// 3xTF32 arithmetic as pw_conv_tc.cu; this is the default path, the synchronous kernel there is the A/B fallback).
//
// Why: in the synchronous kernel every 16-wide K chunk is a round trip  global load -> split -> st.shared -> fence ->
// __syncthreads -> MMA -> commit  with the weight tile loaded by the same threads right before it is needed.  ncu
// (profiles/r1_ncu_final_summary.txt, source page): on the long-K layers (K = 352, 22 chunks, only 200 work items on 148
// SMs) 27 % of all stall samples sit on the first use of the loaded registers and on the barrier - 3.3 us per chunk for
// 24 KB of operands.  Here the three jobs run decoupled, synchronised only through mbarriers:
//   * warp 9 ("activation loader"): streams the fp32 activation chunks (16 channels x 128 pixels) into a WS_SG-deep
//     shared-memory ring, running ahead of the consumers across chunk AND tile boundaries.  One 3-D tensor-map TMA
//     (cp.async.bulk.tensor, box 128 px x 16 ch x 1 frame, zero-filled past P / past the channel count) per chunk;
//     the first version issued 16 row copies of 512 bytes per chunk and the K = 128 layers then sat at ~2.3 TB/s with
//     23 % of the stall samples waiting for a chunk to land - a per-copy cost, not a byte cost.  Chunks that straddle
//     the two sources of a concat-free layer still use the row copies.  The first versions loaded activations into registers two
//     chunks ahead: ncu showed long-scoreboard stalls of 3.8 per issue on the K = 128 layers and 2 TB/s of DRAM - the
//     register prefetch was too shallow for HBM latency; the ring holds 5-8 chunks (40-64 KB) in flight per CTA.
//   * warps 0-7 (256 threads, "producers"): read their pixel's 8 k of a landed chunk from the ring (conflict-free
//     LDS), split them into
//     tf32 hi/lo terms and write them into a ring of A-operand stages in TENSOR MEMORY (tcgen05.st: thread = pixel = TMEM
//     lane, 8 hi + 8 lo columns per thread and chunk; the MMA takes A from TMEM).  The first version kept the ring in
//     shared memory: two STS.128 per 4 k plus a proxy fence, and the tensor core re-read every A tile three times
//     through the shared-memory pipe (ah.bh, al.bh, ah.bl) - on the RCBlock kernel the same change cut 25-35 %.
//     Per-warp mbarrier arrive, no CTA barrier.
//     After the last chunk they prefetch the first two chunks of the CTA's NEXT tile, then run the epilogue of the
//     current one (TMEM -> registers -> LayerNorm / scale / bias / activation / residual -> coalesced stores).
//   * warp 8, one elected thread ("MMA thread"): streams the pre-packed hi/lo weight tiles with 1-D bulk TMA copies
//     (cp.async.bulk + mbarrier complete_tx) into a 4-stage ring, two chunks ahead and across tile boundaries; waits
//     for "activations stored" + "weights landed", issues the 6 tcgen05.mma of the chunk and commits them to the
//     stage's "MMA done" barrier, which is what lets producers and TMA reuse a stage.
// The accumulator (NT TMEM columns) is handed back and forth with two more barriers (acc_full / acc_empty).
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace ach {

// Tile-width dependent shape of a CTA.  Narrow tiles (NT <= 64) run with ONE producer/epilogue thread per pixel
// (128 threads + MMA warp + loader warp = 192) and two A stages, so that TMEM (<= 128 columns), registers and shared
// memory admit 3-4 resident CTAs per SM; NT = 128 keeps two threads per pixel (half the columns / k each) and 2 CTAs.
// Measured: halving the resident CTAs (a deeper ring that no longer fitted twice) cost 58 % on the K=128 -> 32 layers -
// CTA-level overlap of produce / MMA wait / epilogue phases is what hides the serial chain inside one CTA.
__host__ __device__ constexpr int ws_halves(int nt) { return nt == 128 ? 2 : 1; }
__host__ __device__ constexpr int ws_sa(int nt) { return nt == 128 ? 4 : 2; }   // A-operand stages in tensor memory (32 columns each)
constexpr int WS_SB = 4;        // weight ring stages
constexpr int WS_PF = 2;        // weight chunks in flight ahead of the MMA
__host__ __device__ constexpr int ws_prod(int nt) { return 128 * ws_halves(nt); }   // producer / epilogue threads; then the MMA warp, then the loader warp
__host__ __device__ constexpr int ws_threads(int nt) { return ws_prod(nt) + 64; }
__host__ __device__ constexpr int ws_sg(int nt) { return nt == 32 ? 6 : 5; }   // activation ring depth (8 KB per stage), sized so that 3 narrow CTAs fit an SM

__host__ __device__ constexpr int ws_tmem_cols(int nt) {   // accumulator + A stages, rounded up to a power of two >= 32
    const int need = nt + ws_sa(nt) * 32;
    return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
}

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(mbar)
                 : "memory");
}
template <int N_THREADS>
__device__ __forceinline__ void prod_bar() { asm volatile("bar.sync 1, %0;" ::"n"(N_THREADS) : "memory"); }

template <int NT, int ACT, bool RES>
__global__ void __launch_bounds__(ws_threads(NT), NT == 128 ? 2 : 3)
    pw_conv_tc_ws_kernel(const AchPwConv p, const float* __restrict__ w_hi, const float* __restrict__ w_lo,
                         const float* __restrict__ wsum, int n_kchunks, int n_pt, int n_ot, int total_items) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int B_ELEMS = NT * TC_KC;
    constexpr int SG = ws_sg(NT);
    constexpr int WS_SA = ws_sa(NT), HALVES = ws_halves(NT), WS_PROD = ws_prod(NT);
    constexpr int KPT = TC_KC / HALVES;                                   // k per producer thread and chunk (8 or 16)
    constexpr int G_ELEMS = TC_KC * TC_M;                                // fp32 activation chunk: [16 channels][128 pixels]
    float* b_ring = reinterpret_cast<float*>(smem_raw);                  // [SB][b_hi | b_lo]
    float* g_ring = b_ring + WS_SB * 2 * B_ELEMS;                        // [SG][16][128]
    __shared__ __align__(8) uint64_t bar_full_a[WS_SA], bar_full_b[WS_SB], bar_mma[WS_SA], bar_acc_full, bar_acc_empty;
    __shared__ __align__(8) uint64_t bar_g_full[SG], bar_g_empty[SG];
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_ln[HALVES][TC_M][2];
    __shared__ __align__(16) float4 s_ep[NT];   // per output of the current tile: {scale, scale*wsum, bias + scale*pbias, gamma}

    // the shuffle makes the warp index provably warp-uniform for the compiler (uniform registers / branches in the role loops)
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int K = p.c0 + p.c1, P = p.P;
    // TMEM columns: [0, NT) accumulator, then WS_SA A-operand stages of 32 columns (hi k 0..15 | lo k 0..15)
    constexpr int A_COL0 = NT;
    constexpr int TMEM_COLS = ws_tmem_cols(NT);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < WS_SA; ++i) {
            mbar_init(smem_u32(&bar_full_a[i]), WS_PROD / 32);
            mbar_init(smem_u32(&bar_mma[i]), 1);
        }
        for (int i = 0; i < WS_SB; ++i) mbar_init(smem_u32(&bar_full_b[i]), 1);
        for (int i = 0; i < SG; ++i) {
            mbar_init(smem_u32(&bar_g_full[i]), 1);
            mbar_init(smem_u32(&bar_g_empty[i]), WS_PROD / 32);
        }
        mbar_init(smem_u32(&bar_acc_full), 1);
        mbar_init(smem_u32(&bar_acc_empty), WS_PROD / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    const uint32_t b_ring_s = smem_u32(b_ring);
    // programmatic dependent launch: everything above (TMEM allocation, barrier init) may overlap the tail of the previous
    // kernel in the stream; nothing below touches global memory before that kernel has completed and flushed
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == WS_PROD / 32) {
        // ================================================================== MMA + weight-TMA warp
        // Every lane runs the loop and the barrier waits (warp-uniform control flow: descriptors, TMEM and barrier addresses live in
        // uniform registers), ONE lane chosen by elect.sync issues.  Under the former `if (lane == 0)` ptxas wrapped every UTCHMMA /
        // UBLKCP in an ELECT + 4 x R2UR + vote + branch "waterfall" (~10 instructions per MMA on this single warp): the same finding
        // as in mlp_tc.cu, where it bounded the kernel.
        {
            const bool elected = tc_elect_one();
            const uint32_t tmem_u = tc_uniform(tmem_d);
            constexpr uint32_t idesc = tf32_idesc(NT);
            constexpr uint32_t B_LBO = (NT / 8) * 128;
            int pit = 0, p_item = blockIdx.x, p_c = 0;   // weight prefetch cursor: chunk counter, (item, chunk)
            auto issue_b = [&]() {
                if (p_item >= total_items) return;
                const int sb = pit % WS_SB;
                const int prev = pit - WS_SB;            // the chunk whose MMAs last read this stage
                if (prev >= 0) mbar_wait(smem_u32(&bar_mma[prev % WS_SA]), (uint32_t)(prev / WS_SA) & 1u);
                const long long blk = ((long long)(p_item % n_ot) * n_kchunks + p_c) * B_ELEMS;
                const uint32_t full = smem_u32(&bar_full_b[sb]);
                const uint32_t dst = b_ring_s + (uint32_t)sb * 2u * B_ELEMS * 4u;
                if (elected) {
                    mbar_expect_tx(full, 2u * B_ELEMS * 4u);
                    bulk_g2s(dst, w_hi + blk, B_ELEMS * 4u, full);
                    bulk_g2s(dst + B_ELEMS * 4u, w_lo + blk, B_ELEMS * 4u, full);
                }
                ++pit;
                if (++p_c == n_kchunks) {
                    p_c = 0;
                    p_item += gridDim.x;
                }
            };
#pragma unroll 1
            for (int i = 0; i < WS_PF; ++i) issue_b();
            int it = 0, tile_n = 0;
#pragma unroll 1
            for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++tile_n) {
                if (tile_n > 0) {   // the epilogue of the previous tile has drained the accumulator
                    mbar_wait(smem_u32(&bar_acc_empty), (uint32_t)(tile_n - 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
#pragma unroll 1
                for (int c = 0; c < n_kchunks; ++c, ++it) {
                    issue_b();   // chunk it + WS_PF
                    const int sa = it % WS_SA, sb = it % WS_SB;
                    mbar_wait(smem_u32(&bar_full_b[sb]), (uint32_t)(it / WS_SB) & 1u);
                    mbar_wait(smem_u32(&bar_full_a[sa]), (uint32_t)(it / WS_SA) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi_t = tmem_u + (uint32_t)(A_COL0 + sa * 32), a_lo_t = a_hi_t + 16u;
                    const uint32_t b_hi_s = b_ring_s + (uint32_t)sb * 2u * B_ELEMS * 4u;
                    const uint64_t dh = make_desc(b_hi_s, B_LBO, 128u, 0), dl = make_desc(b_hi_s + B_ELEMS * 4u, B_LBO, 128u, 0);
                    if (elected) {
#pragma unroll
                        for (int ks = 0; ks < TC_KC / 8; ++ks) {
                            // NT = 128 (two producer threads per pixel): K-step ks of the stage is [hi 8 | lo 8] at column 16 * ks, written by
                            // half ks with ONE 16-column store; one thread per pixel: [hi 16 | lo 16]
                            const uint32_t ah = HALVES == 2 ? a_hi_t + (uint32_t)ks * 16u : a_hi_t + (uint32_t)ks * 8u;
                            const uint32_t al = HALVES == 2 ? ah + 8u : a_lo_t + (uint32_t)ks * 8u;
                            const uint64_t off = (uint64_t)((ks * 2 * B_LBO) >> 4);   // the descriptor's address field counts 16-byte units
                            mma_tf32_ts(tmem_u, ah, dh + off, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                            mma_tf32_ts(tmem_u, al, dh + off, idesc, 1u);
                            mma_tf32_ts(tmem_u, ah, dl + off, idesc, 1u);
                        }
                        tc_commit(smem_u32(&bar_mma[sa]));
                        if (c == n_kchunks - 1) tc_commit(smem_u32(&bar_acc_full));
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    } else if (warp == WS_PROD / 32 + 1) {
        // ================================================================== activation loader (lanes 0-15: one channel row each)
        const uint32_t g_ring_s = smem_u32(g_ring);
        uint32_t git = 0;
#pragma unroll 1
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const int pt = (item / n_ot) % n_pt, b = item / (n_ot * n_pt);
            const int pp0 = pt * TC_M;
            const uint32_t bytes = (uint32_t)min(TC_M, P - pp0) * 4u;     // multiple of 16: P % 4 == 0
            const float* x0 = p.x0 + (long long)b * p.x0_bs + pp0;
            const float* x1 = p.x1 ? p.x1 + (long long)b * p.x1_bs + pp0 : nullptr;
#pragma unroll 1
            for (int c = 0; c < n_kchunks; ++c, ++git) {
                const uint32_t g = git % SG;
                if (git >= (uint32_t)SG) mbar_wait(smem_u32(&bar_g_empty[g]), (git / SG - 1u) & 1u);   // every producer warp has read it
                const int k0 = c * TC_KC;
                const int nk = min(TC_KC, K - k0);
                const uint32_t full = smem_u32(&bar_g_full[g]);
                // One bulk copy per channel row (16 x 512 B per chunk).  A tensor-map box per chunk (kernel-parameter CUtensorMap, one
                // cp.async.bulk.tensor instead of 16 row copies) measured 3 % faster on these layers but was NOT reproducible run to run once
                // other kernels shared the GPU: back-to-back launches of this kernel on one stream occasionally accumulated a tile from the
                // wrong activations (tools/op_race_probe.py: 19 of 250 runs with the tensor-map path, 0 of 250 with the row copies - every
                // barrier / lane / single-CTA-per-SM variant of the kernel kept the failure, only the copy path mattered).  DESIGN.md §7.
                if (lane == 0) mbar_expect_tx(full, (uint32_t)nk * bytes);
                __syncwarp();
                if (lane < nk) {
                    const int k = k0 + lane;
                    const float* src = (k < p.c0) ? x0 + (long long)k * P : x1 + (long long)(k - p.c0) * P;
                    bulk_g2s(g_ring_s + (g * G_ELEMS + (uint32_t)lane * TC_M) * 4u, src, bytes, full);
                }
            }
        }
    } else {
        // ================================================================== producers + epilogue (256 threads)
        const int px = tid & (TC_M - 1), half = tid >> 7;   // half is warp-uniform
        const uint32_t t_lane = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);

        struct Item {
            const float* x0;
            const float* x1;
            int b, o_base, pp;
            bool p_ok;
        };
        auto decode = [&](int item) {
            Item t;
            const int o_tile = item % n_ot;
            const int pt = (item / n_ot) % n_pt;
            t.b = item / (n_ot * n_pt);
            t.o_base = o_tile * NT;
            t.pp = pt * TC_M + px;
            t.p_ok = t.pp < P;
            t.x0 = p.x0 + (long long)t.b * p.x0_bs;
            t.x1 = p.x1 ? p.x1 + (long long)t.b * p.x1_bs : nullptr;
            return t;
        };
        uint32_t it = 0;
        int tile_n = 0;
        int item = blockIdx.x;
        Item cur = decode(item);
#pragma unroll 1
        for (; item < total_items; ++tile_n) {
            // LayerNorm running sums, shifted by the pixel's first channel to avoid cancellation
            float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
            for (int c = 0; c < n_kchunks; ++c, ++it) {
                // this thread's 8 consecutive k of the chunk (half h: k 8h .. 8h+7 = MMA K-step h) from the landed ring stage
                const uint32_t g = it % SG;
                mbar_wait(smem_u32(&bar_g_full[g]), (it / SG) & 1u);
                const float* gs = g_ring + g * G_ELEMS + px;
                const int k0 = c * TC_KC;
                if (p.ln && c == 0) shift = cur.p_ok ? gs[0] : 0.f;
                float v0[KPT];
#pragma unroll
                for (int e = 0; e < KPT; ++e) {
                    const int kl = half * KPT + e;
                    v0[e] = (cur.p_ok && k0 + kl < K) ? gs[kl * TC_M] : 0.f;   // rows past K / pixels past P were not copied
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_g_empty[g]));
                const uint32_t sa = it % WS_SA;
                if (it >= (uint32_t)WS_SA) {   // MMAs of chunk it-WS_SA have read this stage
                    mbar_wait(smem_u32(&bar_mma[sa]), (it / WS_SA - 1u) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                uint32_t hi[KPT], lo[KPT];
#pragma unroll
                for (int e = 0; e < KPT; ++e) {
                    if (p.ln) {
                        const float d = (k0 + half * KPT + e < K && cur.p_ok) ? v0[e] - shift : 0.f;
                        s1 += d;
                        s2 = fmaf(d, d, s2);
                    }
                    // x = hi + lo with hi = x truncated to tf32 (1 LOP3; the MMA ignores the low 13 mantissa bits anyway) and
                    // lo = x - hi exact in fp32
                    hi[e] = __float_as_uint(v0[e]) & 0xffffe000u;
                    lo[e] = __float_as_uint(v0[e] - __uint_as_float(hi[e]));
                }
                if constexpr (KPT == 8) {
                    // The two threads of a pixel (warps w and w + 4) write the SAME tensor-memory lanes at the same time.  With their
                    // columns interleaved ([hi h0 | hi h1 | lo h0 | lo h1], two 8-column stores each) one warp's store occasionally did
                    // not land when both were released by the same late activation chunk: 32 pixels x all outputs of one tile wrong
                    // (17 of 2 500 isolated launches under HBM-saturating copy traffic, tools/op_race_probe.py; only the NT = 128 build).
                    // Each half now owns one contiguous, 16-column aligned block [hi 8 | lo 8] = exactly one MMA K-step, one store.
                    uint32_t hl[16];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        hl[e] = hi[e];
                        hl[8 + e] = lo[e];
                    }
                    tmem_st16(t_lane + (uint32_t)(A_COL0 + sa * 32 + half * 16), hl);
                } else {
                    const uint32_t a_t = t_lane + (uint32_t)(A_COL0 + sa * 32);
                    tmem_st16(a_t, hi);
                    tmem_st16(a_t + 16u, lo);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // TMEM writes -> ordered before the MMA thread's reads
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_full_a[sa]));
            }
            const int next = item + gridDim.x;
            const Item done = cur;
            if (next < total_items) cur = decode(next);

            // ---- per-output epilogue constants of this (frame, output tile) + LayerNorm partial sums -> shared memory
            if (tid < NT) {
                const int o = done.o_base + tid;
                float4 e = make_float4(0.f, 0.f, 0.f, 1.f);
                if (o < p.O) {
                    e.x = p.scale ? p.scale[o] : 1.f;
                    e.y = p.ln ? e.x * wsum[o] : 0.f;
                    e.z = (p.bias ? p.bias[o] : 0.f) + (p.pbias ? e.x * p.pbias[(long long)done.b * p.O + o] : 0.f);
                    e.w = p.gamma ? p.gamma[o] : 1.f;
                }
                s_ep[tid] = e;
            }
            if (p.ln) {
                s_ln[half][px][0] = s1;
                s_ln[half][px][1] = s2;
            }
            prod_bar<WS_PROD>();
            // y = act(rs * (scale*acc) - ms * (scale*wsum) + c)  with rs = rstd, ms = mean*rstd   (rs = 1, ms = 0 without LayerNorm)
            float rs = 1.f, ms = 0.f;
            if (p.ln) {
                const float t1 = (s_ln[0][px][0] + (HALVES == 2 ? s_ln[HALVES - 1][px][0] : 0.f)) / (float)K;
                const float t2 = (s_ln[0][px][1] + (HALVES == 2 ? s_ln[HALVES - 1][px][1] : 0.f)) / (float)K;
                rs = 1.0f / sqrtf(fmaxf(t2 - t1 * t1, 0.f) + p.ln_eps);
                ms = (shift + t1) * rs;
            }
            // ---- epilogue: thread = pixel (TMEM lane 32*(warp%4) + lane); this half's NT/2 columns, 16 at a time
            constexpr int NH = NT / HALVES;
            const int pp = done.pp;
            float* optr = p.out + (long long)done.b * p.out_bs + (long long)(done.o_base + half * NH) * P + pp;
            const float* rptr = RES ? p.res + (long long)done.b * p.res_bs + (long long)(done.o_base + half * NH) * P + pp : nullptr;
            const int o_lim = p.O - done.o_base;   // valid outputs in this tile
            // residual values are requested one 16-output block ahead - the first block before waiting for the accumulator
            // (ncu: 5.2 long-scoreboard stalls per issue on the K = 128 -> 32 layers when they were loaded at their use)
            float rr[16];
            auto load_res = [&](const float* rp, int n0) {
                if (RES && done.p_ok && n0 + 16 <= o_lim) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) rr[j] = __ldg(rp + (long long)j * P);
                }
            };
            if (RES) load_res(rptr, half * NH);
            mbar_wait(smem_u32(&bar_acc_full), (uint32_t)tile_n & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int n0 = half * NH; n0 < (half + 1) * NH && n0 < o_lim; n0 += 16) {
                uint32_t r[16];
                tmem_ld16(t_lane + (uint32_t)n0, r);
                if (p.reduce_max) {
                    // out (B, O) = max over pixels.  16 outputs x 32 lanes -> butterfly reduce-scatter: every exchange halves
                    // the outputs a lane still carries (xor 16: 8, xor 8: 4, xor 4: 2, xor 2: 1), a last xor-1 exchange
                    // completes the 32-lane max, and the even lanes each own one output: 16 shuffles and ONE atomic
                    // instruction per 16 outputs (the straightforward per-output warp max took 80 shuffles and 16 atomics).
                    float y[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float4 e = s_ep[n0 + j];
                        const float t = fmaf(rs * e.x, __uint_as_float(r[j]), fmaf(-ms, e.y, e.z));
                        y[j] = done.p_ok ? apply_act(t, ACT) : -INFINITY;
                    }
#pragma unroll
                    for (int w = 8; w >= 1; w >>= 1) {   // lanes with bit (2w) set keep the upper w outputs
                        const bool up = (lane & (2 * w)) != 0;
#pragma unroll
                        for (int i = 0; i < w; ++i) {
                            const float send = up ? y[i] : y[i + w];
                            const float keep = up ? y[i + w] : y[i];
                            y[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 2 * w));
                        }
                    }
                    y[0] = fmaxf(y[0], __shfl_xor_sync(0xffffffffu, y[0], 1));
                    const int jo = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                    if ((lane & 1) == 0 && n0 + jo < o_lim) atomic_max_float(p.out + (long long)done.b * p.out_bs + done.o_base + n0 + jo, y[0]);
                } else if (done.p_ok) {
                    if (n0 + 16 <= o_lim) {
                        // full block: no per-output predicate
                        float y[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float4 e = s_ep[n0 + j];
                            y[j] = fmaf(rs * e.x, __uint_as_float(r[j]), fmaf(-ms, e.y, e.z));
                            y[j] = apply_act(y[j], ACT);
                            if (RES) y[j] = fmaf(e.w, y[j], rr[j]);
                        }
                        // next block's residuals before this block's stores (a load cannot move above a possibly aliasing store)
                        if (RES && n0 + 16 < (half + 1) * NH) load_res(rptr + (long long)16 * P, n0 + 16);
#pragma unroll
                        for (int j = 0; j < 16; ++j) optr[(long long)j * P] = y[j];
                    } else {
#pragma unroll 1
                        for (int j = 0; j < 16 && n0 + j < o_lim; ++j) {
                            const float4 e = s_ep[n0 + j];
                            uint32_t rv = r[0];   // r[] must stay in registers: unrolled select instead of dynamic indexing
#pragma unroll
                            for (int q = 1; q < 16; ++q) rv = (j == q) ? r[q] : rv;
                            float y = fmaf(rs * e.x, __uint_as_float(rv), fmaf(-ms, e.y, e.z));
                            y = apply_act(y, ACT);
                            if (RES) y = fmaf(e.w, y, rptr[(long long)j * P]);
                            optr[(long long)j * P] = y;
                        }
                    }
                }
                optr += (long long)16 * P;
                if (RES) rptr += (long long)16 * P;
            }
            // hand the accumulator back to the MMA thread; s_ep / s_ln are rewritten by the next tile
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty));
            prod_bar<WS_PROD>();
            item = next;
        }
    }

    // ---- teardown
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
    }
}

template <int NT, int ACT, bool RES>
static int launch_ws(const AchPwConv& p, const float* w_hi, const float* w_lo, const float* wsum, cudaStream_t st) {
    const int K = p.c0 + p.c1;
    const int n_kchunks = cdiv(K, TC_KC);
    constexpr size_t smem = (size_t)WS_SB * 2 * NT * TC_KC * 4 + (size_t)ws_sg(NT) * TC_KC * TC_M * 4;
    static int ctas_per_wave_dev[ACH_MAX_DEVICES] = {};
    int& ctas_per_wave = ctas_per_wave_dev[current_device()];
    if (!ctas_per_wave) {
        cudaFuncSetAttribute(pw_conv_tc_ws_kernel<NT, ACT, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int per_sm = tc_ctas_per_sm(pw_conv_tc_ws_kernel<NT, ACT, RES>, ws_threads(NT), smem, ws_tmem_cols(NT));
        ctas_per_wave = sms * (per_sm < 1 ? 1 : per_sm);
    }
    const int n_pt = cdiv(p.P, TC_M), n_ot = cdiv(p.O, NT);
    const long long total = (long long)n_pt * n_ot * p.B;
    ACH_REQUIRE(total < (1LL << 31), "ach_pw_conv_tc: too many tiles");
    const int grid = (int)(total < ctas_per_wave ? total : ctas_per_wave);   // persistent: one wave of resident CTAs
    // programmatic dependent launch was measured inside the CUDA graph at 5.664 vs 5.674 ms per step - the graph already hides
    // launch latency - so the attribute is not set
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(ws_threads(NT));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = nullptr;
    cfg.numAttrs = 0;
    const int total_i = (int)total;
    cudaLaunchKernelEx(&cfg, pw_conv_tc_ws_kernel<NT, ACT, RES>, p, w_hi, w_lo, wsum, n_kchunks, n_pt, n_ot, total_i);
    return check_launch("ach_pw_conv_tc");
}

template <int NT>
static int launch_ws_nt(const AchPwConv& p, const float* w_hi, const float* w_lo, const float* wsum, cudaStream_t st) {
    const bool res = p.res != nullptr;
    switch (p.act) {
        case ACT_NONE: return res ? launch_ws<NT, ACT_NONE, true>(p, w_hi, w_lo, wsum, st) : launch_ws<NT, ACT_NONE, false>(p, w_hi, w_lo, wsum, st);
        case ACT_RELU: return res ? launch_ws<NT, ACT_RELU, true>(p, w_hi, w_lo, wsum, st) : launch_ws<NT, ACT_RELU, false>(p, w_hi, w_lo, wsum, st);
        case ACT_SILU: return res ? launch_ws<NT, ACT_SILU, true>(p, w_hi, w_lo, wsum, st) : launch_ws<NT, ACT_SILU, false>(p, w_hi, w_lo, wsum, st);
        case ACT_GELU: return res ? launch_ws<NT, ACT_GELU, true>(p, w_hi, w_lo, wsum, st) : launch_ws<NT, ACT_GELU, false>(p, w_hi, w_lo, wsum, st);
        default: break;
    }
    set_error("ach_pw_conv_tc: activation %d not instantiated", p.act);
    return ACH_ERR_INVALID;
}

// called by ach_pw_conv_tc (pw_conv_tc.cu) after argument validation
int pw_conv_tc_ws_launch(const AchPwConv& p, const float* w_hi, const float* w_lo, const float* wsum, cudaStream_t st) {
    const int NT = p.O <= 32 ? 32 : (p.O <= 64 ? 64 : 128);
    switch (NT) {
        case 32: return launch_ws_nt<32>(p, w_hi, w_lo, wsum, st);
        case 64: return launch_ws_nt<64>(p, w_hi, w_lo, wsum, st);
        default: return launch_ws_nt<128>(p, w_hi, w_lo, wsum, st);
    }
}

}  // namespace ach
