// RCBlock body on the tensor cores (RadarEncoder.py:65-72, dcn.py:49-63), for the high-resolution blocks (C <= 16).
//
// ncu on the SIMT kernel (rcblock.cu) showed it issue-bound with 40-60 % of its instructions in the 27-output
// offset/modulator 3x3 convolution (C*9 x 27 MACs per pixel: 7 LDS.128 + 14 FFMA2 per input value).  Here both dense
// contractions of the block become implicit GEMMs over 128-pixel tiles (thread = pixel = TMEM lane), fp32-accurate
// through the 3xTF32 split:
//   GEMM 1  om[128 px][27]  = im2col(pooled)[128][9*C] . w_om^T          (offsets + modulators, N = 32 columns)
//   (registers)  per tap: bilinear sample of every channel at p + p_k + dp_k, x 2*sigmoid(modulator)
//   GEMM 2  acc[128 px][C]  = sampled[128][9*C] . w_reg^T                 (N = 32 columns)
//   (registers)  1x1 conv + folded BN + ReLU + residual
// Both K axes are TAP-MAJOR (k = tap*C + ch) and the pooled map is CHANNEL-LAST [P][ceil4(C)], so the 3x3 window is
// 9 coalesced 16-byte loads per 4 channels and each bilinear corner one 16-byte gather (the v1 kernel read planes:
// 4 scalar gathers per channel and tap, and only tied the SIMT kernel - profiles/r1_tc_vs_simt.md).  The im2col rows /
// sampled values never leave the SM: each thread writes 4 consecutive k of its pixel as one 16-byte shared store
// into the K-major core-matrix layout, 16 k at a time, into a ring of STAGES chunk buffers; one thread issues the 6
// MMAs of a chunk and commits them to the stage's mbarrier, which is only waited for when the stage is rewritten (or
// when the accumulator is read), so gathers of chunk c+1 overlap the MMAs of chunk c.  All weight tiles (hi + lo of
// both GEMMs) are copied to shared memory once per persistent CTA.
//
// v3: the 128 pixels of a tile are a 16 x 8 block of the image and the pooled map around it (halo RCT_R) is staged in
// shared memory.  ncu on v2 (profiles/r2_ncu_rc_summary.txt): 5 M gather requests touching ~29 sectors each, l1tex
// 61-75 % busy, long-scoreboard 5.4-6.7 stalls per issue - every bilinear corner was a warp-wide gather costing one
// L1 wavefront per 128-byte line touched.  From shared memory a corner is one LDS.128 (a few bank-conflict replays).
// Taps whose 2x2 corner block leaves the staged window (offsets larger than the halo) fall back to the global gather
// per lane, so the result does not depend on the window size.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace ach {

constexpr int RCT_N = 32;   // MMA N for both GEMMs (27 offset/modulator outputs; C <= 16 conv outputs)
constexpr int RCT_ACOL = 2 * RCT_N;      // TMEM columns [0, 64): the two accumulators; from 64: A-operand stages (hi 16 | lo 16)
constexpr int RCT_TMEM_COLS = 128;
// launch bound (CTAs per SM) per channel count: tensor memory allows 4 CTAs per SM (4 x 128 columns), so nothing is gained below 128
// registers - without a bound ptxas squeezed the folded C = 3 build into 72 registers by SPILLING the prefetched residual (the spill
// store then waits for the global load: the top stall line of that build, 0.303 -> 0.326 ms); C = 12 / 16 keep 3 CTAs, C = 24 two
__host__ __device__ constexpr int rct_min_ctas(int C) { return C <= 8 ? 4 : (C <= 16 ? 3 : 2); }

constexpr int RCT_TW = 16, RCT_TH = 8;   // pixel tile (RCT_TW * RCT_TH == TC_M); a warp covers 2 rows of 16 pixels
constexpr int RCT_R = 3;                 // halo of the staged window: 1 (3x3 tap) + |offset| < 2 + the +1 bilinear corner
constexpr int RCT_WW = RCT_TW + 2 * RCT_R, RCT_WH = RCT_TH + 2 * RCT_R;

// PK = k handed to the tensor core per push (16, or 32 with a single A stage: C = 3 has K = 27, so each of its two GEMMs is ONE push -
// half the CTA barriers, tcgen05.wait::st and elected-thread MMA issues per tile)
template <int C, int STAGES, int PK>
__global__ void __launch_bounds__(128, rct_min_ctas(C)) rc_deform_tc_kernel(const AchRcDeform p, const float* __restrict__ wom_hi,
                                                           const float* __restrict__ wom_lo, const float* __restrict__ wreg_hi,
                                                           const float* __restrict__ wreg_lo, int n_tx, int n_ty, int total_items) {
    constexpr int K1 = C * 9;
    constexpr int NCH = (K1 + TC_KC - 1) / TC_KC;   // K chunks of 16 (same count for both GEMMs)
    constexpr int CP = (C + 3) & ~3;
    constexpr int Q = CP / 4;
    constexpr int B_ELEMS = RCT_N * TC_KC;
    static_assert(PK == 16 || PK == 32, "k per push");
    // v5: the host folds everything linear into the two GEMMs (engine.py:rc_tc_fold): offset bias + tap coordinate, -log2(e) on the
    // modulator rows, and 2 * BN scale * weight_conv1 into the deformable-conv weights.  With a spare k column in the last push
    // (ONES) the constants are an extra weight row against a constant-1 operand column; otherwise (C = 16) they are added from p.b_om
    // / p.bias.  Per pixel this removes the C x C epilogue GEMM (C^2 FMAs + C^2/4 shared loads), 27 bias adds, 18 int->float
    // conversions and 18 multiplies.
    constexpr bool ONES = (K1 % PK) != 0;
    static_assert(RCT_ACOL + STAGES * 2 * PK <= RCT_TMEM_COLS, "A stages do not fit the TMEM allocation");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* b_all = reinterpret_cast<float*>(smem_raw);           // [2 GEMMs][NCH][hi | lo][B_ELEMS]
    float4* win = reinterpret_cast<float4*>(b_all + 2 * NCH * 2 * B_ELEMS);   // [Q][RCT_WH][RCT_WW] pooled window (zero outside the image)
    __shared__ float s_bom[32];
    __shared__ float s_bias[CP];
    __shared__ __align__(8) uint64_t mbar[STAGES];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = (int)tc_uniform((uint32_t)(tid >> 5));   // warp index provably uniform
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(RCT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) s_bom[tid] = (tid < 27) ? p.b_om[tid] : 0.f;
    if (tid < CP) s_bias[tid] = (tid < C) ? p.bias[tid] : 0.f;
    // weight tiles: chunk c of GEMM g at b_all + ((g*NCH + c)*2 + {0: hi, 1: lo}) * B_ELEMS
    for (int i = tid; i < NCH * B_ELEMS / 4; i += 128) {
        const int c = i / (B_ELEMS / 4), r = i - c * (B_ELEMS / 4);
        float4* dst = reinterpret_cast<float4*>(b_all);
        dst[((0 * NCH + c) * 2 + 0) * (B_ELEMS / 4) + r] = __ldg(reinterpret_cast<const float4*>(wom_hi) + i);
        dst[((0 * NCH + c) * 2 + 1) * (B_ELEMS / 4) + r] = __ldg(reinterpret_cast<const float4*>(wom_lo) + i);
        dst[((1 * NCH + c) * 2 + 0) * (B_ELEMS / 4) + r] = __ldg(reinterpret_cast<const float4*>(wreg_hi) + i);
        dst[((1 * NCH + c) * 2 + 1) * (B_ELEMS / 4) + r] = __ldg(reinterpret_cast<const float4*>(wreg_lo) + i);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tc_uniform(tmem_base_s);
    const uint32_t t_lane = tmem_d + ((uint32_t)(warp * 32) << 16);
    const uint32_t b_all_s = smem_u32(b_all);
    constexpr uint32_t idesc = tf32_idesc(RCT_N);
    uint32_t n = 0;   // chunks written so far by this CTA (ring position / mbarrier phase bookkeeping)

    const int H = p.H, W = p.W, P = H * W;

    // hands the 16 k in vbuf (this thread's pixel) to the tensor core as chunk `chunk` of GEMM `g`
    auto push_chunk = [&](const float (&vbuf)[PK], int g, int chunk) {   // chunk counts pushes of PK k
        const uint32_t s = n % STAGES;
        if (n >= (uint32_t)STAGES) {   // the MMAs that read this A stage are done
            mbar_wait(smem_u32(&mbar[s]), (n / STAGES - 1u) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        // x = hi + lo, hi = x truncated to tf32 (the MMA ignores the low 13 mantissa bits anyway), lo = x - hi exact.
        // The A operand goes to TENSOR memory: this thread's lane, 16 hi columns then 16 lo columns - no shared-memory
        // store, no proxy fence, and the MMA does not re-read A through the shared-memory pipe (ncu on the shared-memory
        // A ring: l1tex 83 % busy, of which ~35 % were the tensor core's own operand reads at N = 32).
        const uint32_t a_col = (uint32_t)RCT_ACOL + s * (2u * PK);   // stage: PK hi columns, then PK lo columns
#pragma unroll
        for (int h = 0; h < PK / 16; ++h) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                hi[j] = __float_as_uint(vbuf[16 * h + j]) & 0xffffe000u;
                lo[j] = __float_as_uint(vbuf[16 * h + j] - __uint_as_float(hi[j]));
            }
            tmem_st16(t_lane + a_col + 16u * h, hi);
            tmem_st16(t_lane + a_col + PK + 16u * h, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) {   // warp-uniform branch, one elected lane issues (tc_common.cuh: no per-MMA waterfall on the warp all others wait for)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t col = tmem_d + (uint32_t)(g * RCT_N);
            if (tc_elect_one()) {
#pragma unroll
                for (int ks = 0; ks < PK / 8; ++ks) {
                    const int c16 = chunk * (PK / 16) + ks / 2;   // 16-k weight tile (tiles past NCH do not exist: their k are zero padding)
                    if (c16 >= NCH) break;
                    const uint32_t b_hi_s = b_all_s + (uint32_t)((g * NCH + c16) * 2) * B_ELEMS * 4u, b_lo_s = b_hi_s + B_ELEMS * 4u;
                    const uint32_t ah = tmem_d + a_col + (uint32_t)ks * 8u, al = ah + PK;
                    const uint64_t bh = kmajor_desc(b_hi_s, RCT_N, ks & 1), bl = kmajor_desc(b_lo_s, RCT_N, ks & 1);
                    mma_tf32_ts(col, ah, bh, idesc, (chunk == 0 && ks == 0) ? 0u : 1u);
                    mma_tf32_ts(col, al, bh, idesc, 1u);
                    mma_tf32_ts(col, ah, bl, idesc, 1u);
                }
                tc_commit(smem_u32(&mbar[s]));
            }
            __syncwarp();
        }
        ++n;
    };
    // all MMAs issued so far have completed (a commit covers every earlier MMA of the thread)
    auto wait_all = [&]() {
        const uint32_t last = n - 1u;
        mbar_wait(smem_u32(&mbar[last % STAGES]), (last / STAGES) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };

    // asynchronous (cp.async, zero-filled outside the image) copy of an item's pooled window into shared memory
    auto fill_window = [&](int tx_i, int ty_i, int b) {
        const int wy0 = ty_i * RCT_TH - RCT_R, wx0 = tx_i * RCT_TW - RCT_R;
        const float4* __restrict__ pooled = reinterpret_cast<const float4*>(p.pooled + (long long)b * p.pooled_bs);
        for (int i = tid; i < RCT_WH * RCT_WW; i += 128) {
            const int wy = i / RCT_WW, wx = i - wy * RCT_WW;
            const int yy = wy0 + wy, xx = wx0 + wx;
            const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
            const float4* src = pooled + (in ? (yy * W + xx) * Q : 0);
#pragma unroll
            for (int q = 0; q < Q; ++q)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(win + q * (RCT_WH * RCT_WW) + i)), "l"(src + q),
                             "r"(in ? 16u : 0u)
                             : "memory");
        }
    };
    // tile coordinates advance incrementally by the (pre-divided) grid stride: the three runtime divisions per tile of the direct
    // decode (I2F / MUFU.RCP / F2I sequences at the top of the loop) were the single hottest stall site in ncu's source view
    int tx_i = (int)blockIdx.x % n_tx, ty_i = ((int)blockIdx.x / n_tx) % n_ty, b = (int)blockIdx.x / (n_tx * n_ty);
    const int step_x = (int)gridDim.x % n_tx, step_y = ((int)gridDim.x / n_tx) % n_ty, step_b = (int)gridDim.x / (n_tx * n_ty);
    auto advance = [&](int& tx, int& ty, int& bb) {
        tx += step_x;
        ty += step_y;
        bb += step_b;
        if (tx >= n_tx) { tx -= n_tx; ++ty; }
        if (ty >= n_ty) { ty -= n_ty; ++bb; }
    };
    if ((int)blockIdx.x < total_items) fill_window(tx_i, ty_i, b);

#pragma unroll 1
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int ly_ = tid / RCT_TW, lx_ = tid % RCT_TW;     // position inside the tile
        const int y = ty_i * RCT_TH + ly_, x = tx_i * RCT_TW + lx_;
        const bool ok = y < H && x < W;
        const int pix = y * W + x;
        const int wy0 = ty_i * RCT_TH - RCT_R, wx0 = tx_i * RCT_TW - RCT_R;   // image coordinates of the window origin
        const float4* __restrict__ pooled = reinterpret_cast<const float4*>(p.pooled + (long long)b * p.pooled_bs);
        float* __restrict__ out_item = p.out + (long long)b * p.out_bs;   // (b advances to the next item before the epilogue)

        // residual input: requested now, consumed in the epilogue (ncu on the first v3: 24 % of all stall samples sat on
        // these loads when they were issued there)
        float xres[C];
        {
            const float* __restrict__ xr = p.x + (long long)b * p.x_bs + pix;
#pragma unroll
            for (int o = 0; o < C; ++o) xres[o] = ok ? __ldg(xr + (long long)o * P) : 0.f;
        }
        // the window of this item was requested one item ago (or before the loop)
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();

        // ---- GEMM 1: offsets / modulators = 3x3 conv of the pooled map (implicit im2col, k = tap*C + ch, zero padding)
        {
            float vbuf[PK];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float4* src = win + (ly_ + RCT_R + t / 3 - 1) * RCT_WW + (lx_ + RCT_R + t % 3 - 1);
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const float4 v4 = src[q * (RCT_WH * RCT_WW)];
                    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (q * 4 + e < C) {
                            const int k = t * C + q * 4 + e;   // compile-time
                            vbuf[k % PK] = v[e];
                            if ((k % PK) == PK - 1 || k == K1 - 1) {
                                if (k == K1 - 1) {
#pragma unroll
                                    for (int z = (k % PK) + 1; z < PK; ++z) vbuf[z] = (ONES && z == K1 % PK) ? 1.f : 0.f;
                                }
                                push_chunk(vbuf, 0, k / PK);
                            }
                        }
                    }
                }
            }
        }
        wait_all();
        float om[32];
        {
            uint32_t r[16];
            tmem_ld16(t_lane + 0u, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) om[i] = ONES ? __uint_as_float(r[i]) : __uint_as_float(r[i]) + s_bom[i];
            tmem_ld16(t_lane + 16u, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) om[16 + i] = ONES ? __uint_as_float(r[i]) : __uint_as_float(r[i]) + s_bom[16 + i];
        }

        // ---- modulated deformable sampling, k = tap*C + ch, streamed 16 k at a time into GEMM 2
        {
            float vbuf[PK];
            // converted HERE (volatile: ptxas hoisted the two conversions above GEMM 1 and then spilled them - the spill store was the
            // hottest stall line of the first folded build)
            float yf, xf;
            asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(yf) : "r"(y));
            asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(xf) : "r"(x));
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float py = yf + om[2 * t];          // (the tap's own coordinate is part of the folded offset constant)
                const float px = xf + om[2 * t + 1];
                // mask / 2 = sigmoid(z) = 1 / (1 + 2^(-z*log2 e)) on MUFU.EX2 + MUFU.RCP (relative error <= 2^-21; the denominator is >= 1,
                // and z -> -inf gives 1/inf = 0 like the exact form); om[18 + t] already is -z*log2 e, the 2 sits in the GEMM 2 weights
                const float m = rcp_approx(1.0f + ex2_approx(om[18 + t]));
                const float fy = floorf(py), fx = floorf(px);
                const int y0 = (int)fy, x0 = (int)fx;
                const float ly = py - fy, lx = px - fx;
                const float hy = 1.f - ly, hx = 1.f - lx;
                // fast path: the 2x2 corner block lies inside the staged window.  The window is zero outside the image, so
                // torchvision's per-corner validity tests are implied (an invalid corner contributes weight * 0) and the
                // four corners are immediate offsets from one shared-memory address.
                const int ry0 = y0 - wy0, rx0 = x0 - wx0;
                const bool inwin = (unsigned)ry0 < (unsigned)(RCT_WH - 1) && (unsigned)rx0 < (unsigned)(RCT_WW - 1);
                const float w00 = hy * hx, w01 = hy * lx, w10 = ly * hx, w11 = ly * lx;
                // Offsets beyond the halo are rare; the global-gather fallback is a WARP-UNIFORM branch, so the fast path carries none
                // of its instructions (ncu on the if-converted version: 36 predicated-off LDG and ~300 index instructions per tile,
                // a quarter of everything the kernel issued).
                float sv[Q][4];   // bilinear sample of every channel at this tap (before the modulator)
                if (__any_sync(0xffffffffu, !inwin)) {
                    // per-corner validity and clamped global gathers (dcn semantics spelled out); in-window lanes read the window
                    const bool in = (py > -1.f) && (py < (float)H) && (px > -1.f) && (px < (float)W);
                    const bool y0ok = in && y0 >= 0, y1ok = in && (y0 + 1 <= H - 1);
                    const bool x0ok = x0 >= 0, x1ok = (x0 + 1 <= W - 1);
                    const float v00 = (inwin || (y0ok && x0ok)) ? w00 : 0.f, v01 = (inwin || (y0ok && x1ok)) ? w01 : 0.f;
                    const float v10 = (inwin || (y1ok && x0ok)) ? w10 : 0.f, v11 = (inwin || (y1ok && x1ok)) ? w11 : 0.f;
                    const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
                    const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
                    const int i00 = (yc0 * W + xc0) * Q, i01 = (yc0 * W + xc1) * Q, i10 = (yc1 * W + xc0) * Q, i11 = (yc1 * W + xc1) * Q;
                    const float4* wbase = win + (inwin ? ry0 * RCT_WW + rx0 : 0);
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        float4 a4, b4, c4, d4;
                        if (inwin) {
                            const float4* wq = wbase + q * (RCT_WH * RCT_WW);
                            a4 = wq[0], b4 = wq[1], c4 = wq[RCT_WW], d4 = wq[RCT_WW + 1];
                        } else {
                            a4 = __ldg(pooled + i00 + q), b4 = __ldg(pooled + i01 + q), c4 = __ldg(pooled + i10 + q), d4 = __ldg(pooled + i11 + q);
                        }
                        sv[q][0] = v00 * a4.x + v01 * b4.x + v10 * c4.x + v11 * d4.x;
                        sv[q][1] = v00 * a4.y + v01 * b4.y + v10 * c4.y + v11 * d4.y;
                        sv[q][2] = v00 * a4.z + v01 * b4.z + v10 * c4.z + v11 * d4.z;
                        sv[q][3] = v00 * a4.w + v01 * b4.w + v10 * c4.w + v11 * d4.w;
                    }
                } else {
                    const float4* wbase = win + ry0 * RCT_WW + rx0;
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        const float4* wq = wbase + q * (RCT_WH * RCT_WW);
                        const float4 a4 = wq[0], b4 = wq[1], c4 = wq[RCT_WW], d4 = wq[RCT_WW + 1];
                        sv[q][0] = w00 * a4.x + w01 * b4.x + w10 * c4.x + w11 * d4.x;
                        sv[q][1] = w00 * a4.y + w01 * b4.y + w10 * c4.y + w11 * d4.y;
                        sv[q][2] = w00 * a4.z + w01 * b4.z + w10 * c4.z + w11 * d4.z;
                        sv[q][3] = w00 * a4.w + w01 * b4.w + w10 * c4.w + w11 * d4.w;
                    }
                }
                // (the chunk hand-over contains CTA barriers: it stays outside the warp-dependent branch)
#pragma unroll
                for (int q = 0; q < Q; ++q) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (q * 4 + e < C) {
                            const int k = t * C + q * 4 + e;   // compile-time
                            // mask * bilinear value, as in torchvision
                            vbuf[k % PK] = m * sv[q][e];
                            if ((k % PK) == PK - 1 || k == K1 - 1) {
                                if (k == K1 - 1) {
#pragma unroll
                                    for (int z = (k % PK) + 1; z < PK; ++z) vbuf[z] = (ONES && z == K1 % PK) ? 1.f : 0.f;
                                }
                                push_chunk(vbuf, 1, k / PK);
                            }
                        }
                    }
                }
            }
        }
        // every gather of this item precedes the barrier of the last push_chunk: the window can take the next item's data
        advance(tx_i, ty_i, b);   // (the epilogue below uses this item's pix / xres / ok, computed above)
        if (item + (int)gridDim.x < total_items) fill_window(tx_i, ty_i, b);
        wait_all();
        float acc[C > 16 ? 32 : 16];
        {
            uint32_t r[16];
            tmem_ld16(t_lane + (uint32_t)RCT_N, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(r[i]);
            if constexpr (C > 16) {
                tmem_ld16(t_lane + (uint32_t)RCT_N + 16u, r);
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[16 + i] = __uint_as_float(r[i]);
            }
        }

        // ---- (1x1 conv and BatchNorm are inside GEMM 2) ReLU + residual
        if (ok) {
            float* __restrict__ orow = out_item + pix;
#pragma unroll
            for (int o = 0; o < C; ++o) orow[(long long)o * P] = xres[o] + fmaxf(ONES ? acc[o] : acc[o] + s_bias[o], 0.f);
        }
        // the next item's first MMA overwrites the accumulators: order this item's TMEM reads before it
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(RCT_TMEM_COLS) : "memory");
}

template <int C, int STAGES, int PK>
static int launch_rc_tc_s(const AchRcDeform& p, const float* wom_hi, const float* wom_lo, const float* wreg_hi, const float* wreg_lo,
                          cudaStream_t st) {
    constexpr int NCH = (C * 9 + TC_KC - 1) / TC_KC;
    constexpr int Q = (C + 3) / 4;
    constexpr size_t smem = (size_t)(2 * NCH * 2 * RCT_N * TC_KC + Q * RCT_WH * RCT_WW * 4) * sizeof(float);
    static int ctas_per_wave_dev[ACH_MAX_DEVICES] = {};
    int& ctas_per_wave = ctas_per_wave_dev[current_device()];
    if (!ctas_per_wave) {
        cudaFuncSetAttribute(rc_deform_tc_kernel<C, STAGES, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        ctas_per_wave = sms * tc_ctas_per_sm(rc_deform_tc_kernel<C, STAGES, PK>, 128, smem, RCT_TMEM_COLS);
    }
    const int n_tx = cdiv(p.W, RCT_TW), n_ty = cdiv(p.H, RCT_TH);
    const long long total = (long long)n_tx * n_ty * p.B;
    const int grid = (int)(total < ctas_per_wave ? total : ctas_per_wave);
    rc_deform_tc_kernel<C, STAGES, PK><<<grid, 128, smem, st>>>(p, wom_hi, wom_lo, wreg_hi, wreg_lo, n_tx, n_ty, (int)total);
    return check_launch("ach_rc_deform_tc");
}

template <int C>
static int launch_rc_tc(const AchRcDeform& p, const float* wom_hi, const float* wom_lo, const float* wreg_hi, const float* wreg_lo,
                        cudaStream_t st) {
    // two A stages in tensor memory (a 1-stage build was the better one only while A still lived in shared memory: DESIGN.md §7)
    if constexpr (C * 9 <= 32) return launch_rc_tc_s<C, 1, 32>(p, wom_hi, wom_lo, wreg_hi, wreg_lo, st);   // one push per GEMM
    else return launch_rc_tc_s<C, 2, 16>(p, wom_hi, wom_lo, wreg_hi, wreg_lo, st);
}

}  // namespace ach

extern "C" int ach_rc_deform_tc_supported(int C) { return C == 3 || C == 8 || C == 12 || C == 16 || C == 24; }

extern "C" int ach_rc_deform_tc(const AchRcDeform* pp, const float* wom_hi, const float* wom_lo, const float* wreg_hi,
                                const float* wreg_lo, void* stream) {
    using namespace ach;
    const AchRcDeform& p = *pp;
    ACH_REQUIRE(p.x && p.pooled && p.b_om && p.bias && p.out && wom_hi && wom_lo && wreg_hi && wreg_lo,
                "ach_rc_deform_tc: null arg");
    ACH_REQUIRE(p.B > 0 && p.H > 0 && p.W > 0 && (long long)p.B * p.H * p.W < (1LL << 31), "ach_rc_deform_tc: bad dims");
    ACH_REQUIRE(p.pooled_cl && aligned16(p.pooled) && p.pooled_bs % 4 == 0,
                "ach_rc_deform_tc: needs the 16-byte aligned channel-last pooled map (ach_avgpool3_cl)");
    ACH_REQUIRE((long long)p.H * p.W * ((p.C + 3) / 4) < (1LL << 31), "ach_rc_deform_tc: plane too large for 32-bit indexing");
    ACH_REQUIRE(aligned16(wom_hi) && aligned16(wom_lo) && aligned16(wreg_hi) && aligned16(wreg_lo), "ach_rc_deform_tc: weight tiles must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (p.C) {
        case 3: return launch_rc_tc<3>(p, wom_hi, wom_lo, wreg_hi, wreg_lo, st);
        case 8: return launch_rc_tc<8>(p, wom_hi, wom_lo, wreg_hi, wreg_lo, st);
        case 12: return launch_rc_tc<12>(p, wom_hi, wom_lo, wreg_hi, wreg_lo, st);
        case 16: return launch_rc_tc<16>(p, wom_hi, wom_lo, wreg_hi, wreg_lo, st);
        case 24: return launch_rc_tc<24>(p, wom_hi, wom_lo, wreg_hi, wreg_lo, st);   // 20 x 20 maps: the SIMT kernel has one thread per pixel there (5 warps per SM)
        default: break;
    }
    set_error("ach_rc_deform_tc: C=%d not instantiated (3, 8, 12, 16, 24); use ach_rc_deform", p.C);
    return ACH_ERR_INVALID;
}
