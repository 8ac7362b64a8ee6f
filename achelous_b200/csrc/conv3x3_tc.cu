// Dense 3x3 convolution (stride 1, pad 1) as an implicit GEMM on tcgen05, fp32-accurate through the 3xTF32 split.
// Used for the CSP-Dual-FPN Bottlenecks (neck/cspdualfpn.py:42-56: 16->32 at 320^2 is 4 608 MACs per pixel) and the
// MobileViT block convolutions (mobilevit.py:7-21) - the SIMT direct convolution (conv_dense.cu) needs 2.1 ms per
// 16->32 layer at 320^2 x 64 frames.
//
// Same warp-specialised pipeline as pw_conv_tc_ws.cu, with the K axis = (input-channel group of 16, tap, channel):
//   * warp 5 ("window loader"): per (tile, channel group) stages the 10 x 18 halo window of 16 input channels in a
//     shared-memory ring with 4-byte cp.async (zero-filled outside the image / past Cin), completion tracked by
//     cp.async.mbarrier.arrive on the stage's "full" barrier.
//   * warps 0-3 (thread = output pixel of a 16 x 8 tile = TMEM lane): for each of the 9 taps read the pixel's 16
//     channels from the window at the tap's offset (the implicit im2col row), split into tf32 hi/lo and write them
//     to an A-operand stage in TENSOR memory (tcgen05.st); later run the epilogue (folded BN, activation, optional
//     residual = the Bottleneck's "+ x") with coalesced 64-byte row stores.
//   * warp 4, one thread: streams the packed hi/lo weight tile of chunk (group, tap) by TMA (cp.async.bulk) through
//     a 4-stage ring and issues the 6 tcgen05.mma of the chunk with A from TMEM, B from shared memory.
// Weights are packed by ach_pack_pw_tc from a K-major matrix whose rows follow the chunk order:
//   k = (g * 9 + tap) * 16 + c   for input channel 16 g + c (zero rows past Cin).
#include "common.cuh"
#include "tc_common.cuh"

namespace ach {

constexpr int C3_TW = 16, C3_TH = 8;                 // output tile (C3_TW * C3_TH == TC_M)
constexpr int C3_WW = C3_TW + 2, C3_WH = C3_TH + 2;  // window with the 3x3 halo
constexpr int C3_WIN = TC_KC * C3_WH * C3_WW;        // floats per window stage (16 channels)
constexpr int C3_SW = 3;                             // window ring stages
constexpr int C3_SA = 2;                             // A-operand stages in tensor memory
constexpr int C3_SB = 4;                             // weight ring stages
constexpr int C3_PF = 2;                             // weight chunks in flight ahead of the MMA
constexpr int C3_PROD = 128;
constexpr int C3_THREADS = C3_PROD + 64;

__host__ __device__ constexpr int c3_tmem_cols(int nt) {
    const int need = nt + C3_SA * 32;
    return need <= 64 ? 64 : need <= 128 ? 128 : 256;
}

__device__ __forceinline__ void c3_mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void c3_mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}

template <int NT, int ACT, bool RES>
__global__ void __launch_bounds__(C3_THREADS, NT == 128 ? 2 : 3)
    conv3x3_tc_kernel(const AchConv3x3Tc p, const float* __restrict__ w_hi, const float* __restrict__ w_lo, int n_cg, int n_tx, int n_ty,
                      int n_ot, int total_items) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int B_ELEMS = NT * TC_KC;
    float* b_ring = reinterpret_cast<float*>(smem_raw);          // [SB][b_hi | b_lo]
    float* w_ring = b_ring + C3_SB * 2 * B_ELEMS;                // [SW][16][C3_WH][C3_WW]
    __shared__ __align__(8) uint64_t bar_full_a[C3_SA], bar_full_b[C3_SB], bar_mma[C3_SA], bar_w_full[C3_SW], bar_w_empty[C3_SW],
        bar_acc_full, bar_acc_empty;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float2 s_ep[NT];   // per output of the current tile: {scale, bias}

    const int tid = threadIdx.x, warp = (int)tc_uniform((uint32_t)(tid >> 5)), lane = tid & 31;   // warp index provably uniform
    const int H = p.H, W = p.W, P = H * W;
    constexpr int A_COL0 = NT;
    constexpr int TMEM_COLS = c3_tmem_cols(NT);
    const int n_chunks = n_cg * 9;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < C3_SA; ++i) {
            c3_mbar_init(smem_u32(&bar_full_a[i]), C3_PROD / 32);
            c3_mbar_init(smem_u32(&bar_mma[i]), 1);
        }
        for (int i = 0; i < C3_SB; ++i) c3_mbar_init(smem_u32(&bar_full_b[i]), 1);
        for (int i = 0; i < C3_SW; ++i) {
            c3_mbar_init(smem_u32(&bar_w_full[i]), 32);            // one cp.async arrive per loader lane
            c3_mbar_init(smem_u32(&bar_w_empty[i]), C3_PROD / 32);
        }
        c3_mbar_init(smem_u32(&bar_acc_full), 1);
        c3_mbar_init(smem_u32(&bar_acc_empty), C3_PROD / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    const uint32_t b_ring_s = smem_u32(b_ring);

    if (warp == C3_PROD / 32) {
        // ================================================================== MMA + weight-TMA warp (warp-uniform loop, one elected lane issues: tc_common.cuh)
        {
            const bool elected = tc_elect_one();
            const uint32_t tmem_u = tc_uniform(tmem_d);
            constexpr uint32_t idesc = tf32_idesc(NT);
            constexpr uint32_t B_LBO = (NT / 8) * 128;
            int pit = 0, p_item = blockIdx.x, p_c = 0;
            auto issue_b = [&]() {
                if (p_item >= total_items) return;
                const int sb = pit % C3_SB;
                const int prev = pit - C3_SB;
                if (prev >= 0) mbar_wait(smem_u32(&bar_mma[prev % C3_SA]), (uint32_t)(prev / C3_SA) & 1u);
                const long long blk = ((long long)(p_item % n_ot) * n_chunks + p_c) * B_ELEMS;
                const uint32_t full = smem_u32(&bar_full_b[sb]);
                const uint32_t dst = b_ring_s + (uint32_t)sb * 2u * B_ELEMS * 4u;
                if (elected) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(2u * B_ELEMS * 4u) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                                 "l"(w_hi + blk), "r"(B_ELEMS * 4u), "r"(full)
                                 : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + B_ELEMS * 4u),
                                 "l"(w_lo + blk), "r"(B_ELEMS * 4u), "r"(full)
                                 : "memory");
                }
                ++pit;
                if (++p_c == n_chunks) {
                    p_c = 0;
                    p_item += gridDim.x;
                }
            };
#pragma unroll 1
            for (int i = 0; i < C3_PF; ++i) issue_b();
            int it = 0, tile_n = 0;
#pragma unroll 1
            for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++tile_n) {
                if (tile_n > 0) {
                    mbar_wait(smem_u32(&bar_acc_empty), (uint32_t)(tile_n - 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
#pragma unroll 1
                for (int c = 0; c < n_chunks; ++c, ++it) {
                    issue_b();
                    const int sa = it % C3_SA, sb = it % C3_SB;
                    mbar_wait(smem_u32(&bar_full_b[sb]), (uint32_t)(it / C3_SB) & 1u);
                    mbar_wait(smem_u32(&bar_full_a[sa]), (uint32_t)(it / C3_SA) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi_t = tmem_u + (uint32_t)(A_COL0 + sa * 32), a_lo_t = a_hi_t + 16u;
                    const uint32_t b_hi_s = b_ring_s + (uint32_t)sb * 2u * B_ELEMS * 4u;
                    const uint64_t dh = make_desc(b_hi_s, B_LBO, 128u, 0), dl = make_desc(b_hi_s + B_ELEMS * 4u, B_LBO, 128u, 0);
                    if (elected) {
#pragma unroll
                        for (int ks = 0; ks < TC_KC / 8; ++ks) {
                            const uint32_t ah = a_hi_t + (uint32_t)ks * 8u, al = a_lo_t + (uint32_t)ks * 8u;
                            const uint64_t off = (uint64_t)((ks * 2 * B_LBO) >> 4);   // the descriptor's address field counts 16-byte units
                            mma_tf32_ts(tmem_u, ah, dh + off, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                            mma_tf32_ts(tmem_u, al, dh + off, idesc, 1u);
                            mma_tf32_ts(tmem_u, ah, dl + off, idesc, 1u);
                        }
                        tc_commit(smem_u32(&bar_mma[sa]));
                        if (c == n_chunks - 1) tc_commit(smem_u32(&bar_acc_full));
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    } else if (warp == C3_PROD / 32 + 1) {
        // ================================================================== window loader (32 lanes, 4-byte cp.async)
        uint32_t wit = 0;
#pragma unroll 1
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const int t_i = item / n_ot;
            const int tx_i = t_i % n_tx, ty_i = (t_i / n_tx) % n_ty, b = t_i / (n_tx * n_ty);
            const int gy0 = ty_i * C3_TH - 1, gx0 = tx_i * C3_TW - 1;
            const float* __restrict__ xb = p.x + (long long)b * p.x_bs;
#pragma unroll 1
            for (int g = 0; g < n_cg; ++g, ++wit) {
                const uint32_t s = wit % C3_SW;
                if (wit >= (uint32_t)C3_SW) mbar_wait(smem_u32(&bar_w_empty[s]), (wit / C3_SW - 1u) & 1u);
                const uint32_t dst0 = smem_u32(w_ring + s * C3_WIN);
#pragma unroll 4
                for (int idx = lane; idx < C3_WIN; idx += 32) {     // [c][wy][wx] order, 18-float rows -> 72-byte global segments
                    const int c = idx / (C3_WH * C3_WW), rem = idx - c * (C3_WH * C3_WW);
                    const int wy = rem / C3_WW, wx = rem - wy * C3_WW;
                    const int ch = g * TC_KC + c, gy = gy0 + wy, gx = gx0 + wx;
                    const bool ok = ch < p.Cin && gy >= 0 && gy < H && gx >= 0 && gx < W;
                    const float* src = ok ? xb + (long long)ch * P + gy * W + gx : xb;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst0 + (uint32_t)idx * 4u), "l"(src), "r"(ok ? 4u : 0u)
                                 : "memory");
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bar_w_full[s])) : "memory");
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
        // ================================================================== producers + epilogue (128 threads)
        const int ly = tid / C3_TW, lx = tid % C3_TW;
        const uint32_t t_lane = tmem_d + ((uint32_t)(warp * 32) << 16);
        uint32_t it = 0, wit = 0;
        int tile_n = 0;
#pragma unroll 1
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++tile_n) {
            const int o_tile = item % n_ot, t_i = item / n_ot;
            const int tx_i = t_i % n_tx, ty_i = (t_i / n_tx) % n_ty, b = t_i / (n_tx * n_ty);
            const int y = ty_i * C3_TH + ly, x = tx_i * C3_TW + lx;
            const bool p_ok = y < H && x < W;
            const int o_base = o_tile * NT;
#pragma unroll 1
            for (int g = 0; g < n_cg; ++g, ++wit) {
                const uint32_t s = wit % C3_SW;
                mbar_wait(smem_u32(&bar_w_full[s]), (wit / C3_SW) & 1u);
                const float* win = w_ring + s * C3_WIN + ly * C3_WW + lx;
#pragma unroll 1
                for (int t = 0; t < 9; ++t, ++it) {
                    const float* wt = win + (t / 3) * C3_WW + (t % 3);
                    float v[TC_KC];
#pragma unroll
                    for (int c = 0; c < TC_KC; ++c) v[c] = wt[c * (C3_WH * C3_WW)];
                    const uint32_t sa = it % C3_SA;
                    if (it >= (uint32_t)C3_SA) {   // MMAs of chunk it-SA have read this A stage
                        mbar_wait(smem_u32(&bar_mma[sa]), (it / C3_SA - 1u) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    uint32_t hi[TC_KC], lo[TC_KC];
#pragma unroll
                    for (int c = 0; c < TC_KC; ++c) {
                        hi[c] = __float_as_uint(v[c]) & 0xffffe000u;
                        lo[c] = __float_as_uint(v[c] - __uint_as_float(hi[c]));
                    }
                    const uint32_t a_t = t_lane + (uint32_t)(A_COL0 + sa * 32);
                    tmem_st16(a_t, hi);
                    tmem_st16(a_t + 16u, lo);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) c3_mbar_arrive(smem_u32(&bar_full_a[sa]));
                }
                __syncwarp();
                if (lane == 0) c3_mbar_arrive(smem_u32(&bar_w_empty[s]));   // all 9 taps of this window have been read
            }

            // ---- epilogue constants, then the accumulator
            if (tid < NT) {
                const int o = o_base + tid;
                s_ep[tid] = o < p.O ? make_float2(p.scale ? p.scale[o] : 1.f, p.bias ? p.bias[o] : 0.f) : make_float2(0.f, 0.f);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(C3_PROD) : "memory");
            const int pix = y * W + x;
            float* optr = p.out + (long long)b * p.out_bs + (long long)o_base * P + pix;
            const float* rptr = RES ? p.res + (long long)b * p.res_bs + (long long)o_base * P + pix : nullptr;
            const int o_lim = p.O - o_base;
            mbar_wait(smem_u32(&bar_acc_full), (uint32_t)tile_n & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int n0 = 0; n0 < NT && n0 < o_lim; n0 += 16) {
                uint32_t r[16];
                tmem_ld16(t_lane + (uint32_t)n0, r);
                if (p_ok) {
                    if (n0 + 16 <= o_lim) {
                        float rr[16];
                        if (RES) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) rr[j] = __ldg(rptr + (long long)j * P);
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float2 e = s_ep[n0 + j];
                            float yv = apply_act(fmaf(e.x, __uint_as_float(r[j]), e.y), ACT);
                            if (RES) yv += rr[j];
                            optr[(long long)j * P] = yv;
                        }
                    } else {
#pragma unroll 1
                        for (int j = 0; j < 16 && n0 + j < o_lim; ++j) {
                            const float2 e = s_ep[n0 + j];
                            uint32_t rv = r[0];
#pragma unroll
                            for (int q = 1; q < 16; ++q) rv = (j == q) ? r[q] : rv;
                            float yv = apply_act(fmaf(e.x, __uint_as_float(rv), e.y), ACT);
                            if (RES) yv += __ldg(rptr + (long long)j * P);
                            optr[(long long)j * P] = yv;
                        }
                    }
                }
                optr += (long long)16 * P;
                if (RES) rptr += (long long)16 * P;
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) c3_mbar_arrive(smem_u32(&bar_acc_empty));
            asm volatile("bar.sync 1, %0;" ::"n"(C3_PROD) : "memory");
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
static int launch_c3(const AchConv3x3Tc& p, const float* w_hi, const float* w_lo, cudaStream_t st) {
    constexpr size_t smem = (size_t)C3_SB * 2 * NT * TC_KC * 4 + (size_t)C3_SW * C3_WIN * 4;
    static int ctas_per_wave_dev[ACH_MAX_DEVICES] = {};
    int& ctas_per_wave = ctas_per_wave_dev[current_device()];
    if (!ctas_per_wave) {
        cudaFuncSetAttribute(conv3x3_tc_kernel<NT, ACT, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int per_sm = tc_ctas_per_sm(conv3x3_tc_kernel<NT, ACT, RES>, C3_THREADS, smem, c3_tmem_cols(NT));
        ctas_per_wave = sms * (per_sm < 1 ? 1 : per_sm);
    }
    const int n_cg = cdiv(p.Cin, TC_KC), n_tx = cdiv(p.W, C3_TW), n_ty = cdiv(p.H, C3_TH), n_ot = cdiv(p.O, NT);
    const long long total = (long long)n_tx * n_ty * n_ot * p.B;
    ACH_REQUIRE(total < (1LL << 31), "ach_conv3x3_tc: too many tiles");
    const int grid = (int)(total < ctas_per_wave ? total : ctas_per_wave);
    conv3x3_tc_kernel<NT, ACT, RES><<<grid, C3_THREADS, smem, st>>>(p, w_hi, w_lo, n_cg, n_tx, n_ty, n_ot, (int)total);
    return check_launch("ach_conv3x3_tc");
}

template <int NT>
static int launch_c3_nt(const AchConv3x3Tc& p, const float* w_hi, const float* w_lo, cudaStream_t st) {
    const bool res = p.res != nullptr;
    switch (p.act) {
        case ACT_NONE: return res ? launch_c3<NT, ACT_NONE, true>(p, w_hi, w_lo, st) : launch_c3<NT, ACT_NONE, false>(p, w_hi, w_lo, st);
        case ACT_RELU: return res ? launch_c3<NT, ACT_RELU, true>(p, w_hi, w_lo, st) : launch_c3<NT, ACT_RELU, false>(p, w_hi, w_lo, st);
        case ACT_SILU: return res ? launch_c3<NT, ACT_SILU, true>(p, w_hi, w_lo, st) : launch_c3<NT, ACT_SILU, false>(p, w_hi, w_lo, st);
        default: break;
    }
    set_error("ach_conv3x3_tc: activation %d not instantiated", p.act);
    return ACH_ERR_INVALID;
}

}  // namespace ach

extern "C" int ach_conv3x3_tc_k(int Cin) { return ((Cin + 15) / 16) * 9 * 16; }

extern "C" int ach_conv3x3_tc(const AchConv3x3Tc* pp, const float* w_hi, const float* w_lo, void* stream) {
    using namespace ach;
    const AchConv3x3Tc& p = *pp;
    ACH_REQUIRE(p.x && p.out && w_hi && w_lo, "ach_conv3x3_tc: null x/out/weights");
    ACH_REQUIRE(p.B > 0 && p.Cin > 0 && p.O > 0 && p.H > 0 && p.W > 0, "ach_conv3x3_tc: bad dims");
    ACH_REQUIRE(aligned16(w_hi) && aligned16(w_lo), "ach_conv3x3_tc: weight tiles must be 16-byte aligned");
    ACH_REQUIRE((long long)p.H * p.W < (1LL << 30), "ach_conv3x3_tc: plane too large for 32-bit indexing");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int NT = p.O <= 32 ? 32 : (p.O <= 64 ? 64 : 128);   // == the tile width ach_pack_pw_tc chose for O outputs
    switch (NT) {
        case 32: return launch_c3_nt<32>(p, w_hi, w_lo, st);
        case 64: return launch_c3_nt<64>(p, w_hi, w_lo, st);
        default: return launch_c3_nt<128>(p, w_hi, w_lo, st);
    }
}
