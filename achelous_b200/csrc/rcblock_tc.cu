// RCBlock body on the tensor cores (RadarEncoder.py:65-72, dcn.py:49-63), for the high-resolution blocks (C <= 12).
//
// ncu on the SIMT kernel (rcblock.cu) showed it issue-bound with 40-60 % of its instructions in the 27-output
// offset/modulator 3x3 convolution (C*9 x 27 MACs per pixel).  Here both dense contractions of the block become
// implicit GEMMs over 128-pixel tiles (thread = pixel = TMEM lane), fp32-accurate through the 3xTF32 split:
//   GEMM 1  om[128 px][27]  = im2col(pooled)[128][C*9] . w_om^T          (offsets + modulators, N = 32 columns)
//   (registers)  per tap: bilinear sample of every channel at p + p_k + dp_k, x 2*sigmoid(modulator)
//   GEMM 2  acc[128 px][C]  = sampled[128][9*C] . w_reg^T                 (tap-major K, N = 32 columns)
//   (registers)  1x1 conv + folded BN + ReLU + residual
// The im2col rows / sampled values never leave the SM: each thread writes the 4 consecutive k of its pixel as
// one 16-byte shared store into the K-major core-matrix layout (conflict-free, see pw_conv_tc.cu), 16 k at a
// time; one thread issues the MMAs and commits to an mbarrier.  Persistent CTAs of 128 threads, ~20 KB of shared
// memory and 64 TMEM columns each, so many CTAs share an SM and overlap each other's gather / MMA / epilogue.
#include "common.cuh"
#include "tc_common.cuh"

namespace ach {

constexpr int RCT_N = 32;   // MMA N for both GEMMs (27 offset/modulator outputs; C <= 16 conv outputs)

template <int C>
__global__ void __launch_bounds__(128) rc_deform_tc_kernel(const AchRcDeform p, const float* __restrict__ wom_hi,
                                                           const float* __restrict__ wom_lo, const float* __restrict__ wreg_hi,
                                                           const float* __restrict__ wreg_lo, int n_pt, int total_items) {
    constexpr int K1 = C * 9;
    constexpr int NCH = (K1 + TC_KC - 1) / TC_KC;   // K chunks of 16 (same count for both GEMMs)
    constexpr int CP = (C + 3) & ~3;
    __shared__ __align__(128) float a_hi[TC_KC * TC_M];
    __shared__ __align__(128) float a_lo[TC_KC * TC_M];
    __shared__ __align__(128) float b_hi[RCT_N * TC_KC];
    __shared__ __align__(128) float b_lo[RCT_N * TC_KC];
    __shared__ __align__(16) float s_w1[C * CP];    // [c][o]
    __shared__ float s_bom[32];
    __shared__ float s_scale[CP], s_bias[CP];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < C * CP; i += 128) {
        const int c = i / CP, o = i - c * CP;
        s_w1[i] = (o < C) ? p.w1[c * C + o] : 0.f;
    }
    if (tid < 32) s_bom[tid] = (tid < 27) ? p.b_om[tid] : 0.f;
    if (tid < CP) {
        s_scale[tid] = (tid < C) ? p.scale[tid] : 0.f;
        s_bias[tid] = (tid < C) ? p.bias[tid] : 0.f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    const uint32_t t_lane = tmem_d + ((uint32_t)(warp * 32) << 16);
    const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo), b_hi_s = smem_u32(b_hi), b_lo_s = smem_u32(b_lo);
    const uint32_t mbar_s = smem_u32(&mbar);
    constexpr uint32_t idesc = tf32_idesc(RCT_N);
    uint32_t commits = 0;

    const int H = p.H, W = p.W, P = H * W;

    // one K chunk: B tile copy, proxy fence, barrier, 6 MMAs into TMEM column `col`, commit, wait
    auto mma_chunk = [&](const float* __restrict__ w_h, const float* __restrict__ w_l, int chunk, uint32_t col, bool first) {
        reinterpret_cast<float4*>(b_hi)[tid] = __ldg(reinterpret_cast<const float4*>(w_h + (long long)chunk * (RCT_N * TC_KC)) + tid);
        reinterpret_cast<float4*>(b_lo)[tid] = __ldg(reinterpret_cast<const float4*>(w_l + (long long)chunk * (RCT_N * TC_KC)) + tid);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < TC_KC / 8; ++ks) {
                const uint64_t ah = kmajor_desc(a_hi_s, TC_M, ks), al = kmajor_desc(a_lo_s, TC_M, ks);
                const uint64_t bh = kmajor_desc(b_hi_s, RCT_N, ks), bl = kmajor_desc(b_lo_s, RCT_N, ks);
                mma_tf32(tmem_d + col, ah, bh, idesc, (first && ks == 0) ? 0u : 1u);
                mma_tf32(tmem_d + col, al, bh, idesc, 1u);
                mma_tf32(tmem_d + col, ah, bl, idesc, 1u);
            }
            tc_commit(mbar_s);
        }
        mbar_wait(mbar_s, commits & 1);
        commits += 1;
    };

    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int pt = item % n_pt, b = item / n_pt;
        const int pix = pt * TC_M + tid;
        const bool ok = pix < P;
        const int y = ok ? pix / W : 0, x = ok ? pix - (pix / W) * W : 0;
        const float* __restrict__ pooled = p.pooled + (long long)b * p.pooled_bs;

        // ---- GEMM 1: offsets / modulators = 3x3 conv of the pooled map (implicit im2col, k = ch*9 + tap)
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = c * TC_KC + j * 4 + e;   // compile-time
                    float t = 0.f;
                    if (k < K1) {
                        const int ch = k / 9, tap = k % 9;
                        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
                        if (ok && yy >= 0 && yy < H && xx >= 0 && xx < W) t = __ldg(pooled + (long long)ch * P + yy * W + xx);
                    }
                    v[e] = t;
                }
                split_store(a_hi + j * (TC_M * 4) + tid * 4, a_lo + j * (TC_M * 4) + tid * 4, v);
            }
            mma_chunk(wom_hi, wom_lo, c, 0u, c == 0);
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float om[32];
        {
            uint32_t r[16];
            tmem_ld16(t_lane + 0u, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) om[i] = __uint_as_float(r[i]) + s_bom[i];
            tmem_ld16(t_lane + 16u, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) om[16 + i] = __uint_as_float(r[i]) + s_bom[16 + i];
        }

        // ---- modulated deformable sampling, tap-major k = tap*C + ch, streamed 16 k at a time into GEMM 2
        float vbuf[TC_KC];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float py = (float)(y - 1 + t / 3) + om[2 * t];
            const float px = (float)(x - 1 + t % 3) + om[2 * t + 1];
            const float m = 2.0f * sigmoidf_(om[18 + t]);
            const float fy = floorf(py), fx = floorf(px);
            const int y0 = (int)fy, x0 = (int)fx;
            const float ly = py - fy, lx = px - fx;
            const float hy = 1.f - ly, hx = 1.f - lx;
            const bool in = ok && (py > -1.f) && (py < (float)H) && (px > -1.f) && (px < (float)W);
            const bool y0ok = in && y0 >= 0, y1ok = in && (y0 + 1 <= H - 1);
            const bool x0ok = x0 >= 0, x1ok = (x0 + 1 <= W - 1);
            const float w00 = (y0ok && x0ok) ? m * (hy * hx) : 0.f;
            const float w01 = (y0ok && x1ok) ? m * (hy * lx) : 0.f;
            const float w10 = (y1ok && x0ok) ? m * (ly * hx) : 0.f;
            const float w11 = (y1ok && x1ok) ? m * (ly * lx) : 0.f;
            const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
            const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
            const int i00 = yc0 * W + xc0, i01 = yc0 * W + xc1, i10 = yc1 * W + xc0, i11 = yc1 * W + xc1;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                const int k = t * C + ch;                    // compile-time
                const float* pc = pooled + (long long)ch * P;
                vbuf[k % TC_KC] = w00 * __ldg(pc + i00) + w01 * __ldg(pc + i01) + w10 * __ldg(pc + i10) + w11 * __ldg(pc + i11);
                if ((k % TC_KC) == TC_KC - 1 || k == K1 - 1) {
                    // chunk complete (the last one is zero padded): hand it to the tensor core
                    const int chunk = k / TC_KC;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float v[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[e] = (chunk * TC_KC + j * 4 + e < K1) ? vbuf[j * 4 + e] : 0.f;
                        split_store(a_hi + j * (TC_M * 4) + tid * 4, a_lo + j * (TC_M * 4) + tid * 4, v);
                    }
                    mma_chunk(wreg_hi, wreg_lo, chunk, (uint32_t)RCT_N, chunk == 0);
                }
            }
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float acc[16];
        {
            uint32_t r[16];
            tmem_ld16(t_lane + (uint32_t)RCT_N, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(r[i]);
        }

        // ---- 1x1 conv + folded BN + ReLU + residual
        if (ok) {
            float z[CP];
#pragma unroll
            for (int o = 0; o < CP; ++o) z[o] = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float4* w4 = reinterpret_cast<const float4*>(s_w1 + c * CP);
#pragma unroll
                for (int i = 0; i < CP / 4; ++i) {
                    const float4 wv = w4[i];
                    z[4 * i + 0] = fmaf(acc[c], wv.x, z[4 * i + 0]);
                    z[4 * i + 1] = fmaf(acc[c], wv.y, z[4 * i + 1]);
                    z[4 * i + 2] = fmaf(acc[c], wv.z, z[4 * i + 2]);
                    z[4 * i + 3] = fmaf(acc[c], wv.w, z[4 * i + 3]);
                }
            }
            const float* __restrict__ xr = p.x + (long long)b * p.x_bs + pix;
            float* __restrict__ orow = p.out + (long long)b * p.out_bs + pix;
#pragma unroll
            for (int o = 0; o < C; ++o) orow[(long long)o * P] = xr[(long long)o * P] + fmaxf(fmaf(s_scale[o], z[o], s_bias[o]), 0.f);
        }
        // the next item's first MMA overwrites the accumulators: order this item's TMEM reads before it
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(64) : "memory");
}

template <int C>
static int launch_rc_tc(const AchRcDeform& p, const float* wom_hi, const float* wom_lo, const float* wreg_hi, const float* wreg_lo,
                        cudaStream_t st) {
    static int ctas_per_wave = 0;
    if (!ctas_per_wave) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        ctas_per_wave = sms * tc_ctas_per_sm(rc_deform_tc_kernel<C>, 128, 0, 64);
    }
    const int n_pt = cdiv((long long)p.H * p.W, TC_M);
    const long long total = (long long)n_pt * p.B;
    const int grid = (int)(total < ctas_per_wave ? total : ctas_per_wave);
    rc_deform_tc_kernel<C><<<grid, 128, 0, st>>>(p, wom_hi, wom_lo, wreg_hi, wreg_lo, n_pt, (int)total);
    return check_launch("ach_rc_deform_tc");
}

}  // namespace ach

extern "C" int ach_rc_deform_tc_supported(int C) { return C == 3 || C == 8 || C == 12; }

extern "C" int ach_rc_deform_tc(const AchRcDeform* pp, const float* wom_hi, const float* wom_lo, const float* wreg_hi,
                                const float* wreg_lo, void* stream) {
    using namespace ach;
    const AchRcDeform& p = *pp;
    ACH_REQUIRE(p.x && p.pooled && p.b_om && p.w1 && p.scale && p.bias && p.out && wom_hi && wom_lo && wreg_hi && wreg_lo,
                "ach_rc_deform_tc: null arg");
    ACH_REQUIRE(p.B > 0 && p.H > 0 && p.W > 0 && (long long)p.B * p.H * p.W < (1LL << 31), "ach_rc_deform_tc: bad dims");
    ACH_REQUIRE(!p.pooled_cl, "ach_rc_deform_tc: needs the channel-major pooled map (ach_avgpool3)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (p.C) {
        case 3: return launch_rc_tc<3>(p, wom_hi, wom_lo, wreg_hi, wreg_lo, st);
        case 8: return launch_rc_tc<8>(p, wom_hi, wom_lo, wreg_hi, wreg_lo, st);
        case 12: return launch_rc_tc<12>(p, wom_hi, wom_lo, wreg_hi, wreg_lo, st);
        default: break;
    }
    set_error("ach_rc_deform_tc: C=%d not instantiated (3, 8, 12); use ach_rc_deform", p.C);
    return ACH_ERR_INVALID;
}
