// Pointwise convolution / Linear / Conv1d(k=1) / bmm as a per-frame fp32 GEMM
//   OUT[O][P] = W[O][K] * X[K][P]     (X, OUT channel-major planes, P contiguous)
// with the prologues/epilogues of the reference blocks fused in (see AchPwConv in the header):
// channel concat of two sources, LayerNorm over channels, folded BN scale/bias, per-frame bias,
// ReLU/SiLU/GELU, layer-scale + residual, max over points.
//
// Tiling: one CTA = 128 pixels x (8*TM) outputs, 256 threads, each thread a TM x 4 register tile
// (4 consecutive pixels -> float4 global loads/stores, fully coalesced: a warp row covers 64 px = 256 B).
// K is streamed in chunks of 16 through double-buffered shared memory with register staging, so the
// global loads of chunk c+1 are in flight while chunk c is being multiplied.  Shared-memory reads are
// conflict-free: the X fragment is 16 distinct float4 per warp (256 contiguous bytes), the W fragment
// is a 2-address multicast.
#include "common.cuh"

namespace ach {

constexpr int PW_TP = 128;
constexpr int PW_KC = 16;

template <int TM>
__global__ void __launch_bounds__(256) pw_conv_kernel(const AchPwConv p) {
    constexpr int TO = 8 * TM;
    __shared__ __align__(16) float Xs[2][PW_KC][PW_TP];
    __shared__ __align__(16) float Ws[2][PW_KC][TO];
    __shared__ float s_mean[PW_TP];
    __shared__ float s_rstd[PW_TP];
    __shared__ float s_part[2][PW_TP];

    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int p_base = blockIdx.x * PW_TP;
    const int o_base = blockIdx.y * TO;
    const int K = p.c0 + p.c1;
    const int P = p.P;
    const float* __restrict__ x0 = p.x0 + (long long)b * p.x0_bs;
    const float* __restrict__ x1 = p.x1 ? p.x1 + (long long)b * p.x1_bs : nullptr;
    const float* __restrict__ wt = p.wt + (long long)b * p.wt_bs;

    // ---- optional LayerNorm statistics over the K channels of each of the tile's pixels (two-pass)
    if (p.ln) {
        const int px = tid & (PW_TP - 1);
        const int half = tid >> 7;
        const int pp = p_base + px;
        float s = 0.f;
        if (pp < P)
            for (int k = half; k < K; k += 2) s += (k < p.c0) ? x0[(long long)k * P + pp] : x1[(long long)(k - p.c0) * P + pp];
        s_part[half][px] = s;
        __syncthreads();
        const float mean = (s_part[0][px] + s_part[1][px]) / (float)K;
        float v = 0.f;
        if (pp < P)
            for (int k = half; k < K; k += 2) {
                const float d = ((k < p.c0) ? x0[(long long)k * P + pp] : x1[(long long)(k - p.c0) * P + pp]) - mean;
                v += d * d;
            }
        __syncthreads();
        s_part[half][px] = v;
        __syncthreads();
        if (half == 0) {
            s_mean[px] = mean;
            s_rstd[px] = 1.0f / sqrtf((s_part[0][px] + s_part[1][px]) / (float)K + p.ln_eps);
        }
        __syncthreads();
    }

    float4 xr[2];
    float4 wr = make_float4(0.f, 0.f, 0.f, 0.f);

    auto load_chunk = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * 256;
            const int row = idx >> 5;
            const int c4 = idx & 31;
            const int kk = k0 + row;
            const int pp = p_base + c4 * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kk < K && pp < P) {
                const float* src = (kk < p.c0) ? x0 + (long long)kk * P : x1 + (long long)(kk - p.c0) * P;
                v = __ldg(reinterpret_cast<const float4*>(src + pp));
                if (p.ln) {
                    const int pl = c4 * 4;
                    v.x = (v.x - s_mean[pl + 0]) * s_rstd[pl + 0];
                    v.y = (v.y - s_mean[pl + 1]) * s_rstd[pl + 1];
                    v.z = (v.z - s_mean[pl + 2]) * s_rstd[pl + 2];
                    v.w = (v.w - s_mean[pl + 3]) * s_rstd[pl + 3];
                }
            }
            xr[i] = v;
        }
        if (tid < PW_KC * TO / 4) {
            const int row = tid / (TO / 4);
            const int c4 = tid % (TO / 4);
            const int kk = k0 + row;
            const int o = o_base + c4 * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kk < K && o < p.ldw) v = __ldg(reinterpret_cast<const float4*>(wt + (long long)kk * p.ldw + o));
            wr = v;
        }
    };
    auto store_chunk = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * 256;
            *reinterpret_cast<float4*>(&Xs[buf][idx >> 5][(idx & 31) * 4]) = xr[i];
        }
        if (tid < PW_KC * TO / 4) *reinterpret_cast<float4*>(&Ws[buf][tid / (TO / 4)][(tid % (TO / 4)) * 4]) = wr;
    };

    const int warp = tid >> 5, lane = tid & 31;
    const int wp = warp & 1, wo = warp >> 1, pg = lane & 15, og = lane >> 4;
    const int ol = (wo * 2 + og) * TM;
    const int pl = wp * 64 + pg * 4;

    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

    const int nk = (K + PW_KC - 1) / PW_KC;
    load_chunk(0);
    store_chunk(0);
    __syncthreads();
    for (int c = 0; c < nk; ++c) {
        const int buf = c & 1;
        if (c + 1 < nk) load_chunk((c + 1) * PW_KC);
#pragma unroll
        for (int kk = 0; kk < PW_KC; ++kk) {
            const float4 bv = *reinterpret_cast<const float4*>(&Xs[buf][kk][pl]);
            float a[TM];
            if constexpr (TM == 8) {
                const float4 a0 = *reinterpret_cast<const float4*>(&Ws[buf][kk][ol]);
                const float4 a1 = *reinterpret_cast<const float4*>(&Ws[buf][kk][ol + 4]);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
                a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            } else if constexpr (TM == 4) {
                const float4 a0 = *reinterpret_cast<const float4*>(&Ws[buf][kk][ol]);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            } else {
                const float2 a0 = *reinterpret_cast<const float2*>(&Ws[buf][kk][ol]);
                a[0] = a0.x; a[1] = a0.y;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i) fma4_bcast(acc[i], a[i], bv);   // FFMA2: 2 issue slots per 4 MACs
        }
        if (c + 1 < nk) store_chunk((c + 1) & 1);
        __syncthreads();
    }

    // ---- epilogue
    const int pp = p_base + pl;
    const bool p_ok = pp < P;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int o = o_base + ol + i;
        const bool o_ok = o < p.O;
        const float s = (o_ok && p.scale) ? p.scale[o] : 1.f;
        const float bi = (o_ok && p.bias) ? p.bias[o] : 0.f;
        const float pb = (o_ok && p.pbias) ? p.pbias[(long long)b * p.O + o] : 0.f;
        float y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = apply_act(fmaf(s, acc[i][j] + pb, bi), p.act);
        if (p.reduce_max) {
            float m = p_ok ? fmaxf(fmaxf(y[0], y[1]), fmaxf(y[2], y[3])) : -INFINITY;
#pragma unroll
            for (int sft = 1; sft < 16; sft <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, sft));
            if (pg == 0 && o_ok) atomic_max_float(p.out + (long long)b * p.out_bs + o, m);
        } else if (o_ok && p_ok) {
            if (p.res) {
                const float4 r = *reinterpret_cast<const float4*>(p.res + (long long)b * p.res_bs + (long long)o * P + pp);
                const float g = p.gamma ? p.gamma[o] : 1.f;
                y[0] = fmaf(g, y[0], r.x);
                y[1] = fmaf(g, y[1], r.y);
                y[2] = fmaf(g, y[2], r.z);
                y[3] = fmaf(g, y[3], r.w);
            }
            *reinterpret_cast<float4*>(p.out + (long long)b * p.out_bs + (long long)o * P + pp) =
                make_float4(y[0], y[1], y[2], y[3]);
        }
    }
}

}  // namespace ach

extern "C" int ach_pw_conv(const AchPwConv* pp, void* stream) {
    using namespace ach;
    const AchPwConv& p = *pp;
    ACH_REQUIRE(p.x0 && p.wt && p.out, "ach_pw_conv: null x0/wt/out");
    ACH_REQUIRE(p.B > 0 && p.O > 0 && p.P > 0 && p.c0 > 0 && p.c1 >= 0, "ach_pw_conv: bad dims B=%d O=%d P=%d c0=%d c1=%d",
                p.B, p.O, p.P, p.c0, p.c1);
    ACH_REQUIRE((p.c1 == 0) == (p.x1 == nullptr), "ach_pw_conv: x1/c1 mismatch");
    ACH_REQUIRE(p.P % 4 == 0, "ach_pw_conv: P=%d must be a multiple of 4", p.P);
    ACH_REQUIRE(p.ldw % 4 == 0 && p.ldw >= p.O, "ach_pw_conv: ldw=%d must be a multiple of 4 and >= O=%d", p.ldw, p.O);
    ACH_REQUIRE(aligned16(p.x0) && aligned16(p.x1) && aligned16(p.wt) && aligned16(p.res) && (p.reduce_max || aligned16(p.out)),
                "ach_pw_conv: views must be 16-byte aligned");
    ACH_REQUIRE(p.x0_bs % 4 == 0 && p.x1_bs % 4 == 0 && p.wt_bs % 4 == 0 && p.res_bs % 4 == 0 && (p.reduce_max || p.out_bs % 4 == 0),
                "ach_pw_conv: batch strides must be multiples of 4 elements");
    ACH_REQUIRE(!(p.reduce_max && p.res), "ach_pw_conv: reduce_max excludes a residual");
    ACH_REQUIRE(p.B <= 65535, "ach_pw_conv: B too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p.O <= 16) {
        dim3 grid(cdiv(p.P, PW_TP), cdiv(p.O, 16), p.B);
        pw_conv_kernel<2><<<grid, 256, 0, st>>>(p);
    } else if (p.O <= 32) {
        dim3 grid(cdiv(p.P, PW_TP), cdiv(p.O, 32), p.B);
        pw_conv_kernel<4><<<grid, 256, 0, st>>>(p);
    } else {
        dim3 grid(cdiv(p.P, PW_TP), cdiv(p.O, 64), p.B);
        pw_conv_kernel<8><<<grid, 256, 0, st>>>(p);
    }
    return check_launch("ach_pw_conv");
}
