// Dense spatial convolution for the small-channel layers (patchify stem / downsample, RCNet 3x3
// stride-2 convs, offset-free 3x3 convs of MobileViT).  Direct convolution: a CTA owns a 16x16 tile of
// output pixels of one frame and OT output channels; input channels are streamed in chunks through a
// shared-memory halo tile, weights (packed [Cin][k*k][ldo], output-contiguous) through a shared slab read
// as float4 broadcasts.  Each thread keeps OT accumulators for its pixel.  Epilogue: folded scale/bias,
// activation, optional channels-first LayerNorm over the outputs (EdgeNeXt stem).
#include "common.cuh"

namespace ach {

template <int OT>
__global__ void __launch_bounds__(256) conv_dense_kernel(const AchConvDense p, int CC, int tiles_x) {
    extern __shared__ __align__(16) float smem[];
    const int k = p.k, S = p.stride;
    const int IH = 15 * S + k, IW = 15 * S + k;
    const int IWp = IW | 1;
    float* ws = smem;                       // [CC][k*k][OT]
    float* tile = smem + CC * k * k * OT;   // [CC][IH][IWp]

    const int b = blockIdx.z;
    const int o_base = blockIdx.y * OT;
    const int ty0 = (blockIdx.x / tiles_x) * 16;
    const int tx0 = (blockIdx.x % tiles_x) * 16;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int iy0 = ty0 * S - p.pad, ix0 = tx0 * S - p.pad;
    const long long plane_in = (long long)p.H * p.W;
    const float* __restrict__ xb = p.x + (long long)b * p.x_bs;

    float acc[OT];
#pragma unroll
    for (int i = 0; i < OT; ++i) acc[i] = 0.f;

    const int kk = k * k;
    for (int c0 = 0; c0 < p.Cin; c0 += CC) {
        const int nc = min(CC, p.Cin - c0);
        __syncthreads();
        // weights: rows (c, tap) of OT floats
        for (int i = threadIdx.x; i < nc * kk * (OT / 4); i += 256) {
            const int row = i / (OT / 4);
            const int o4 = (i - row * (OT / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (o_base + o4 < p.ldo)
                v = __ldg(reinterpret_cast<const float4*>(p.w + ((long long)c0 * kk + row) * p.ldo + o_base + o4));
            *reinterpret_cast<float4*>(ws + row * OT + o4) = v;
        }
        const int per_ch = IH * IW;
        for (int i = threadIdx.x; i < nc * per_ch; i += 256) {
            const int c = i / per_ch;
            const int r = i - c * per_ch;
            const int yy = r / IW;
            const int xx = r - yy * IW;
            const int gy = iy0 + yy, gx = ix0 + xx;
            float v = 0.f;
            if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) v = __ldg(xb + (long long)(c0 + c) * plane_in + (long long)gy * p.W + gx);
            tile[(c * IH + yy) * IWp + xx] = v;
        }
        __syncthreads();
        for (int c = 0; c < nc; ++c) {
            const float* t = tile + (c * IH + ty * S) * IWp + tx * S;
            const float* w = ws + c * kk * OT;
            for (int ky = 0; ky < k; ++ky)
                for (int kx = 0; kx < k; ++kx) {
                    const float xv = t[ky * IWp + kx];
                    const float4* w4 = reinterpret_cast<const float4*>(w + (ky * k + kx) * OT);
#pragma unroll
                    for (int i = 0; i < OT / 4; ++i) fma4_bcast(acc + 4 * i, xv, w4[i]);
                }
        }
    }

    const int oy = ty0 + ty, ox = tx0 + tx;
    if (oy >= p.Ho || ox >= p.Wo) return;
    const long long plane_out = (long long)p.Ho * p.Wo;
    float* ob = p.out + (long long)b * p.out_bs + (long long)oy * p.Wo + ox;
    if (p.ln_out) {
        // channels-first LayerNorm over the O (<= OT) outputs of this pixel (biased variance)
        float mean = 0.f;
#pragma unroll
        for (int i = 0; i < OT; ++i) {
            if (i < p.O) {
                const float s = p.scale ? p.scale[i] : 1.f;
                const float bi = p.bias ? p.bias[i] : 0.f;
                acc[i] = apply_act(fmaf(s, acc[i], bi), p.act);
                mean += acc[i];
            }
        }
        mean /= (float)p.O;
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < OT; ++i)
            if (i < p.O) {
                const float d = acc[i] - mean;
                var = fmaf(d, d, var);
            }
        const float rstd = 1.0f / sqrtf(var / (float)p.O + p.ln_eps);
#pragma unroll
        for (int i = 0; i < OT; ++i)
            if (i < p.O) ob[(long long)i * plane_out] = p.ln_w[i] * ((acc[i] - mean) * rstd) + p.ln_b[i];
        return;
    }
#pragma unroll
    for (int i = 0; i < OT; ++i) {
        const int o = o_base + i;
        if (o < p.O) {
            const float s = p.scale ? p.scale[o] : 1.f;
            const float bi = p.bias ? p.bias[o] : 0.f;
            ob[(long long)o * plane_out] = apply_act(fmaf(s, acc[i], bi), p.act);
        }
    }
}

template <int OT>
static int launch_conv_dense(const AchConvDense& p, cudaStream_t st) {
    const int S = p.stride, k = p.k;
    const int IH = 15 * S + k, IW = 15 * S + k, IWp = IW | 1;
    const size_t per_c = (size_t)(k * k * OT + IH * IWp) * sizeof(float);
    int CC = (int)((64 * 1024) / per_c);
    if (CC < 1) CC = 1;
    CC = min(CC, p.Cin);
    const size_t smem = per_c * CC;
    ACH_REQUIRE(smem <= 200 * 1024, "ach_conv_dense: tile does not fit shared memory (k=%d s=%d)", k, S);
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(conv_dense_kernel<OT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    const int tiles_x = cdiv(p.Wo, 16), tiles_y = cdiv(p.Ho, 16);
    dim3 grid(tiles_x * tiles_y, cdiv(p.O, OT), p.B);
    conv_dense_kernel<OT><<<grid, 256, smem, st>>>(p, CC, tiles_x);
    return check_launch("ach_conv_dense");
}

}  // namespace ach

extern "C" int ach_conv_dense(const AchConvDense* pp, void* stream) {
    using namespace ach;
    const AchConvDense& p = *pp;
    ACH_REQUIRE(p.x && p.w && p.out, "ach_conv_dense: null x/w/out");
    ACH_REQUIRE(p.B > 0 && p.Cin > 0 && p.O > 0 && p.k >= 1 && p.k <= 9 && p.stride >= 1 && p.stride <= 4,
                "ach_conv_dense: bad dims");
    ACH_REQUIRE(p.Ho == (p.H + 2 * p.pad - p.k) / p.stride + 1 && p.Wo == (p.W + 2 * p.pad - p.k) / p.stride + 1,
                "ach_conv_dense: output size (%d,%d) inconsistent", p.Ho, p.Wo);
    ACH_REQUIRE(p.ldo % 4 == 0 && p.ldo >= p.O && aligned16(p.w), "ach_conv_dense: weights must be [Cin][k*k][ldo], ldo %% 4 == 0, 16B aligned");
    ACH_REQUIRE(p.B <= 65535, "ach_conv_dense: B too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p.ln_out) {
        ACH_REQUIRE(p.O <= 32 && p.ln_w && p.ln_b, "ach_conv_dense: ln_out needs O <= 32 and ln_w/ln_b");
        return launch_conv_dense<32>(p, st);
    }
    if (p.O <= 8) return launch_conv_dense<8>(p, st);
    if (p.O <= 16 || (p.O > 32 && p.O <= 48)) return launch_conv_dense<16>(p, st);
    return launch_conv_dense<32>(p, st);
}
