// Dense spatial convolution for the small-channel layers (patchify stem / downsample, RCNet 3x3
// stride-2 convs, offset-free 3x3 convs of MobileViT).  Direct convolution: a CTA owns a 16x16 tile of
// output pixels of one frame and OT output channels; input channels are streamed in chunks through a
// shared-memory halo tile, weights (packed [Cin][k*k][ldo], output-contiguous) through a shared slab read
// as float4 broadcasts.  Each thread keeps OT accumulators for its pixel.  Epilogue: folded scale/bias,
// activation, optional channels-first LayerNorm over the outputs (EdgeNeXt stem).
#include "common.cuh"
#include "tma_common.cuh"

namespace ach {

// use_tma: the halo tile [nc][IH][IWp] (IWp = IW rounded up to 4) arrives as ONE tensor-map box per channel chunk, zero-filled outside
// the image - no per-element index arithmetic, bounds tests or scalar loads (ncu on the staging loop at rc0.down: issue slots 70 %
// busy, half of all instructions in the loop); views a tensor map cannot describe keep the loop.
template <int OT>
__global__ void __launch_bounds__(256) conv_dense_kernel(const AchConvDense p, int CC, int tiles_x, const __grid_constant__ CUtensorMap tmx,
                                                         int use_tma) {
    extern __shared__ __align__(128) float smem[];
    const int k = p.k, S = p.stride;
    const int IH = 15 * S + k, IW = 15 * S + k;
    // a tensor-map box must start at a 16-byte aligned global address (x coordinate % 4 == 0, else the copy faults): the box starts
    // xo = (-pad) mod 4 columns to the left of the halo tile (tile origins are multiples of 16 * stride)
    const int xo = use_tma ? (4 - (p.pad & 3)) & 3 : 0;
    const int IWp = use_tma ? (IW + xo + 3) & ~3 : IW | 1;
    float* tile = smem;                                  // [CC][IH][IWp]  (first: a TMA destination wants 128-byte alignment)
    float* ws = smem + ((CC * IH * IWp + 31) & ~31);     // [CC][k*k][OT]
    __shared__ __align__(8) uint64_t mbar;
    if (use_tma && threadIdx.x == 0) tma_mbar_init(tma_smem_u32(&mbar), 1);
    uint32_t phase = 0;

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
        if (use_tma) {
            if (threadIdx.x == 0) {
                // the box always has CC channels: channels past Cin are zero-filled (and never read)
                tma_mbar_expect_tx(tma_smem_u32(&mbar), (uint32_t)(CC * IH * IWp) * 4u);
                tma_load_4d(tma_smem_u32(tile), &tmx, ix0 - xo, iy0, c0, b, tma_smem_u32(&mbar));
            }
            tma_mbar_wait(tma_smem_u32(&mbar), phase);
            phase ^= 1u;
        }
        const int per_ch = IH * IW;
        for (int i = threadIdx.x; !use_tma && i < nc * per_ch; i += 256) {
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
            const float* t = tile + (c * IH + ty * S) * IWp + tx * S + xo;
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

// Patchify convolution (k == stride == 4, pad 0, Cin = 3: the EdgeNeXt stem, edgenext.py:24-27) without the shared-memory halo tile:
// a thread owns one output pixel, its 4 x 4 x 3 patch arrives as 12 aligned 16-byte loads issued back to back (a warp reads 512
// contiguous bytes per channel row), the 48 x 32 weights sit in shared memory and are read as float4 broadcasts.  The generic kernel
// above staged a 64 x 65 x 3 halo tile with scalar loads per CTA: ncu showed 6.6 long-scoreboard stalls per issue and 0.8 TB/s.
__global__ void __launch_bounds__(256) patchify4_kernel(const AchConvDense p) {
    constexpr int OT = 32, CIN = 3, KK = 16;
    __shared__ __align__(16) float ws[CIN * KK * OT];
    for (int i = threadIdx.x; i < CIN * KK * OT / 4; i += 256) {
        const int row = i / (OT / 4), o4 = (i - row * (OT / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (o4 < p.ldo) v = __ldg(reinterpret_cast<const float4*>(p.w + (long long)row * p.ldo + o4));
        *reinterpret_cast<float4*>(ws + row * OT + o4) = v;
    }
    const int pix = blockIdx.x * 256 + threadIdx.x;
    const int P = p.Ho * p.Wo;
    const bool live = pix < P;
    const int oy = live ? pix / p.Wo : 0, ox = live ? pix - oy * p.Wo : 0;
    const float* __restrict__ xb = p.x + (long long)blockIdx.y * p.x_bs + (long long)(4 * oy) * p.W + 4 * ox;
    float4 in[CIN * 4];
#pragma unroll
    for (int c = 0; c < CIN; ++c)
#pragma unroll
        for (int ky = 0; ky < 4; ++ky)
            in[c * 4 + ky] = live ? __ldg(reinterpret_cast<const float4*>(xb + (long long)c * p.H * p.W + (long long)ky * p.W)) : make_float4(0.f, 0.f, 0.f, 0.f);
    // the 48 patch values go through shared memory ([k][thread], conflict-free) so that the k loop can stay ROLLED: fully unrolled the
    // kernel was > 4096 instructions of straight-line code that every warp runs exactly once (ncu: 0.64 no-instruction stalls per issue)
    extern __shared__ __align__(16) float xs_raw[];
    float (*xs)[256] = reinterpret_cast<float (*)[256]>(xs_raw);   // [CIN * KK][256], dynamic: 48 KB
#pragma unroll
    for (int r = 0; r < CIN * 4; ++r) {
        xs[r * 4 + 0][threadIdx.x] = in[r].x;
        xs[r * 4 + 1][threadIdx.x] = in[r].y;
        xs[r * 4 + 2][threadIdx.x] = in[r].z;
        xs[r * 4 + 3][threadIdx.x] = in[r].w;
    }
    __syncthreads();
    float acc[OT];
#pragma unroll
    for (int i = 0; i < OT; ++i) acc[i] = 0.f;
#pragma unroll 4
    for (int k = 0; k < CIN * KK; ++k) {   // same (channel, ky, kx) accumulation order as the generic kernel
        const float xv = xs[k][threadIdx.x];
        const float4* w4 = reinterpret_cast<const float4*>(ws + k * OT);
#pragma unroll
        for (int i = 0; i < OT / 4; ++i) fma4_bcast(acc + 4 * i, xv, w4[i]);
    }
    if (!live) return;
    float* ob = p.out + (long long)blockIdx.y * p.out_bs + pix;
    if (p.ln_out) {
        // channels-first LayerNorm over the O (<= 32) outputs of this pixel (biased variance), as in the generic kernel
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
            if (i < p.O) ob[(long long)i * P] = p.ln_w[i] * ((acc[i] - mean) * rstd) + p.ln_b[i];
        return;
    }
#pragma unroll
    for (int i = 0; i < OT; ++i) {
        if (i < p.O) {
            const float s = p.scale ? p.scale[i] : 1.f;
            const float bi = p.bias ? p.bias[i] : 0.f;
            ob[(long long)i * P] = apply_act(fmaf(s, acc[i], bi), p.act);
        }
    }
}

template <int OT>
static int launch_conv_dense(const AchConvDense& p, cudaStream_t st) {
    const int S = p.stride, k = p.k;
    const int IH = 15 * S + k, IW = 15 * S + k;
    alignas(64) CUtensorMap tmx;
    memset(&tmx, 0, sizeof(tmx));
    int use_tma = 0, CC = 1, IWp = IW | 1;
    size_t per_c = 0;
    auto size_for = [&](int pitch) {
        IWp = pitch;
        per_c = (size_t)(k * k * OT + IH * IWp) * sizeof(float);
        CC = (int)((64 * 1024) / per_c);
        if (CC < 1) CC = 1;
        CC = min(CC, p.Cin);
    };
    const int xo = (4 - (p.pad & 3)) & 3;   // the box starts at a 16-byte aligned column
    if (p.W % 4 == 0 && ((IW + xo + 3) & ~3) <= 256 && IH <= 256) {   // tensor-map staging; views it cannot describe keep the loop
        size_for((IW + xo + 3) & ~3);
        use_tma = tma_map_planes(&tmx, p.x, p.W, p.H, p.Cin, p.B, p.x_bs, IWp, IH, CC) ? 1 : 0;
    }
    if (!use_tma) size_for(IW | 1);
    const size_t smem = per_c * CC + 128;
    ACH_REQUIRE(smem <= 200 * 1024, "ach_conv_dense: tile does not fit shared memory (k=%d s=%d)", k, S);
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(conv_dense_kernel<OT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    const int tiles_x = cdiv(p.Wo, 16), tiles_y = cdiv(p.Ho, 16);
    dim3 grid(tiles_x * tiles_y, cdiv(p.O, OT), p.B);
    conv_dense_kernel<OT><<<grid, 256, smem, st>>>(p, CC, tiles_x, tmx, use_tma);
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
    if (p.ln_out) ACH_REQUIRE(p.O <= 32 && p.ln_w && p.ln_b, "ach_conv_dense: ln_out needs O <= 32 and ln_w/ln_b");
    if (p.k == 4 && p.stride == 4 && p.pad == 0 && p.Cin == 3 && p.O <= 32 && p.ldo <= 32 && p.W % 4 == 0 && p.H % 4 == 0 && aligned16(p.x) &&
        p.x_bs % 4 == 0) {
        static PerDeviceOnce patch_once;
        if (patch_once.first()) cudaFuncSetAttribute(patchify4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 256 * 4);
        patchify4_kernel<<<dim3(cdiv((long long)p.Ho * p.Wo, 256), p.B), 256, 48 * 256 * 4, st>>>(p);
        return check_launch("ach_conv_dense");
    }
    if (p.ln_out) return launch_conv_dense<32>(p, st);
    if (p.O <= 8) return launch_conv_dense<8>(p, st);
    if (p.O <= 16 || (p.O > 32 && p.O <= 48)) return launch_conv_dense<16>(p, st);
    return launch_conv_dense<32>(p, st);
}
