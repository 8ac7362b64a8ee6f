// Depthwise k x k convolution (k in {3,5,7,9}; stride 1/2; pad k/2), HBM-bound stencil:
// one CTA stages a (TH*s + k - s) x (TW*s + k - s) halo tile for CPB channels of one frame in shared
// memory (coalesced row loads, optional fused pre-add of a second tensor - the SDTA cascade), then each
// thread produces outputs from the staged tile with the per-channel taps held in shared memory.
// Epilogue: folded BN/bias, activation, optional broadcast post-add (positional encoding).
#include "common.cuh"

namespace ach {

template <int KS, int S>
__global__ void __launch_bounds__(256) dw_conv_kernel(const AchDwConv p, int TH, int TW, int CPB, int tiles_x) {
    extern __shared__ float smem[];
    const int IH = (TH - 1) * S + KS;
    const int IW = (TW - 1) * S + KS;
    const int IWp = IW | 1;  // odd row pitch: no bank conflicts between rows
    float* tile = smem;                   // [CPB][IH][IWp]
    float* wsm = smem + CPB * IH * IWp;   // [CPB][KS*KS]

    const int b = blockIdx.z;
    const int c_base = blockIdx.y * CPB;
    const int ty0 = (blockIdx.x / tiles_x) * TH;
    const int tx0 = (blockIdx.x % tiles_x) * TW;
    const int pad = KS / 2;
    const int iy0 = ty0 * S - pad;
    const int ix0 = tx0 * S - pad;
    const int nch = min(CPB, p.C - c_base);
    const long long plane_in = (long long)p.H * p.W;

    for (int i = threadIdx.x; i < nch * KS * KS; i += 256) wsm[i] = p.w[(long long)c_base * KS * KS + i];

    const int per_ch = IH * IW;
    for (int i = threadIdx.x; i < nch * per_ch; i += 256) {
        const int c = i / per_ch;
        const int r = i - c * per_ch;
        const int yy = r / IW;
        const int xx = r - yy * IW;
        const int gy = iy0 + yy, gx = ix0 + xx;
        float v = 0.f;
        if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
            const long long off = (long long)(c_base + c) * plane_in + (long long)gy * p.W + gx;
            v = p.x[(long long)b * p.x_bs + off];
            if (p.xadd) v += p.xadd[(long long)b * p.xadd_bs + off];
        }
        tile[(c * IH + yy) * IWp + xx] = v;
    }
    __syncthreads();

    const int th = min(TH, p.Ho - ty0);
    const int tw = min(TW, p.Wo - tx0);
    const int per_out = TH * TW;
    const long long plane_out = (long long)p.Ho * p.Wo;
    for (int i = threadIdx.x; i < nch * per_out; i += 256) {
        const int c = i / per_out;
        const int r = i - c * per_out;
        const int oy = r / TW;
        const int ox = r - oy * TW;
        if (oy >= th || ox >= tw) continue;
        const float* t = tile + (c * IH + oy * S) * IWp + ox * S;
        const float* w = wsm + c * KS * KS;
        float acc = 0.f;
#pragma unroll
        for (int ky = 0; ky < KS; ++ky)
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) acc = fmaf(t[ky * IWp + kx], w[ky * KS + kx], acc);
        const int ch = c_base + c;
        const float s = p.scale ? p.scale[ch] : 1.f;
        const float bi = p.bias ? p.bias[ch] : 0.f;
        float y = apply_act(fmaf(s, acc, bi), p.act);
        const long long po = (long long)(ty0 + oy) * p.Wo + (tx0 + ox);
        if (p.post) y += p.post[(long long)ch * plane_out + po];
        p.out[(long long)b * p.out_bs + (long long)ch * plane_out + po] = y;
    }
}

template <int KS, int S>
static int launch_dw(const AchDwConv& p, cudaStream_t st) {
    const int TW = min(p.Wo, 32);
    const int TH = min(p.Ho, 32);
    int CPB = max(1, 1024 / (TH * TW));
    CPB = min(CPB, p.C);
    const int IH = (TH - 1) * S + KS, IW = (TW - 1) * S + KS;
    const int IWp = IW | 1;
    while (CPB > 1 && (size_t)CPB * (IH * IWp + KS * KS) * 4 > 96 * 1024) --CPB;
    const size_t smem = (size_t)CPB * (IH * IWp + KS * KS) * sizeof(float);
    const int tiles_x = cdiv(p.Wo, TW), tiles_y = cdiv(p.Ho, TH);
    static bool attr_set = false;  // benign race: idempotent
    if (!attr_set) {
        cudaFuncSetAttribute(dw_conv_kernel<KS, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        attr_set = true;
    }
    dim3 grid(tiles_x * tiles_y, cdiv(p.C, CPB), p.B);
    dw_conv_kernel<KS, S><<<grid, 256, smem, st>>>(p, TH, TW, CPB, tiles_x);
    return check_launch("ach_dw_conv");
}

}  // namespace ach

extern "C" int ach_dw_conv(const AchDwConv* pp, void* stream) {
    using namespace ach;
    const AchDwConv& p = *pp;
    ACH_REQUIRE(p.x && p.w && p.out, "ach_dw_conv: null x/w/out");
    ACH_REQUIRE(p.B > 0 && p.C > 0 && p.H > 0 && p.W > 0, "ach_dw_conv: bad dims");
    ACH_REQUIRE(p.stride == 1 || p.stride == 2, "ach_dw_conv: stride %d unsupported", p.stride);
    const int pad = p.k / 2;
    ACH_REQUIRE(p.Ho == (p.H + 2 * pad - p.k) / p.stride + 1 && p.Wo == (p.W + 2 * pad - p.k) / p.stride + 1,
                "ach_dw_conv: output size (%d,%d) inconsistent with input (%d,%d) k=%d s=%d", p.Ho, p.Wo, p.H, p.W, p.k, p.stride);
    ACH_REQUIRE(p.B <= 65535 && p.C <= 65535, "ach_dw_conv: grid too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int key = p.k * 10 + p.stride;
    switch (key) {
        case 31: return launch_dw<3, 1>(p, st);
        case 32: return launch_dw<3, 2>(p, st);
        case 51: return launch_dw<5, 1>(p, st);
        case 71: return launch_dw<7, 1>(p, st);
        case 91: return launch_dw<9, 1>(p, st);
        default: break;
    }
    set_error("ach_dw_conv: k=%d stride=%d unsupported", p.k, p.stride);
    return ACH_ERR_INVALID;
}
