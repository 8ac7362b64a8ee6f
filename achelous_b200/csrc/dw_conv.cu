// Depthwise k x k convolution (k in {3,5,7,9}; stride 1/2; pad k/2), HBM-bound stencil.
//
// One CTA stages the halo tile of CPB channels of one frame in shared memory (warp-per-row coalesced
// loads, optional fused pre-add of a second tensor - the SDTA cascade `conv(sp + spx[i])`), then every
// thread produces a strip of 4 horizontally adjacent outputs: per kernel row it pulls the
// (3*S + KS)-wide input segment with conflict-free 128-bit shared loads into registers and reuses it
// for all 4 outputs and KS taps (0.3-0.5 shared loads per FMA instead of 2), with the KS*KS taps of the
// channel held in registers.  Epilogue: folded BN/bias, activation, optional broadcast post-add
// (positional encoding), 128-bit stores when the row pitch allows.
#include "common.cuh"

namespace ach {

constexpr int DW_NX = 4;  // outputs per thread along x

template <int KS, int S>
__global__ void __launch_bounds__(256) dw_conv_kernel(const AchDwConv p, int TH, int TW, int CPB, int tiles_x) {
    extern __shared__ __align__(16) float smem[];
    constexpr int SEG = (DW_NX - 1) * S + KS;      // input values per kernel row per thread
    constexpr int SEG4 = (SEG + 3) / 4;            // as float4 loads
    const int IH = (TH - 1) * S + KS;
    const int IW = (TW - 1) * S + KS;
    const int IWp = ((IW + 3) & ~3) + 4;           // 16B-aligned rows + slack for the last segment's over-read
    float* tile = smem;                            // [CPB][IH][IWp]
    float* wsm = smem + CPB * IH * IWp;            // [CPB][KS*KS]

    const int b = blockIdx.z;
    const int c_base = blockIdx.y * CPB;
    const int ty0 = (blockIdx.x / tiles_x) * TH;
    const int tx0 = (blockIdx.x % tiles_x) * TW;
    constexpr int pad = KS / 2;
    const int iy0 = ty0 * S - pad;
    const int ix0 = tx0 * S - pad;
    const int nch = min(CPB, p.C - c_base);
    const long long plane_in = (long long)p.H * p.W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < nch * KS * KS; i += 256) wsm[i] = p.w[(long long)c_base * KS * KS + i];

    // ---- stage the halo tile: one warp per (channel, row)
    const float* xb = p.x + (long long)b * p.x_bs;
    const float* ab = p.xadd ? p.xadd + (long long)b * p.xadd_bs : nullptr;
    for (int r = warp; r < nch * IH; r += 8) {
        const int c = r / IH;
        const int yy = r - c * IH;
        const int gy = iy0 + yy;
        const bool row_ok = gy >= 0 && gy < p.H;
        const long long roff = (long long)(c_base + c) * plane_in + (long long)gy * p.W;
        float* trow = tile + (c * IH + yy) * IWp;
        for (int xx = lane; xx < IWp; xx += 32) {
            const int gx = ix0 + xx;
            float v = 0.f;
            if (row_ok && xx < IW && gx >= 0 && gx < p.W) {
                v = xb[roff + gx];
                if (ab) v += ab[roff + gx];
            }
            trow[xx] = v;
        }
    }
    __syncthreads();

    const int th = min(TH, p.Ho - ty0);
    const int tw = min(TW, p.Wo - tx0);
    const int strips = (TW + DW_NX - 1) / DW_NX;
    const int per_ch = TH * strips;
    const long long plane_out = (long long)p.Ho * p.Wo;
    const bool vec_store = (p.Wo % 4 == 0) && (tx0 % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) && (p.out_bs % 4 == 0);
    for (int i = threadIdx.x; i < nch * per_ch; i += 256) {
        const int c = i / per_ch;
        const int r = i - c * per_ch;
        const int oy = r / strips;
        const int ox = (r - oy * strips) * DW_NX;
        if (oy >= th || ox >= tw) continue;
        float wk[KS * KS];
#pragma unroll
        for (int t = 0; t < KS * KS; ++t) wk[t] = wsm[c * KS * KS + t];
        float acc[DW_NX];
#pragma unroll
        for (int j = 0; j < DW_NX; ++j) acc[j] = 0.f;
        const float* t0 = tile + (c * IH + oy * S) * IWp + ox * S;   // 16B aligned: ox*S multiple of 4
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
            float seg[SEG4 * 4];
#pragma unroll
            for (int q = 0; q < SEG4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(t0 + ky * IWp + 4 * q);
                seg[4 * q + 0] = v.x; seg[4 * q + 1] = v.y; seg[4 * q + 2] = v.z; seg[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int kx = 0; kx < KS; ++kx)
#pragma unroll
                for (int j = 0; j < DW_NX; ++j) acc[j] = fmaf(seg[j * S + kx], wk[ky * KS + kx], acc[j]);
        }
        const int ch = c_base + c;
        const float s = p.scale ? p.scale[ch] : 1.f;
        const float bi = p.bias ? p.bias[ch] : 0.f;
        const long long po = (long long)(ty0 + oy) * p.Wo + (tx0 + ox);
        float y[DW_NX];
#pragma unroll
        for (int j = 0; j < DW_NX; ++j) {
            y[j] = apply_act(fmaf(s, acc[j], bi), p.act);
            if (p.post && ox + j < tw) y[j] += p.post[(long long)ch * plane_out + po + j];
        }
        float* op = p.out + (long long)b * p.out_bs + (long long)ch * plane_out + po;
        if (vec_store && ox + DW_NX <= tw) {
            *reinterpret_cast<float4*>(op) = make_float4(y[0], y[1], y[2], y[3]);
        } else {
#pragma unroll
            for (int j = 0; j < DW_NX; ++j)
                if (ox + j < tw) op[j] = y[j];
        }
    }
}

template <int KS, int S>
static int launch_dw(const AchDwConv& p, cudaStream_t st) {
    // tile = (TH x TW) outputs x CPB channels with TW a multiple of 4 that divides the (rounded) row evenly, so that
    // e.g. 40-wide planes are one 40-wide tile instead of a 32-wide tile plus a mostly empty one
    const int Wr = (p.Wo + 3) & ~3;
    const int n_tx = cdiv(Wr, 64);
    const int TW = (cdiv(Wr, n_tx) + 3) & ~3;
    const int strips = TW / DW_NX;
    int TH = min(p.Ho, max(1, 256 / strips));
    TH = cdiv(p.Ho, cdiv(p.Ho, TH));           // balance the rows over the tiles
    int CPB = max(1, 256 / (TH * strips));     // >= 1 strip per thread per pass
    CPB = min(CPB, p.C);
    const int IH = (TH - 1) * S + KS, IW = (TW - 1) * S + KS;
    const int IWp = ((IW + 3) & ~3) + 4;
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
