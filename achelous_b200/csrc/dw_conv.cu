// Depthwise k x k convolution (k in {3,5,7,9}; stride 1/2; pad k/2), HBM-bound stencil.
// Two kernels: dw_rows_kernel (stride 1, the hot one: register sliding window, see below) and the tiled
// dw_conv_kernel (stride 2, MobileViT MV2 blocks).
//
// Tiled kernel: one CTA stages the halo tile of CPB channels of one frame in shared memory (warp-per-row coalesced
// loads, optional fused pre-add of a second tensor - the SDTA cascade `conv(sp + spx[i])`), then every
// thread produces a strip of 4 horizontally adjacent outputs: per kernel row it pulls the
// (3*S + KS)-wide input segment with conflict-free 128-bit shared loads into registers and reuses it
// for all 4 outputs and KS taps (0.3-0.5 shared loads per FMA instead of 2), with the KS*KS taps of the
// channel held in registers.  Epilogue: folded BN/bias, activation, optional broadcast post-add
// (positional encoding), 128-bit stores when the row pitch allows.
#include "common.cuh"
#include "tma_common.cuh"
#include <cstdlib>

namespace ach {

constexpr int DW_NX = 4;  // outputs per thread along x

// use_tma: the halo tile of the CPB channels arrives as ONE tensor-map box (zero-filled outside the image = the conv's zero padding);
// the box has to start at a 16-byte aligned column, XO = (-pad) mod 4 columns left of the tile, which only shifts the register
// segment (ncu on the warp-per-row staging loop at the MobileViT stride-2 layers: 9 long-scoreboard stalls per issue, 1.2 TB/s).
template <int KS, int S>
__global__ void __launch_bounds__(256) dw_conv_kernel(const AchDwConv p, int TH, int TW, int CPB, int tiles_x, const __grid_constant__ CUtensorMap tmx,
                                                      int use_tma) {
    extern __shared__ __align__(128) float smem[];
    constexpr int XO_T = (4 - ((KS / 2) & 3)) & 3;
    constexpr int SEG = (DW_NX - 1) * S + KS + XO_T;   // input values per kernel row per thread (incl. the TMA column shift)
    constexpr int SEG4 = (SEG + 3) / 4;            // as float4 loads
    const int xo = use_tma ? XO_T : 0;
    const int IH = (TH - 1) * S + KS;
    const int IW = (TW - 1) * S + KS;
    const int IWp = ((IW + XO_T + 3) & ~3) + 4;    // 16B-aligned rows + slack for the last segment's over-read
    __shared__ __align__(8) uint64_t mbar;
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

    if (use_tma && threadIdx.x == 0) {
        tma_mbar_init(tma_smem_u32(&mbar), 1);
        tma_mbar_expect_tx(tma_smem_u32(&mbar), (uint32_t)(CPB * IH * IWp) * 4u);   // always the full box: missing channels are zero-filled
        tma_load_4d(tma_smem_u32(tile), &tmx, ix0 - XO_T, iy0, c_base, b, tma_smem_u32(&mbar));
    }
    for (int i = threadIdx.x; i < nch * KS * KS; i += 256) wsm[i] = p.w[(long long)c_base * KS * KS + i];

    // ---- stage the halo tile: one warp per (channel, row)
    const float* xb = p.x + (long long)b * p.x_bs;
    const float* ab = p.xadd ? p.xadd + (long long)b * p.xadd_bs : nullptr;
    for (int r = warp; !use_tma && r < nch * IH; r += 8) {
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
    if (use_tma) tma_mbar_wait(tma_smem_u32(&mbar), 0);

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
                for (int j = 0; j < DW_NX; ++j) acc[j] = fmaf(seg[j * S + kx + xo], wk[ky * KS + kx], acc[j]);
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

// ---- stride-1 path: register sliding window, no shared memory, no barrier.
// Thread = (frame, channel, block of R output rows, strip of 4 output columns); consecutive threads take consecutive
// strips, so every row access of a warp is a run of contiguous 16-byte loads.  The thread walks the R + KS - 1 input
// rows of its block once; each row segment (4 + 2*pad values: the strip plus its halo, the halo coming from L1 where
// the neighbouring thread's strip already is) feeds the up to KS output rows it overlaps while it is in registers.
// Everything is unrolled at compile time: KS*KS taps and R*4 accumulators live in registers, the instruction stream
// is KS*KS FMAs per output plus ~3 loads per 4*KS FMAs.  The tiled kernel above spends 5-10x that on staging
// arithmetic (ncu: 160 instructions per output for a 3x3).  Summation order per output is (ky, kx) like the tiled kernel.
template <int KS, int R, bool AL>
__global__ void __launch_bounds__(128) dw_rows_kernel(const AchDwConv p, int strips, int nrb, long long total) {
    constexpr int pad = KS / 2;
    constexpr int SEG = 4 + 2 * pad;
    const long long idx = (long long)blockIdx.x * 128 + threadIdx.x;
    if (idx >= total) return;
    const int s = (int)(idx % strips);
    long long t = idx / strips;
    const int rb = (int)(t % nrb);
    t /= nrb;
    const int c = (int)(t % p.C);
    const int b = (int)(t / p.C);
    const int H = p.H, W = p.W;
    const int x0 = 4 * s, r0 = rb * R;
    const long long plane = (long long)H * W;
    const float* __restrict__ xp = p.x + (long long)b * p.x_bs + (long long)c * plane;
    const float* __restrict__ ap = p.xadd ? p.xadd + (long long)b * p.xadd_bs + (long long)c * plane : nullptr;

    float wk[KS * KS];
#pragma unroll
    for (int i = 0; i < KS * KS; ++i) wk[i] = __ldg(p.w + (long long)c * KS * KS + i);

    float acc[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;

    const bool has_l = s > 0, has_r = s < strips - 1;
    auto load_seg = [&](const float* __restrict__ row, float (&seg)[SEG]) {
        if constexpr (AL) {
            const float4 cur = __ldg(reinterpret_cast<const float4*>(row + x0));
            seg[pad + 0] = cur.x; seg[pad + 1] = cur.y; seg[pad + 2] = cur.z; seg[pad + 3] = cur.w;
            if constexpr (pad == 1) {
                seg[0] = has_l ? __ldg(row + x0 - 1) : 0.f;
                seg[5] = has_r ? __ldg(row + x0 + 4) : 0.f;
            } else if constexpr (pad == 2) {
                const float2 l = has_l ? __ldg(reinterpret_cast<const float2*>(row + x0 - 2)) : make_float2(0.f, 0.f);
                const float2 r = has_r ? __ldg(reinterpret_cast<const float2*>(row + x0 + 4)) : make_float2(0.f, 0.f);
                seg[0] = l.x; seg[1] = l.y; seg[6] = r.x; seg[7] = r.y;
            } else {
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 l = has_l ? __ldg(reinterpret_cast<const float4*>(row + x0 - 4)) : z;
                const float4 r = has_r ? __ldg(reinterpret_cast<const float4*>(row + x0 + 4)) : z;
                const float lv[4] = {l.x, l.y, l.z, l.w}, rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int i = 0; i < pad; ++i) {
                    seg[i] = lv[4 - pad + i];
                    seg[pad + 4 + i] = rv[i];
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < SEG; ++i) {
                const int xx = x0 - pad + i;
                seg[i] = (xx >= 0 && xx < W) ? __ldg(row + xx) : 0.f;
            }
        }
    };

    // small windows: every row segment of the block is requested before the first FMA (ncu on the k = 3 layers at 80 x 80: 7.5
    // long-scoreboard stalls per issued instruction with the loads interleaved row by row)
    constexpr bool PRELOAD = (R + 2 * pad) * SEG <= 64;
    float pre[PRELOAD ? R + 2 * pad : 1][SEG];
    if constexpr (PRELOAD) {
#pragma unroll
        for (int i = 0; i < R + 2 * pad; ++i) {
            const int y = r0 - pad + i;
#pragma unroll
            for (int q = 0; q < SEG; ++q) pre[i][q] = 0.f;
            if (y >= 0 && y < H) load_seg(xp + (long long)y * W, pre[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < R + 2 * pad; ++i) {
        const int y = r0 - pad + i;
        float seg[SEG];
#pragma unroll
        for (int q = 0; q < SEG; ++q) seg[q] = PRELOAD ? pre[PRELOAD ? i : 0][q] : 0.f;
        if (y >= 0 && y < H) {
            if constexpr (!PRELOAD) load_seg(xp + (long long)y * W, seg);
            if (ap) {
                float sa[SEG];
                load_seg(ap + (long long)y * W, sa);
#pragma unroll
                for (int q = 0; q < SEG; ++q) seg[q] += sa[q];
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int ky = i - r;   // input row y = (r0 + r) - pad + ky
            if (ky >= 0 && ky < KS) {
#pragma unroll
                for (int kx = 0; kx < KS; ++kx)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[r][j] = fmaf(seg[j + kx], wk[ky * KS + kx], acc[r][j]);
            }
        }
    }

    const float sc = p.scale ? p.scale[c] : 1.f;
    const float bi = p.bias ? p.bias[c] : 0.f;
    float* __restrict__ op = p.out + (long long)b * p.out_bs + (long long)c * plane;
    const float* __restrict__ pp = p.post ? p.post + (long long)c * plane : nullptr;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int y = r0 + r;
        if (y >= H) break;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = apply_act(fmaf(sc, acc[r][j], bi), p.act);
        const long long po = (long long)y * W + x0;
        if constexpr (AL) {
            if (pp) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(pp + po));
                v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w;
            }
            *reinterpret_cast<float4*>(op + po) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x0 + j < W) op[po + j] = pp ? v[j] + pp[po + j] : v[j];
        }
    }
}

template <int KS, int R>
static int launch_rows(const AchDwConv& p, cudaStream_t st) {
    const bool al = (p.W % 4 == 0) && aligned16(p.x) && aligned16(p.xadd) && aligned16(p.out) && aligned16(p.post) &&
                    p.x_bs % 4 == 0 && p.xadd_bs % 4 == 0 && p.out_bs % 4 == 0;
    const int strips = cdiv(p.W, 4), nrb = cdiv(p.H, R);
    const long long total = (long long)p.B * p.C * nrb * strips;
    const unsigned grid = (unsigned)cdiv(total, 128);
    if (al) dw_rows_kernel<KS, R, true><<<grid, 128, 0, st>>>(p, strips, nrb, total);
    else dw_rows_kernel<KS, R, false><<<grid, 128, 0, st>>>(p, strips, nrb, total);
    return check_launch("ach_dw_conv");
}

template <int KS>
static int launch_rows_k(const AchDwConv& p, cudaStream_t st) {
    // rows per thread: 8 on tall planes (halo re-reads (R + KS - 1) / R stay small), otherwise a divisor of H
    if (p.H >= 40 && KS <= 5) return launch_rows<KS, 8>(p, st);
    // small planes (10 x 10): few threads each running thousands of straight-line instructions once (64 KB+ of SASS at k = 9,
    // R = 5) are instruction-fetch / latency bound - 2 rows per thread gives 2.5x the threads and a third of the code; the extra halo
    // re-reads hit L1 / L2 (the planes are tiny)
    if (p.H <= 10) return launch_rows<KS, 2>(p, st);   // measured: 10 x 10 k = 9 0.036 -> 0.022 ms; 20 x 20 k = 7 slightly slower (stays at 5 rows)
    if (p.H % 5 == 0) return launch_rows<KS, 5>(p, st);
    return launch_rows<KS, 4>(p, st);
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
    constexpr int XO_T = (4 - ((KS / 2) & 3)) & 3;
    const int IH = (TH - 1) * S + KS, IW = (TW - 1) * S + KS;
    const int IWp = ((IW + XO_T + 3) & ~3) + 4;
    while (CPB > 1 && (size_t)CPB * (IH * IWp + KS * KS) * 4 > 96 * 1024) --CPB;
    const size_t smem = (size_t)CPB * (IH * IWp + KS * KS) * sizeof(float);
    alignas(64) CUtensorMap tmx;
    memset(&tmx, 0, sizeof(tmx));
    const int use_tma = (!p.xadd && p.W % 4 == 0 && IWp <= 256 && IH <= 256 && CPB <= 256 &&
                         tma_map_planes(&tmx, p.x, p.W, p.H, p.C, p.B, p.x_bs, IWp, IH, CPB)) ? 1 : 0;
    const int tiles_x = cdiv(p.Wo, TW), tiles_y = cdiv(p.Ho, TH);
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(dw_conv_kernel<KS, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    }
    dim3 grid(tiles_x * tiles_y, cdiv(p.C, CPB), p.B);
    dw_conv_kernel<KS, S><<<grid, 256, smem, st>>>(p, TH, TW, CPB, tiles_x, tmx, use_tma);
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
    if (p.stride == 1) {
        switch (p.k) {
            case 3: return launch_rows_k<3>(p, st);
            case 5: return launch_rows_k<5>(p, st);
            case 7: return launch_rows_k<7>(p, st);
            case 9: return launch_rows_k<9>(p, st);
            default: break;
        }
    }
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
