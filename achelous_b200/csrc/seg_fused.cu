// Fused segmentation-decoder stages of the Ghost-Dual-FPN (ghostdualfpn.py:175-197).
//
// The decoder is the byte-dominant part of the network (85 % of block-level traffic):
// each stage is  1x1 conv+BN+ReLU -> bilinear x2 (align_corners) -> GhostModule(1x1+BN+ReLU | dw3x3+BN+ReLU).
// Bilinear interpolation and the Ghost "primary" 1x1 conv + BN affine are both linear and act on
// different axes (space vs channels), so they commute:  primary(up(t)) == up(primary_linear(t)) + b.
// The host therefore runs the primary conv at LOW resolution (4x fewer MACs, half the channels) and
// these kernels do everything that happens at HIGH resolution in one pass over shared memory:
//
//   ach_up_ghost       v (B,Ci,h,w) -> out (B,Ci+Cn,2h,2w):  x1 = relu(up(v) + b1);  x2 = relu(s2*dw3x3(x1) + b2)
//   ach_up_ghost_head  v (B,16,h,w) -> logits (B,K,2h,2w): the same stage, immediately followed by the head
//                      GhostModule (1x1 32->INIT +BN+ReLU | dw3x3 +BN+ReLU, slice to K) WITHOUT ever writing
//                      the 32-channel full-resolution tensor (13 MB/frame in the reference) to HBM.
//
// Tiles: the halo of each stage is recomputed (cheap ALU) instead of exchanged through memory.  All
// per-channel weights of ach_up_ghost_head travel as kernel parameters (constant bank) so the inner FMAs
// take them as immediate constant operands - no load instructions for weights at all.
#include "common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace ach {

// ATen upsample_bilinear2d (align_corners=True) source coordinate for destination index d
__device__ __forceinline__ void bilin_src(int d, float scale, int in_size, int& i0, int& i1, float& l1) {
    const float f = scale * (float)d;
    i0 = (int)f;
    i1 = i0 + (i0 < in_size - 1);
    l1 = f - (float)i0;
}

// ------------------------------------------------------------------------------------------------
// ach_up_ghost: per channel, 32x32 output tile.
constexpr int UG_T = 32;                 // output tile edge
constexpr int UG_X1 = UG_T + 2;          // x1 tile edge (halo 1 for the dw3x3)
constexpr int UG_V = UG_X1 / 2 + 3;      // low-res tile edge (ceil(34 * 0.5) + interpolation neighbour + slack)
constexpr int UG_CPB = 4;                // channels per CTA

__global__ void __launch_bounds__(256) up_ghost_kernel(const AchUpGhost p) {
    __shared__ float vs[UG_V][UG_V + 1];
    __shared__ float x1s[UG_X1][UG_X1 + 1];
    const int H = 2 * p.h, W = 2 * p.w;
    const int tiles_x = (W + UG_T - 1) / UG_T;
    const int ty0 = (blockIdx.x / tiles_x) * UG_T, tx0 = (blockIdx.x % tiles_x) * UG_T;
    const int b = blockIdx.z;
    const float sy = (float)(p.h - 1) / (float)(H - 1), sx = (float)(p.w - 1) / (float)(W - 1);
    // low-res origin of this tile: source row/col of the first x1 row/col (clamped to the image)
    const int vy0 = (int)(sy * (float)max(ty0 - 1, 0)), vx0 = (int)(sx * (float)max(tx0 - 1, 0));
    const long long plane_lo = (long long)p.h * p.w, plane_hi = (long long)H * W;

    for (int cc = 0; cc < UG_CPB; ++cc) {
        const int c = blockIdx.y * UG_CPB + cc;
        if (c >= p.Ci) break;
        const float* vp = p.v + (long long)b * p.v_bs + (long long)c * plane_lo;
        __syncthreads();
        for (int i = threadIdx.x; i < UG_V * UG_V; i += 256) {
            const int yy = i / UG_V, xx = i - yy * UG_V;
            const int gy = min(vy0 + yy, p.h - 1), gx = min(vx0 + xx, p.w - 1);
            vs[yy][xx] = vp[gy * p.w + gx];
        }
        __syncthreads();
        const float b1 = p.b1[c];
        float* o1 = p.out + (long long)b * p.out_bs + (long long)c * plane_hi;
        for (int i = threadIdx.x; i < UG_X1 * UG_X1; i += 256) {
            const int yy = i / UG_X1, xx = i - yy * UG_X1;
            const int gy = ty0 - 1 + yy, gx = tx0 - 1 + xx;
            float val = 0.f;  // zero padding of the dw3x3 outside the image
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                int y0, y1, x0, x1;
                float ly, lx;
                bilin_src(gy, sy, p.h, y0, y1, ly);
                bilin_src(gx, sx, p.w, x0, x1, lx);
                const float hy = 1.f - ly, hx = 1.f - lx;
                y0 -= vy0; y1 -= vy0; x0 -= vx0; x1 -= vx0;
                val = hy * (hx * vs[y0][x0] + lx * vs[y0][x1]) + ly * (hx * vs[y1][x0] + lx * vs[y1][x1]);
                val = fmaxf(val + b1, 0.f);
                if (yy >= 1 && yy <= UG_T && xx >= 1 && xx <= UG_T) o1[(long long)gy * W + gx] = val;
            }
            x1s[yy][xx] = val;
        }
        __syncthreads();
        if (c < p.Cn) {
            float wk[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) wk[t] = p.w2[c * 9 + t];
            const float s2 = p.s2[c], b2 = p.b2[c];
            float* o2 = p.out + (long long)b * p.out_bs + (long long)(p.Ci + c) * plane_hi;
            for (int i = threadIdx.x; i < UG_T * UG_T; i += 256) {
                const int yy = i / UG_T, xx = i - yy * UG_T;
                const int gy = ty0 + yy, gx = tx0 + xx;
                if (gy >= H || gx >= W) continue;
                float acc = 0.f;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) acc = fmaf(x1s[yy + ky][xx + kx], wk[ky * 3 + kx], acc);
                o2[(long long)gy * W + gx] = fmaxf(fmaf(s2, acc, b2), 0.f);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ach_up_ghost_head: last decoder stage + head GhostModule.  All per-channel weights travel as kernel parameters.
constexpr int UH_C = 16;                 // channels of v / x1 / x2

template <int INIT, int KOUT>
struct UpGhostHeadParams {
    // stage ghost module
    float b1[UH_C];
    float w2[UH_C * 9];
    float s2[UH_C];
    float b2[UH_C];
    // head ghost module: primary (BN scale folded into w3), cheap dw
    float w3[2 * UH_C * INIT];  // [c][i]
    float b3[INIT];
    float w4[(KOUT - INIT) * 9];
    float s4[KOUT - INIT];
    float b4[KOUT - INIT];
};

// ------------------------------------------------------------------------------------------------
// ach_up_ghost_head ("v3").  ARGMAX: instead of the K logit planes the kernel writes the per-pixel class index (first maximum,
// like torch.argmax over the very same fp32 logits) as one byte - achelous.py:283-297 only ever uses the argmax of these maps;
// classes whose bit in keep_mask is clear are mapped to 0 (`output_seg[(output_seg != 0) & (output_seg != 8)] = 0`).
// History (profiles/r2_head_kernel.md): the round-1 kernel kept the 16-channel x1 tile in shared memory (103 KB, one scalar LDS
// per value and consumer; ncu: issue 62 %, l1tex 64 %, 24 % warps active; 0.411 ms).  A version with x1 built chunk-wise in shared
// memory by separable passes cut the instruction count by a third and got SLOWER (0.47 ms) - ten CTA barriers per tile at 2-3
// resident CTAs leave it latency bound (issue 35 %, barrier stall 1.6 per issue).  This version (0.247 ms) has NO barrier and no
// shared-memory round trip in its main loop:
//   * bilinear x2 with align_corners has a STATIC index pattern: out[2m] mixes v[m-1], v[m]; out[2m+1] mixes v[m], v[m+1]
//     (2m*(w-1)/(2w-1) = m - m/(2w-1)); only the weights move with the position.  So the 4 x 6 x1 window a thread needs for its
//     2 x 4 pixels is the separable interpolation of a 4 x 5 low-resolution patch, and is rebuilt in REGISTERS from the patch
//     (12 8-byte shared loads) with per-thread weights computed once - image borders, the clamp at the last row / column and the
//     dw conv's zero padding are folded into those weights (v + b1 is interpolated, so zero weights give relu(0) = 0);
//   * the low-resolution tile of all 16 channels (24 x 24 x 16 floats, 36 KB) arrives by ONE 4-D TMA box copy per CTA
//     (cp.async.bulk.tensor, out-of-image rows / columns zero-filled by the copy engine: no index arithmetic, no LDG / STS);
//   * 32 x 40 output tiles, 2 x 4 pixels per thread, packed FFMA2 on horizontally adjacent pixels with the weights as uniform
//     scalars from the constant bank, STG.128 / one STG.32 of four class bytes (ARGMAX) on the way out.
#ifndef H3_UNROLL_N
#define H3_UNROLL_N 4
#endif
constexpr int H3_UNROLL = H3_UNROLL_N;   // channel-loop unroll of the head kernel (16 = full: 92 KB of SASS, ncu: 1.75 no-instruction stalls per issue)
constexpr int H3_TW = 32, H3_TH = 40, H3_T = 192;
constexpr int H3_PH = H3_TH + 2, H3_PP = 40;          // p tile: rows (global ty0 - 1 + pr), pitch (36 columns used: tx0 - 1 + pc)
constexpr int H3_VR = 24, H3_VP = 24;                 // low-resolution tile: rows (ty0/2 - 2 + r), columns (tx0/2 - 4 + c)
constexpr int H3_VC = 2;                              // first used tile column: the box starts 2 columns early because the innermost
                                                      // TMA start coordinate must be a multiple of 16 bytes (tools/probe/tma_probe.cu)
constexpr int H3_VELEMS = UH_C * H3_VR * H3_VP;
static_assert(H3_TW / 2 + 5 + H3_VC <= H3_VP + 1 && H3_TH / 2 + 4 <= H3_VR && (H3_TW / 2) % 4 == 0, "low-resolution tile too small for the output tile");

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

// weights (wa, wb) on the static source pair (a, a + 1) for destination index g of an axis of out_size = 2 * in_size pixels;
// both 0 outside the image.  ATen's own source index / lambda (bilin_src) decides; the static pair is only the addressing.
__device__ __forceinline__ void static_lerp_weights(int g, int out_size, int in_size, float scale, int a, float& wa, float& wb) {
    wa = 0.f;
    wb = 0.f;
    if (g < 0 || g >= out_size) return;
    int i0, i1;
    float l;
    bilin_src(g, scale, in_size, i0, i1, l);
    if (i1 == i0) l = 0.f;                       // clamped at the last row / column: the value is v[i0]
    const int d = i0 - a;
    if (d == 0) { wa = 1.f - l; wb = l; }
    else if (d == 1) { wb = 1.f; }                // g == 0: l == 0 exactly, the value is v[0] = v[a + 1]
    else if (d == -1) { wb = 0.f; wa = 1.f; }     // scale * (out_size - 1) rounded just below in_size - 1: l = 1 - O(1e-6), value v[a]
    else __trap();                                // cannot happen for out_size == 2 * in_size
}

template <int INIT, int KOUT, bool ARGMAX>
__global__ void __maxnreg__(112) up_ghost_head3_kernel(const __grid_constant__ CUtensorMap tmv, const float* __restrict__ v, long long v_bs,
                                                        float* __restrict__ out, long long out_bs, int h, int w,
                                                        const __grid_constant__ UpGhostHeadParams<INIT, KOUT> P,
                                                        unsigned char* __restrict__ mask, long long mask_bs, unsigned keep_mask, int vec_ok,
                                                        int use_tma) {
    extern __shared__ __align__(128) float smem[];
    float* vs = smem;                                    // [16][H3_VR][H3_VP]
    float* ps = smem;                                    // [INIT][H3_PH][H3_PP]   (after the channel loop)
    static_assert(INIT * H3_PH * H3_PP <= H3_VELEMS && KOUT - INIT <= INIT, "p tile must fit in the low-resolution tile");
    __shared__ __align__(8) unsigned long long mbar;

    const int H = 2 * h, W = 2 * w, tid = threadIdx.x;
    const int tiles_x = (W + H3_TW - 1) / H3_TW;
    const int ty0 = (blockIdx.x / tiles_x) * H3_TH, tx0 = (blockIdx.x % tiles_x) * H3_TW;
    const int b = blockIdx.y;
    const int R0 = ty0 / 2 - 2, C0 = tx0 / 2 - 2 - H3_VC;
    const long long plane_hi = (long long)H * W;

    if (use_tma) {
        if (tid == 0) {
            tma_mbar_init(tma_smem_u32(&mbar), 1);
            tma_mbar_expect_tx(tma_smem_u32(&mbar), (uint32_t)H3_VELEMS * 4u);
            tma_load_4d(tma_smem_u32(vs), &tmv, C0, R0, 0, b, tma_smem_u32(&mbar));
        }
    } else {   // views a tensor map cannot describe (row pitch not a multiple of 16 bytes): the same tile by plain loads
        const float* vb = v + (long long)b * v_bs;
        for (int i = tid; i < H3_VELEMS; i += H3_T) {
            const int c = i / (H3_VR * H3_VP), r = (i / H3_VP) % H3_VR, x = i % H3_VP;
            const int gy = R0 + r, gx = C0 + x;
            vs[i] = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? __ldg(vb + ((long long)c * h + gy) * w + gx) : 0.f;
        }
    }

    // ---- per-thread interpolation weights (the tile is in flight meanwhile).  thread -> p rows 2*tr, 2*tr+1, p columns 4*tc .. 4*tc+3;
    // x1 window rows i = 0..3 (global ty0 - 2 + 2*tr + i), columns j = 0..5 (global tx0 - 2 + 4*tc + j)
    const int tr = tid / 9, tc = tid - tr * 9;
    const bool p_thread = tid < (H3_PH / 2) * 9;
    const float sy = (float)(h - 1) / (float)(H - 1), sx = (float)(w - 1) / (float)(W - 1);
    float2 cwa[3], cwb[3];          // column weights, pairs (j, j + 1) for j = 0, 2, 4
    float rwa[4], rwb[4];           // row weights
    {
        const int m0 = tx0 / 2 - 1 + 2 * tc, n0 = ty0 / 2 - 1 + tr;
        const int gx0 = tx0 - 2 + 4 * tc, gy0 = ty0 - 2 + 2 * tr;
#pragma unroll
        for (int q = 0; q < 3; ++q) {   // static source columns of j = 2q: (m0 + q - 1, m0 + q); of j = 2q + 1: (m0 + q, m0 + q + 1)
            static_lerp_weights(gx0 + 2 * q, W, w, sx, m0 + q - 1, cwa[q].x, cwb[q].x);
            static_lerp_weights(gx0 + 2 * q + 1, W, w, sx, m0 + q, cwa[q].y, cwb[q].y);
        }
        static_lerp_weights(gy0 + 0, H, h, sy, n0 - 1, rwa[0], rwb[0]);
        static_lerp_weights(gy0 + 1, H, h, sy, n0, rwa[1], rwb[1]);
        static_lerp_weights(gy0 + 2, H, h, sy, n0, rwa[2], rwb[2]);
        static_lerp_weights(gy0 + 3, H, h, sy, n0 + 1, rwa[3], rwb[3]);
    }
    float2 acc[2][2][INIT];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int i = 0; i < INIT; ++i) acc[r][q][i] = f2(P.b3[i], P.b3[i]);

    if (use_tma) {
        __syncthreads();                               // mbarrier initialised before anybody polls it
        tma_mbar_wait(tma_smem_u32(&mbar), 0);
    } else {
        __syncthreads();
    }

    if (p_thread) {
        const float* vt = vs + tr * H3_VP + H3_VC + 2 * tc;    // patch rows tr .. tr+3, columns H3_VC + 2*tc .. + 4
#pragma unroll H3_UNROLL
        for (int c = 0; c < UH_C; ++c) {
            const float* vp = vt + c * (H3_VR * H3_VP);
            const float b1 = P.b1[c];
            // horizontal pass on (v + b1): t[r][q] = pair of x1 columns (2q, 2q + 1) of patch row r
            float2 t[4][3];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float2 p01 = *reinterpret_cast<const float2*>(vp + r * H3_VP);
                const float2 p23 = *reinterpret_cast<const float2*>(vp + r * H3_VP + 2);
                const float2 p45 = *reinterpret_cast<const float2*>(vp + r * H3_VP + 4);
                const float v0 = p01.x + b1, v1 = p01.y + b1, v2 = p23.x + b1, v3 = p23.y + b1, v4 = p45.x + b1;
                t[r][0] = __ffma2_rn(cwb[0], f2(v1, v2), __fmul2_rn(cwa[0], f2(v0, v1)));
                t[r][1] = __ffma2_rn(cwb[1], f2(v2, v3), __fmul2_rn(cwa[1], f2(v1, v2)));
                t[r][2] = __ffma2_rn(cwb[2], f2(v3, v4), __fmul2_rn(cwa[2], f2(v2, v3)));
            }
            // vertical pass + relu: window rows 0..3 use patch rows (0,1), (1,2), (1,2), (2,3)
            float2 x1w[4][3];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ra = (i + 1) / 2;     // 0, 1, 1, 2
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float2 val = __ffma2_rn(f2(rwb[i], rwb[i]), t[ra + 1][q], __fmul2_rn(f2(rwa[i], rwa[i]), t[ra][q]));
                    x1w[i][q] = f2(fmaxf(val.x, 0.f), fmaxf(val.y, 0.f));
                }
            }
            // Ghost cheap op (dw 3x3 + BN + ReLU) and the head's primary 1x1 on the 2 x 4 pixels
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float2 d = f2(0.f, 0.f);
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const float2 e0 = x1w[r + ky][q], e2 = x1w[r + ky][q + 1];
                        const float w0 = P.w2[c * 9 + ky * 3], w1 = P.w2[c * 9 + ky * 3 + 1], w2_ = P.w2[c * 9 + ky * 3 + 2];
                        d = __ffma2_rn(e0, f2(w0, w0), d);
                        d = __ffma2_rn(f2(e0.y, e2.x), f2(w1, w1), d);
                        d = __ffma2_rn(e2, f2(w2_, w2_), d);
                    }
                    d = __ffma2_rn(f2(P.s2[c], P.s2[c]), d, f2(P.b2[c], P.b2[c]));
                    const float2 x2 = f2(fmaxf(d.x, 0.f), fmaxf(d.y, 0.f));
                    const float2 x1c = f2(x1w[r + 1][q].y, x1w[r + 1][q + 1].x);
#pragma unroll
                    for (int i = 0; i < INIT; ++i) {
                        const float wa = P.w3[c * INIT + i], wb = P.w3[(UH_C + c) * INIT + i];
                        acc[r][q][i] = __ffma2_rn(x1c, f2(wa, wa), acc[r][q][i]);
                        acc[r][q][i] = __ffma2_rn(x2, f2(wb, wb), acc[r][q][i]);
                    }
                }
        }
    }
    __syncthreads();                                   // every thread is done with the low-resolution tile: p may overwrite it

    // ---- p = relu(acc), 0 outside the image (= zero padding of the head's dw conv)
    if (p_thread) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int pr = 2 * tr + r, gy = ty0 - 1 + pr;
            const bool rin = gy >= 0 && gy < H;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int gx = tx0 - 1 + 4 * tc + 2 * q;
                const bool in0 = rin && gx >= 0 && gx < W, in1 = rin && gx + 1 >= 0 && gx + 1 < W;
#pragma unroll
                for (int i = 0; i < INIT; ++i)
                    *reinterpret_cast<float2*>(ps + (i * H3_PH + pr) * H3_PP + 4 * tc + 2 * q) =
                        f2(in0 ? fmaxf(acc[r][q][i].x, 0.f) : 0.f, in1 ? fmaxf(acc[r][q][i].y, 0.f) : 0.f);
            }
        }
    }
    __syncthreads();

    // ---- outputs: thread -> rows 2*orow, 2*orow + 1, columns 4*og .. 4*og + 3; channels [0, INIT) = p, [INIT, KOUT) = relu(s4*dw3x3(p)+b4)
    if (tid < (H3_TH / 2) * (H3_TW / 4)) {
        const int og = tid & 7, orow = tid >> 3;
        const int gx = tx0 + 4 * og;
        float best[2][4];
        int arg[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int x = 0; x < 4; ++x) { best[r][x] = -INFINITY; arg[r][x] = 0; }
        // one channel's four values of output row 2*orow + r: stored at once, or folded into the running first-maximum
        auto emit = [&](int r, int k, float a0, float a1, float a2, float a3) {
            const int gy = ty0 + 2 * orow + r;
            if (ARGMAX) {
                const float a[4] = {a0, a1, a2, a3};
#pragma unroll
                for (int x = 0; x < 4; ++x)
                    if (a[x] > best[r][x] || (a[x] == best[r][x] && k < arg[r][x])) { best[r][x] = a[x]; arg[r][x] = k; }
            } else if (gy < H && gx < W) {
                float* ob = out + (long long)b * out_bs + (long long)k * plane_hi + (long long)gy * W + gx;
                if (vec_ok) {
                    *reinterpret_cast<float4*>(ob) = make_float4(a0, a1, a2, a3);
                } else {
                    ob[0] = a0;
                    if (gx + 1 < W) ob[1] = a1;
                    if (gx + 2 < W) ob[2] = a2;
                    if (gx + 3 < W) ob[3] = a3;
                }
            }
        };
#pragma unroll
        for (int k = 0; k < INIT; ++k) {
            const float* pb = ps + (k * H3_PH + 2 * orow) * H3_PP + 4 * og;
            float win[4][6];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (k < KOUT - INIT || r == 1 || r == 2) {
                    const float4 a4 = *reinterpret_cast<const float4*>(pb + r * H3_PP);
                    const float2 b2 = *reinterpret_cast<const float2*>(pb + r * H3_PP + 4);
                    win[r][0] = a4.x; win[r][1] = a4.y; win[r][2] = a4.z; win[r][3] = a4.w; win[r][4] = b2.x; win[r][5] = b2.y;
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) emit(r, k, win[r + 1][1], win[r + 1][2], win[r + 1][3], win[r + 1][4]);
            if (k < KOUT - INIT) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    float2 d[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        d[q] = f2(0.f, 0.f);
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const float wk = P.w4[k * 9 + ky * 3 + kx];
                                d[q] = __ffma2_rn(f2(win[r + ky][2 * q + kx], win[r + ky][2 * q + kx + 1]), f2(wk, wk), d[q]);
                            }
                        d[q] = __ffma2_rn(f2(P.s4[k], P.s4[k]), d[q], f2(P.b4[k], P.b4[k]));
                    }
                    emit(r, INIT + k, fmaxf(d[0].x, 0.f), fmaxf(d[0].y, 0.f), fmaxf(d[1].x, 0.f), fmaxf(d[1].y, 0.f));
                }
            }
        }
        if (ARGMAX) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int gy = ty0 + 2 * orow + r;
                if (gy >= H || gx >= W) continue;
                unsigned packed = 0;
#pragma unroll
                for (int x = 0; x < 4; ++x) packed |= (unsigned)(((keep_mask >> arg[r][x]) & 1u) ? arg[r][x] : 0) << (8 * x);
                unsigned char* mb = mask + (long long)b * mask_bs + (long long)gy * W + gx;
                if (vec_ok) {
                    *reinterpret_cast<unsigned*>(mb) = packed;
                } else {
#pragma unroll
                    for (int x = 0; x < 4; ++x)
                        if (gx + x < W) mb[x] = (unsigned char)(packed >> (8 * x));
                }
            }
        }
    }
}

template <int INIT, int KOUT, bool ARGMAX>
static int launch_head3(const AchUpGhostHead& a, cudaStream_t st, unsigned char* mask = nullptr, long long mask_bs = 0,
                        unsigned keep_mask = 0xffffffffu) {
    UpGhostHeadParams<INIT, KOUT> P;
    memcpy(P.b1, a.b1, sizeof(P.b1));
    memcpy(P.w2, a.w2, sizeof(P.w2));
    memcpy(P.s2, a.s2, sizeof(P.s2));
    memcpy(P.b2, a.b2, sizeof(P.b2));
    memcpy(P.w3, a.w3, sizeof(P.w3));
    memcpy(P.b3, a.b3, sizeof(P.b3));
    memcpy(P.w4, a.w4, sizeof(P.w4));
    memcpy(P.s4, a.s4, sizeof(P.s4));
    memcpy(P.b4, a.b4, sizeof(P.b4));
    constexpr size_t smem = (size_t)H3_VELEMS * sizeof(float);
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(up_ghost_head3_kernel<INIT, KOUT, ARGMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    const int H = 2 * a.h, W = 2 * a.w;
    // 16-byte stores need every row start aligned: W % 4 == 0 and aligned bases / batch strides
    int vec_ok = W % 4 == 0;
    if (ARGMAX)
        vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(mask) & 3u) == 0 && mask_bs % 4 == 0;
    else
        vec_ok = vec_ok && aligned16(a.out) && a.out_bs % 4 == 0;
    alignas(64) CUtensorMap tmv;
    memset(&tmv, 0, sizeof(tmv));
    const int use_tma = tma_map_planes(&tmv, a.v, a.w, a.h, UH_C, a.B, a.v_bs, H3_VP, H3_VR, UH_C) ? 1 : 0;
    dim3 grid(cdiv(W, H3_TW) * cdiv(H, H3_TH), a.B);
    up_ghost_head3_kernel<INIT, KOUT, ARGMAX><<<grid, H3_T, smem, st>>>(tmv, a.v, a.v_bs, a.out, a.out_bs, a.h, a.w, P, mask, mask_bs, keep_mask,
                                                                         vec_ok, use_tma);
    return check_launch("ach_up_ghost_head");
}

}  // namespace ach

extern "C" int ach_up_ghost(const AchUpGhost* pp, void* stream) {
    using namespace ach;
    const AchUpGhost& p = *pp;
    ACH_REQUIRE(p.v && p.b1 && p.out, "ach_up_ghost: null arg");
    ACH_REQUIRE(p.Cn == 0 || (p.w2 && p.s2 && p.b2), "ach_up_ghost: null cheap-op weights");
    ACH_REQUIRE(p.B > 0 && p.B <= 65535 && p.Ci > 0 && p.Cn >= 0 && p.Cn <= p.Ci && p.h > 1 && p.w > 1, "ach_up_ghost: bad dims");
    const int H = 2 * p.h, W = 2 * p.w;
    dim3 grid(cdiv(W, UG_T) * cdiv(H, UG_T), cdiv(p.Ci, UG_CPB), p.B);
    up_ghost_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("ach_up_ghost");
}

extern "C" int ach_up_ghost_head_supported(int c_in, int init, int k_out) {
    return c_in == ach::UH_C && ((init == 1 && k_out == 2) || (init == 5 && k_out == 9));
}

extern "C" int ach_up_ghost_head(const AchUpGhostHead* pp, void* stream) {
    using namespace ach;
    const AchUpGhostHead& a = *pp;
    ACH_REQUIRE(a.v && a.out && a.b1 && a.w2 && a.s2 && a.b2 && a.w3 && a.b3 && a.w4 && a.s4 && a.b4, "ach_up_ghost_head: null arg");
    ACH_REQUIRE(a.B > 0 && a.B <= 65535 && a.h > 1 && a.w > 1, "ach_up_ghost_head: bad dims");
    ACH_REQUIRE(ach_up_ghost_head_supported(a.C, a.init, a.K), "ach_up_ghost_head: (C=%d, init=%d, K=%d) not instantiated", a.C, a.init, a.K);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (a.init == 1) return launch_head3<1, 2, false>(a, st);
    return launch_head3<5, 9, false>(a, st);
}

extern "C" int ach_up_ghost_head_argmax(const AchUpGhostHead* pp, unsigned char* mask, long long mask_bs, unsigned keep_mask,
                                        void* stream) {
    using namespace ach;
    const AchUpGhostHead& a = *pp;
    ACH_REQUIRE(a.v && mask && a.b1 && a.w2 && a.s2 && a.b2 && a.w3 && a.b3 && a.w4 && a.s4 && a.b4, "ach_up_ghost_head_argmax: null arg");
    ACH_REQUIRE(a.B > 0 && a.B <= 65535 && a.h > 1 && a.w > 1, "ach_up_ghost_head_argmax: bad dims");
    ACH_REQUIRE(mask_bs >= 4LL * a.h * a.w, "ach_up_ghost_head_argmax: mask batch stride smaller than one map");
    ACH_REQUIRE(ach_up_ghost_head_supported(a.C, a.init, a.K), "ach_up_ghost_head_argmax: (C=%d, init=%d, K=%d) not instantiated", a.C, a.init, a.K);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (a.init == 1) return launch_head3<1, 2, true>(a, st, mask, mask_bs, keep_mask);
    return launch_head3<5, 9, true>(a, st, mask, mask_bs, keep_mask);
}

// ------------------------------------------------------------------------------------------------
// ach_up_ghost_pw2: one decoder stage END TO END at the output resolution of the stage:
//   x1 = relu(up(v) + b1), x2 = relu(s2*dw3x3(x1) + b2)          (GhostModule of stage s, as ach_up_ghost)
//   t  = relu(W1 [x1, x2] + c1)                                   (next stage's Upsample 1x1 conv + BN + ReLU, 32 ch)
//   v' = W2 t                                                      (next stage's Ghost primary conv, BN scale folded, 16 ch)
// so the stage's 2*Ci-channel output and the 32-channel t never touch HBM: the kernel reads Ci channels at (h, w)
// and writes 16 channels at (2h, 2w) - what the next ach_up_ghost_pw2 / ach_up_ghost_head consumes.
// Tile 16 x 32 outputs; x1 on the 18 x 34 halo tile in shared memory; every thread owns a vertical pixel pair,
// keeps t[32][2] in registers and streams the weights from shared memory as float4 broadcasts into FFMA2s.
namespace ach {

constexpr int UP_TH = 16, UP_TW = 32;
constexpr int UP_XH = UP_TH + 2, UP_XW = UP_TW + 2, UP_XP = UP_XW + 1;
constexpr int UP_VH = UP_XH / 2 + 3, UP_VW = UP_XW / 2 + 3;
constexpr int UP_C1 = 32, UP_N2 = 16;

template <int CI>
__global__ void __launch_bounds__(256, 2) up_ghost_pw2_kernel(const AchUpGhostPw2 p) {
    extern __shared__ __align__(16) float smem[];
    float* x1s = smem;                                   // [CI][18][35]
    float* vs = x1s + CI * UP_XH * UP_XP;                // [CI][12][20+1]
    float* w1s = vs + CI * UP_VH * (UP_VW + 1);          // [2*CI][32]
    float* w2s = w1s + 2 * CI * UP_C1;                   // [32][16]
    float* dws = w2s + UP_C1 * UP_N2;                    // [CI][12]: 9 taps, s2, b2, b1
    float* c1s = dws + CI * 12;                          // [32]

    const int h = p.h, w = p.w, H = 2 * h, W = 2 * w;
    const int tiles_x = (W + UP_TW - 1) / UP_TW;
    const int ty0 = (blockIdx.x / tiles_x) * UP_TH, tx0 = (blockIdx.x % tiles_x) * UP_TW;
    const int b = blockIdx.y, tid = threadIdx.x;
    const float sy = (float)(h - 1) / (float)(H - 1), sx = (float)(w - 1) / (float)(W - 1);
    const int vy0 = (int)(sy * (float)max(ty0 - 1, 0)), vx0 = (int)(sx * (float)max(tx0 - 1, 0));
    const long long plane_lo = (long long)h * w, plane_hi = (long long)H * W;
    const float* vb = p.v + (long long)b * p.v_bs;

    for (int i = tid; i < 2 * CI * UP_C1; i += 256) w1s[i] = p.w1t[i];
    for (int i = tid; i < UP_C1 * UP_N2; i += 256) w2s[i] = p.w2t[i];
    for (int i = tid; i < CI * 12; i += 256) {
        const int c = i / 12, k = i - c * 12;
        dws[i] = k < 9 ? p.w2[c * 9 + k] : (k == 9 ? p.s2[c] : (k == 10 ? p.b2[c] : p.b1[c]));
    }
    if (tid < UP_C1) c1s[tid] = p.c1[tid];
    {
        // every load of the low-resolution tile is issued before the first store: ncu showed 35 % of ALL stall samples of the tensor-core
        // kernel on the one STS of the former load -> store loop (15 dependent global-load latencies per CTA, back to back)
        constexpr int V_IT = (CI * UP_VH * UP_VW + 255) / 256;
        float vreg[V_IT];
#pragma unroll
        for (int it = 0; it < V_IT; ++it) {
            const int i = tid + it * 256;
            const int c = i / (UP_VH * UP_VW);
            const int r = i - c * (UP_VH * UP_VW);
            const int yy = r / UP_VW, xx = r - yy * UP_VW;
            const int gy = min(vy0 + yy, h - 1), gx = min(vx0 + xx, w - 1);
            vreg[it] = i < CI * UP_VH * UP_VW ? __ldg(vb + (long long)c * plane_lo + gy * w + gx) : 0.f;
        }
#pragma unroll
        for (int it = 0; it < V_IT; ++it) {
            const int i = tid + it * 256;
            const int c = i / (UP_VH * UP_VW);
            const int r = i - c * (UP_VH * UP_VW);
            const int yy = r / UP_VW, xx = r - yy * UP_VW;
            if (i < CI * UP_VH * UP_VW) vs[(c * UP_VH + yy) * (UP_VW + 1) + xx] = vreg[it];
        }
    }
    __syncthreads();

    // ---- x1 on the halo tile (0 outside the image: dw zero padding)
    for (int i = tid; i < UP_XH * UP_XW; i += 256) {
        const int yy = i / UP_XW, xx = i - yy * UP_XW;
        const int gy = ty0 - 1 + yy, gx = tx0 - 1 + xx;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        int y0 = 0, y1 = 0, x0 = 0, x1 = 0;
        float ly = 0.f, lx = 0.f;
        if (in) {
            bilin_src(gy, sy, h, y0, y1, ly);
            bilin_src(gx, sx, w, x0, x1, lx);
            y0 -= vy0; y1 -= vy0; x0 -= vx0; x1 -= vx0;
        }
        const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll 4
        for (int c = 0; c < CI; ++c) {
            const float* vc = vs + c * UP_VH * (UP_VW + 1);
            float val = hy * (hx * vc[y0 * (UP_VW + 1) + x0] + lx * vc[y0 * (UP_VW + 1) + x1]) +
                        ly * (hx * vc[y1 * (UP_VW + 1) + x0] + lx * vc[y1 * (UP_VW + 1) + x1]);
            x1s[(c * UP_XH + yy) * UP_XP + xx] = in ? fmaxf(val + dws[c * 12 + 11], 0.f) : 0.f;
        }
    }
    __syncthreads();

    // ---- per vertical pixel pair: t = relu(W1 [x1, x2] + c1)
    const int col = tid & 31, rp = tid >> 5;          // output rows 2*rp, 2*rp+1; x1 tile rows 2*rp .. 2*rp+3
    float t[2][UP_C1];
#pragma unroll
    for (int o = 0; o < UP_C1; ++o) t[0][o] = t[1][o] = c1s[o];
#pragma unroll 2
    for (int c = 0; c < CI; ++c) {
        const float* xc = x1s + (c * UP_XH + 2 * rp) * UP_XP + col;
        float win[4][3];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) win[r][k] = xc[r * UP_XP + k];
        const float* dk = dws + c * 12;
        float g1[2], g2[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float d = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) d = fmaf(win[r + ky][kx], dk[ky * 3 + kx], d);
            g1[r] = win[r + 1][1];
            g2[r] = fmaxf(fmaf(dk[9], d, dk[10]), 0.f);
        }
        const float4* wa = reinterpret_cast<const float4*>(w1s + c * UP_C1);
        const float4* wb = reinterpret_cast<const float4*>(w1s + (CI + c) * UP_C1);
#pragma unroll
        for (int i = 0; i < UP_C1 / 4; ++i) {
            const float4 a = wa[i], bq = wb[i];
            fma4_bcast(t[0] + 4 * i, g1[0], a);
            fma4_bcast(t[1] + 4 * i, g1[1], a);
            fma4_bcast(t[0] + 4 * i, g2[0], bq);
            fma4_bcast(t[1] + 4 * i, g2[1], bq);
        }
    }
    // ---- v' = W2 relu(t)
    float vo[2][UP_N2];
#pragma unroll
    for (int o = 0; o < UP_N2; ++o) vo[0][o] = vo[1][o] = 0.f;
#pragma unroll
    for (int k = 0; k < UP_C1; ++k) {
        const float a0 = fmaxf(t[0][k], 0.f), a1 = fmaxf(t[1][k], 0.f);
        const float4* w4 = reinterpret_cast<const float4*>(w2s + k * UP_N2);
#pragma unroll
        for (int i = 0; i < UP_N2 / 4; ++i) {
            const float4 wv = w4[i];
            fma4_bcast(vo[0] + 4 * i, a0, wv);
            fma4_bcast(vo[1] + 4 * i, a1, wv);
        }
    }
    float* ob = p.out + (long long)b * p.out_bs;
    const int gx = tx0 + col;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int gy = ty0 + 2 * rp + r;
        if (gy < H && gx < W) {
#pragma unroll
            for (int o = 0; o < UP_N2; ++o) ob[(long long)o * plane_hi + (long long)gy * W + gx] = vo[r][o];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core version of ach_up_ghost_pw2: the two 1x1 convolutions (2*CI -> 32 -> 16, 1 536 of the ~1 750 FMAs per
// output pixel) run on tcgen05 as 3xTF32 GEMMs; the upsampling and the depthwise 3x3 stay on the CUDA cores.
// Same 16 x 32 output tile and shared-memory staging.  The CTA's 512 pixels are 4 M-tiles of 128: each half of the CTA
// (warps 0-3 / 4-7) owns one M-tile per round (round r = the thread's pixel 2*rp + r) with its own TMEM columns,
// mbarrier and named barrier.  Per round and half:
//   thread = pixel = TMEM lane: [x1 | x2] (2*CI values) split hi/lo -> tcgen05.st (A operand in tensor memory)
//   GEMM 1 (N = 32) -> D;  D + c1, ReLU, split -> tcgen05.st over the dead A columns;  GEMM 2 (N = 32, 16 used) -> D
//   D -> 16 coalesced channel-plane stores.
// Weights: ach_pack_pw_tc tiles of w1t (K = 2*CI, O = 32) and w2t (K = 32, O = 16), resident in shared memory.
// The depthwise 3x3 taps and the per-channel affines travel as KERNEL PARAMETERS (constant bank), like the head kernel's: with them in
// shared memory every FMA of the dw 3x3 had a broadcast LDS of its weight next to the LDS of its x1 value (ncu: LSU pipe 49 %,
// l1tex 68 %, mio-throttle 1.3 per issue) - as immediates of the FMA they cost nothing.
template <int CI>
struct UpDwParams {
    float w[CI * 12];   // per channel: 9 taps, s2, b2, b1
};

template <int CI>
__global__ void __launch_bounds__(256, 2) up_ghost_pw2_tc_kernel(const AchUpGhostPw2 p, const __grid_constant__ UpDwParams<CI> dwp,
                                                                 const float* __restrict__ w1_hi,
                                                                 const float* __restrict__ w1_lo, const float* __restrict__ w2_hi,
                                                                 const float* __restrict__ w2_lo) {
    constexpr int K1 = 2 * CI, NCH1 = K1 / TC_KC, NCH2 = UP_C1 / TC_KC;
    constexpr int AK = K1 > UP_C1 ? K1 : UP_C1;              // columns of one A half (hi or lo)
    constexpr int HALF_COLS = 2 * AK + 32;                   // [A hi | A lo | D]
    constexpr int TMEM_COLS = 2 * HALF_COLS <= 256 ? 256 : 512;
    constexpr int BT = 32 * TC_KC;                           // floats per packed weight tile (NT = 32)
    static_assert(K1 % TC_KC == 0, "2*CI must be a multiple of 16");
    extern __shared__ __align__(128) uint8_t smem_tc_raw[];
    float* b1t = reinterpret_cast<float*>(smem_tc_raw);  // [NCH1][hi | lo][BT]
    float* b2t = b1t + NCH1 * 2 * BT;                    // [NCH2][hi | lo][BT]
    float* x1s = b2t + NCH2 * 2 * BT;                    // [CI][18][35]
    float* vs = x1s + CI * UP_XH * UP_XP;                // [CI][12][20+1]
    float* c1s = vs + CI * UP_VH * (UP_VW + 1);          // [32]
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;

    const int h = p.h, w = p.w, H = 2 * h, W = 2 * w;
    const int tiles_x = (W + UP_TW - 1) / UP_TW;
    const int ty0 = (blockIdx.x / tiles_x) * UP_TH, tx0 = (blockIdx.x % tiles_x) * UP_TW;
    const int b = blockIdx.y, tid = threadIdx.x, warp = (int)tc_uniform((uint32_t)(tid >> 5)), half = warp >> 2;   // provably warp-uniform
    const float sy = (float)(h - 1) / (float)(H - 1), sx = (float)(w - 1) / (float)(W - 1);
    const int vy0 = (int)(sy * (float)max(ty0 - 1, 0)), vx0 = (int)(sx * (float)max(tx0 - 1, 0));
    const long long plane_lo = (long long)h * w, plane_hi = (long long)H * W;
    const float* vb = p.v + (long long)b * p.v_bs;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < NCH1 * BT / 4; i += 256) {
        const int c = i / (BT / 4), r = i - c * (BT / 4);
        reinterpret_cast<float4*>(b1t)[(c * 2 + 0) * (BT / 4) + r] = __ldg(reinterpret_cast<const float4*>(w1_hi) + i);
        reinterpret_cast<float4*>(b1t)[(c * 2 + 1) * (BT / 4) + r] = __ldg(reinterpret_cast<const float4*>(w1_lo) + i);
    }
    for (int i = tid; i < NCH2 * BT / 4; i += 256) {
        const int c = i / (BT / 4), r = i - c * (BT / 4);
        reinterpret_cast<float4*>(b2t)[(c * 2 + 0) * (BT / 4) + r] = __ldg(reinterpret_cast<const float4*>(w2_hi) + i);
        reinterpret_cast<float4*>(b2t)[(c * 2 + 1) * (BT / 4) + r] = __ldg(reinterpret_cast<const float4*>(w2_lo) + i);
    }
    if (tid < UP_C1) c1s[tid] = p.c1[tid];
    {
        // every load of the low-resolution tile is issued before the first store: ncu showed 35 % of ALL stall samples of the tensor-core
        // kernel on the one STS of the former load -> store loop (15 dependent global-load latencies per CTA, back to back)
        constexpr int V_IT = (CI * UP_VH * UP_VW + 255) / 256;
        float vreg[V_IT];
#pragma unroll
        for (int it = 0; it < V_IT; ++it) {
            const int i = tid + it * 256;
            const int c = i / (UP_VH * UP_VW);
            const int r = i - c * (UP_VH * UP_VW);
            const int yy = r / UP_VW, xx = r - yy * UP_VW;
            const int gy = min(vy0 + yy, h - 1), gx = min(vx0 + xx, w - 1);
            vreg[it] = i < CI * UP_VH * UP_VW ? __ldg(vb + (long long)c * plane_lo + gy * w + gx) : 0.f;
        }
#pragma unroll
        for (int it = 0; it < V_IT; ++it) {
            const int i = tid + it * 256;
            const int c = i / (UP_VH * UP_VW);
            const int r = i - c * (UP_VH * UP_VW);
            const int yy = r / UP_VW, xx = r - yy * UP_VW;
            if (i < CI * UP_VH * UP_VW) vs[(c * UP_VH + yy) * (UP_VW + 1) + xx] = vreg[it];
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // weight tiles are read by the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- x1 on the halo tile (0 outside the image: dw zero padding)
    for (int i = tid; i < UP_XH * UP_XW; i += 256) {
        const int yy = i / UP_XW, xx = i - yy * UP_XW;
        const int gy = ty0 - 1 + yy, gx = tx0 - 1 + xx;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        int y0 = 0, y1 = 0, x0 = 0, x1 = 0;
        float ly = 0.f, lx = 0.f;
        if (in) {
            bilin_src(gy, sy, h, y0, y1, ly);
            bilin_src(gx, sx, w, x0, x1, lx);
            y0 -= vy0; y1 -= vy0; x0 -= vx0; x1 -= vx0;
        }
        const float hy = 1.f - ly, hx = 1.f - lx;
#pragma unroll
        for (int c = 0; c < CI; ++c) {
            const float* vc = vs + c * UP_VH * (UP_VW + 1);
            float val = hy * (hx * vc[y0 * (UP_VW + 1) + x0] + lx * vc[y0 * (UP_VW + 1) + x1]) +
                        ly * (hx * vc[y1 * (UP_VW + 1) + x0] + lx * vc[y1 * (UP_VW + 1) + x1]);
            x1s[(c * UP_XH + yy) * UP_XP + xx] = in ? fmaxf(val + dwp.w[c * 12 + 11], 0.f) : 0.f;
        }
    }
    __syncthreads();

    const int col = tid & 31, rp = tid >> 5;
    const uint32_t tm = tc_uniform(tmem_base_s) + (uint32_t)(half * HALF_COLS);          // this half's columns
    const uint32_t t_lane = tm + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t mb = smem_u32(&mbar[half]);
    const uint32_t b1_s = smem_u32(b1t), b2_s = smem_u32(b2t);
    constexpr uint32_t idesc = tf32_idesc(32);
    float* ob = p.out + (long long)b * p.out_bs;
    const int gx = tx0 + col;
    uint32_t commits = 0;

    auto put16 = [&](const float (&a)[16], int col0) {      // 16 A values of this pixel -> hi / lo columns col0 .. col0+15
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            hi[j] = __float_as_uint(a[j]) & 0xffffe000u;
            lo[j] = __float_as_uint(a[j] - __uint_as_float(hi[j]));
        }
        tmem_st16(t_lane + (uint32_t)col0, hi);
        tmem_st16(t_lane + (uint32_t)(AK + col0), lo);
    };
    auto gemm = [&](uint32_t b_s, int n_chunks) {            // elected thread of the half: D = A . B over n_chunks K chunks
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
        if ((warp & 3) == 0) {   // warp-uniform branch, one elected lane issues (tc_common.cuh)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tc_elect_one()) {
                for (int c = 0; c < n_chunks; ++c) {
                    const uint32_t bh_s = b_s + (uint32_t)(c * 2) * BT * 4u, bl_s = bh_s + BT * 4u;
#pragma unroll
                    for (int ks = 0; ks < TC_KC / 8; ++ks) {
                        const uint32_t ah = tm + (uint32_t)(c * TC_KC + ks * 8), al = ah + (uint32_t)AK;
                        const uint64_t bh = kmajor_desc(bh_s, 32, ks), bl = kmajor_desc(bl_s, 32, ks);
                        mma_tf32_ts(tm + 2u * AK, ah, bh, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                        mma_tf32_ts(tm + 2u * AK, al, bh, idesc, 1u);
                        mma_tf32_ts(tm + 2u * AK, ah, bl, idesc, 1u);
                    }
                }
                tc_commit(mb);
            }
            __syncwarp();
        }
        mbar_wait(mb, commits & 1u);
        ++commits;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };

#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
        // ---- A1 = [x1 | x2] of pixel (2*rp + r, col), 16 k at a time
        const float* xrow = x1s + (2 * rp + r) * UP_XP + col;          // x1 halo row of the output row above this pixel
#pragma unroll
        for (int piece = 0; piece < NCH1; ++piece) {
            float a[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int k = piece * 16 + j;                           // compile-time
                if (k < CI) {
                    a[j] = xrow[(k * UP_XH + 1) * UP_XP + 1];           // x1: the window centre
                } else {
                    const int c = k - CI;
                    const float* xc = xrow + c * UP_XH * UP_XP;
                    const float* dk = dwp.w + c * 12;                   // compile-time offsets: constant-bank operands
                    float d = 0.f;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) d = fmaf(xc[ky * UP_XP + kx], dk[ky * 3 + kx], d);
                    a[j] = fmaxf(fmaf(dk[9], d, dk[10]), 0.f);
                }
            }
            put16(a, piece * 16);
        }
        gemm(b1_s, NCH1);
        // ---- t = relu(D + c1) -> A2 (over the dead A1 columns)
#pragma unroll
        for (int piece = 0; piece < NCH2; ++piece) {
            uint32_t d[16];
            tmem_ld16(t_lane + (uint32_t)(2 * AK + piece * 16), d);
            float a[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = fmaxf(__uint_as_float(d[j]) + c1s[piece * 16 + j], 0.f);
            put16(a, piece * 16);
        }
        gemm(b2_s, NCH2);
        // ---- v' = D[:, 0:16]
        {
            uint32_t d[16];
            tmem_ld16(t_lane + (uint32_t)(2 * AK), d);
            const int gy = ty0 + 2 * rp + r;
            if (gy < H && gx < W) {
#pragma unroll
                for (int o = 0; o < UP_N2; ++o) ob[(long long)o * plane_hi + (long long)gy * W + gx] = __uint_as_float(d[o]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // D / A reads before the next round's writes
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(TMEM_COLS) : "memory");
}

template <int CI>
static int launch_up_ghost_pw2_tc(const AchUpGhostPw2& p, const float* w1_hi, const float* w1_lo, const float* w2_hi, const float* w2_lo,
                                  const float* dw_host, cudaStream_t st) {
    constexpr int NCH1 = 2 * CI / TC_KC, NCH2 = UP_C1 / TC_KC;
    const size_t smem = (size_t)((NCH1 + NCH2) * 2 * 32 * TC_KC + CI * UP_XH * UP_XP + CI * UP_VH * (UP_VW + 1) + UP_C1) * sizeof(float);
    UpDwParams<CI> dwp;
    memcpy(dwp.w, dw_host, sizeof(dwp.w));
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(up_ghost_pw2_tc_kernel<CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    const int H = 2 * p.h, W = 2 * p.w;
    dim3 grid(cdiv(W, UP_TW) * cdiv(H, UP_TH), p.B);
    up_ghost_pw2_tc_kernel<CI><<<grid, 256, smem, st>>>(p, dwp, w1_hi, w1_lo, w2_hi, w2_lo);
    return check_launch("ach_up_ghost_pw2_tc");
}

template <int CI>
static int launch_up_ghost_pw2(const AchUpGhostPw2& p, cudaStream_t st) {
    const size_t smem = (size_t)(CI * UP_XH * UP_XP + CI * UP_VH * (UP_VW + 1) + 2 * CI * UP_C1 + UP_C1 * UP_N2 + CI * 12 + UP_C1) * sizeof(float);
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(up_ghost_pw2_kernel<CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    const int H = 2 * p.h, W = 2 * p.w;
    dim3 grid(cdiv(W, UP_TW) * cdiv(H, UP_TH), p.B);
    up_ghost_pw2_kernel<CI><<<grid, 256, smem, st>>>(p);
    return check_launch("ach_up_ghost_pw2");
}

}  // namespace ach

extern "C" int ach_up_ghost_pw2_supported(int ci, int c1, int n2) {
    return (ci == 16 || ci == 24 || ci == 32) && c1 == ach::UP_C1 && n2 == ach::UP_N2;
}

extern "C" int ach_up_ghost_pw2(const AchUpGhostPw2* pp, void* stream) {
    using namespace ach;
    const AchUpGhostPw2& p = *pp;
    ACH_REQUIRE(p.v && p.out && p.b1 && p.w2 && p.s2 && p.b2 && p.w1t && p.c1 && p.w2t, "ach_up_ghost_pw2: null arg");
    ACH_REQUIRE(p.B > 0 && p.B <= 65535 && p.h > 1 && p.w > 1, "ach_up_ghost_pw2: bad dims");
    ACH_REQUIRE(ach_up_ghost_pw2_supported(p.Ci, p.C1, p.N2), "ach_up_ghost_pw2: (Ci=%d, C1=%d, N2=%d) not instantiated", p.Ci, p.C1, p.N2);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (p.Ci) {
        case 16: return launch_up_ghost_pw2<16>(p, st);
        case 24: return launch_up_ghost_pw2<24>(p, st);
        default: return launch_up_ghost_pw2<32>(p, st);
    }
}

extern "C" int ach_up_ghost_pw2_tc_supported(int ci, int c1, int n2) {
    return (ci == 16 || ci == 24 || ci == 32) && c1 == ach::UP_C1 && n2 == ach::UP_N2;
}

extern "C" int ach_up_ghost_pw2_tc(const AchUpGhostPw2* pp, const float* w1_hi, const float* w1_lo, const float* w2_hi, const float* w2_lo,
                                   const float* dw_host, void* stream) {
    using namespace ach;
    const AchUpGhostPw2& p = *pp;
    ACH_REQUIRE(p.v && p.out && p.c1 && w1_hi && w1_lo && w2_hi && w2_lo && dw_host, "ach_up_ghost_pw2_tc: null arg");
    ACH_REQUIRE(p.B > 0 && p.B <= 65535 && p.h > 1 && p.w > 1, "ach_up_ghost_pw2_tc: bad dims");
    ACH_REQUIRE(ach_up_ghost_pw2_tc_supported(p.Ci, p.C1, p.N2), "ach_up_ghost_pw2_tc: (Ci=%d, C1=%d, N2=%d) not instantiated", p.Ci, p.C1, p.N2);
    ACH_REQUIRE(aligned16(w1_hi) && aligned16(w1_lo) && aligned16(w2_hi) && aligned16(w2_lo), "ach_up_ghost_pw2_tc: weight tiles must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (p.Ci) {
        case 16: return launch_up_ghost_pw2_tc<16>(p, w1_hi, w1_lo, w2_hi, w2_lo, dw_host, st);
        case 24: return launch_up_ghost_pw2_tc<24>(p, w1_hi, w1_lo, w2_hi, w2_lo, dw_host, st);
        default: return launch_up_ghost_pw2_tc<32>(p, w1_hi, w1_lo, w2_hi, w2_lo, dw_host, st);
    }
}
