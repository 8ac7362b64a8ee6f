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
// ach_up_ghost_head: 30x30 output tile; x1 on 34x34 (16 channels, shared memory), x2 recomputed on the fly,
// head primary p on 32x32 (one 1x4 column strip per thread = exactly 256 work items), then the head's
// cheap dw3x3 from shared p.  p aliases the low-res tile (dead after x1 is built).
constexpr int UH_T = 30;
constexpr int UH_P = UH_T + 2;           // 32
constexpr int UH_X1 = UH_T + 4;          // 34
constexpr int UH_V = UH_X1 / 2 + 3;      // 20
constexpr int UH_C = 16;                 // channels of v / x1 / x2
constexpr int UH_X1P = UH_X1 + 1;        // row pitch

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

// ARGMAX: instead of the K logit planes the kernel writes the per-pixel class index (first maximum, like torch.argmax
// over the very same fp32 logits) as one byte - 4*K bytes per pixel less to write, copy out and all-gather
// (achelous.py:283-297 only ever uses the argmax of these maps).  Classes whose bit in keep_mask is clear are mapped to 0
// (the reference's `output_seg[(output_seg != 0) & (output_seg != 8)] = 0`).
template <int INIT, int KOUT, bool ARGMAX>
__global__ void __launch_bounds__(256, 2) up_ghost_head_kernel(const float* __restrict__ v, long long v_bs, float* __restrict__ out,
                                                               long long out_bs, int h, int w,
                                                               const __grid_constant__ UpGhostHeadParams<INIT, KOUT> P,
                                                               unsigned char* __restrict__ mask, long long mask_bs, unsigned keep_mask) {
    extern __shared__ __align__(16) float smem[];
    float* x1s = smem;                                  // [16][34][35]
    float* vs = smem + UH_C * UH_X1 * UH_X1P;           // [16][20][21]
    float* ps = vs;                                     // [INIT][32][33]  (aliases vs)
    constexpr int VP = UH_V + 1, PP = UH_P + 1;
    static_assert(INIT * UH_P * PP <= UH_C * UH_V * VP, "p tile must fit in the low-res tile");

    const int H = 2 * h, W = 2 * w;
    const int tiles_x = (W + UH_T - 1) / UH_T;
    const int ty0 = (blockIdx.x / tiles_x) * UH_T, tx0 = (blockIdx.x % tiles_x) * UH_T;
    const int b = blockIdx.y;
    const float sy = (float)(h - 1) / (float)(H - 1), sx = (float)(w - 1) / (float)(W - 1);
    const int vy0 = (int)(sy * (float)max(ty0 - 2, 0)), vx0 = (int)(sx * (float)max(tx0 - 2, 0));
    const long long plane_lo = (long long)h * w, plane_hi = (long long)H * W;
    const float* vb = v + (long long)b * v_bs;

    // ---- low-res tile, all 16 channels
    for (int i = threadIdx.x; i < UH_C * UH_V * UH_V; i += 256) {
        const int c = i / (UH_V * UH_V);
        const int r = i - c * (UH_V * UH_V);
        const int yy = r / UH_V, xx = r - yy * UH_V;
        const int gy = min(vy0 + yy, h - 1), gx = min(vx0 + xx, w - 1);
        vs[(c * UH_V + yy) * VP + xx] = __ldg(vb + (long long)c * plane_lo + gy * w + gx);
    }
    __syncthreads();

    // ---- x1 = relu(up(v) + b1) on the 34x34 halo tile (0 outside the image = dw zero padding)
    for (int i = threadIdx.x; i < UH_X1 * UH_X1; i += 256) {
        const int yy = i / UH_X1, xx = i - yy * UH_X1;
        const int gy = ty0 - 2 + yy, gx = tx0 - 2 + xx;
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
        for (int c = 0; c < UH_C; ++c) {
            const float* vc = vs + c * UH_V * VP;
            float val = hy * (hx * vc[y0 * VP + x0] + lx * vc[y0 * VP + x1]) + ly * (hx * vc[y1 * VP + x0] + lx * vc[y1 * VP + x1]);
            val = in ? fmaxf(val + P.b1[c], 0.f) : 0.f;
            x1s[(c * UH_X1 + yy) * UH_X1P + xx] = val;
        }
    }
    __syncthreads();  // vs is dead from here on; ps (alias) may be written

    // ---- head primary p = relu(w3 . [x1, x2] + b3) on the 32x32 tile, x2 = relu(s2*dw3x3(x1)+b2) on the fly.
    // thread -> column `col` (0..31), strip of 4 rows starting at 4*strip (0..7): 6x3 x1 values serve 4 pixels.
    {
        const int col = threadIdx.x & 31, strip = threadIdx.x >> 5;
        float acc[4][INIT];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < INIT; ++i) acc[r][i] = P.b3[i];
#pragma unroll
        for (int c = 0; c < UH_C; ++c) {
            const float* xc = x1s + (c * UH_X1 + 4 * strip) * UH_X1P + col;  // x1 row (4*strip), col: p pixel (r, col) centre = x1[r+1][col+1]
            float win[6][3];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k) win[r][k] = xc[r * UH_X1P + k];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float d = 0.f;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) d = fmaf(win[r + ky][kx], P.w2[c * 9 + ky * 3 + kx], d);
                const float x2 = fmaxf(fmaf(P.s2[c], d, P.b2[c]), 0.f);
                const float x1c = win[r + 1][1];
#pragma unroll
                for (int i = 0; i < INIT; ++i) {
                    acc[r][i] = fmaf(x1c, P.w3[c * INIT + i], acc[r][i]);
                    acc[r][i] = fmaf(x2, P.w3[(UH_C + c) * INIT + i], acc[r][i]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int py = 4 * strip + r;
            const int gy = ty0 - 1 + py, gx = tx0 - 1 + col;
            const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
#pragma unroll
            for (int i = 0; i < INIT; ++i) ps[(i * UH_P + py) * PP + col] = in ? fmaxf(acc[r][i], 0.f) : 0.f;
        }
    }
    __syncthreads();

    // ---- outputs: channels [0, INIT) = p, [INIT, KOUT) = relu(s4 * dw3x3(p) + b4)
    float* ob = ARGMAX ? nullptr : out + (long long)b * out_bs;
    unsigned char* mb = ARGMAX ? mask + (long long)b * mask_bs : nullptr;
    for (int i = threadIdx.x; i < UH_T * UH_T; i += 256) {
        const int yy = i / UH_T, xx = i - yy * UH_T;
        const int gy = ty0 + yy, gx = tx0 + xx;
        if (gy >= H || gx >= W) continue;
        const long long o = (long long)gy * W + gx;
        float best = -INFINITY;
        int arg = 0;
#pragma unroll
        for (int k = 0; k < INIT; ++k) {
            const float val = ps[(k * UH_P + yy + 1) * PP + xx + 1];
            if (ARGMAX) {
                if (val > best) { best = val; arg = k; }
            } else {
                ob[k * plane_hi + o] = val;
            }
        }
#pragma unroll
        for (int k = 0; k < KOUT - INIT; ++k) {
            float d = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) d = fmaf(ps[(k * UH_P + yy + ky) * PP + xx + kx], P.w4[k * 9 + ky * 3 + kx], d);
            const float val = fmaxf(fmaf(P.s4[k], d, P.b4[k]), 0.f);
            if (ARGMAX) {
                if (val > best) { best = val; arg = INIT + k; }
            } else {
                ob[(INIT + k) * plane_hi + o] = val;
            }
        }
        if (ARGMAX) mb[o] = (unsigned char)(((keep_mask >> arg) & 1u) ? arg : 0);
    }
}

template <int INIT, int KOUT, bool ARGMAX>
static int launch_head(const AchUpGhostHead& a, cudaStream_t st, unsigned char* mask = nullptr, long long mask_bs = 0,
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
    constexpr size_t smem = (size_t)(UH_C * UH_X1 * UH_X1P + UH_C * UH_V * (UH_V + 1)) * sizeof(float);
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(up_ghost_head_kernel<INIT, KOUT, ARGMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    const int H = 2 * a.h, W = 2 * a.w;
    dim3 grid(cdiv(W, UH_T) * cdiv(H, UH_T), a.B);
    up_ghost_head_kernel<INIT, KOUT, ARGMAX><<<grid, 256, smem, st>>>(a.v, a.v_bs, a.out, a.out_bs, a.h, a.w, P, mask, mask_bs, keep_mask);
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
    if (a.init == 1) return launch_head<1, 2, false>(a, st);
    return launch_head<5, 9, false>(a, st);
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
    if (a.init == 1) return launch_head<1, 2, true>(a, st, mask, mask_bs, keep_mask);
    return launch_head<5, 9, true>(a, st, mask, mask_bs, keep_mask);
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
    for (int i = tid; i < CI * UP_VH * UP_VW; i += 256) {
        const int c = i / (UP_VH * UP_VW);
        const int r = i - c * (UP_VH * UP_VW);
        const int yy = r / UP_VW, xx = r - yy * UP_VW;
        const int gy = min(vy0 + yy, h - 1), gx = min(vx0 + xx, w - 1);
        vs[(c * UP_VH + yy) * (UP_VW + 1) + xx] = __ldg(vb + (long long)c * plane_lo + gy * w + gx);
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
template <int CI>
__global__ void __launch_bounds__(256, 2) up_ghost_pw2_tc_kernel(const AchUpGhostPw2 p, const float* __restrict__ w1_hi,
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
    float* dws = vs + CI * UP_VH * (UP_VW + 1);          // [CI][12]: 9 taps, s2, b2, b1
    float* c1s = dws + CI * 12;                          // [32]
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;

    const int h = p.h, w = p.w, H = 2 * h, W = 2 * w;
    const int tiles_x = (W + UP_TW - 1) / UP_TW;
    const int ty0 = (blockIdx.x / tiles_x) * UP_TH, tx0 = (blockIdx.x % tiles_x) * UP_TW;
    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, half = tid >> 7;
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
    for (int i = tid; i < CI * 12; i += 256) {
        const int c = i / 12, k = i - c * 12;
        dws[i] = k < 9 ? p.w2[c * 9 + k] : (k == 9 ? p.s2[c] : (k == 10 ? p.b2[c] : p.b1[c]));
    }
    if (tid < UP_C1) c1s[tid] = p.c1[tid];
    for (int i = tid; i < CI * UP_VH * UP_VW; i += 256) {
        const int c = i / (UP_VH * UP_VW);
        const int r = i - c * (UP_VH * UP_VW);
        const int yy = r / UP_VW, xx = r - yy * UP_VW;
        const int gy = min(vy0 + yy, h - 1), gx = min(vx0 + xx, w - 1);
        vs[(c * UP_VH + yy) * (UP_VW + 1) + xx] = __ldg(vb + (long long)c * plane_lo + gy * w + gx);
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
#pragma unroll 4
        for (int c = 0; c < CI; ++c) {
            const float* vc = vs + c * UP_VH * (UP_VW + 1);
            float val = hy * (hx * vc[y0 * (UP_VW + 1) + x0] + lx * vc[y0 * (UP_VW + 1) + x1]) +
                        ly * (hx * vc[y1 * (UP_VW + 1) + x0] + lx * vc[y1 * (UP_VW + 1) + x1]);
            x1s[(c * UP_XH + yy) * UP_XP + xx] = in ? fmaxf(val + dws[c * 12 + 11], 0.f) : 0.f;
        }
    }
    __syncthreads();

    const int col = tid & 31, rp = tid >> 5;
    const uint32_t tm = tmem_base_s + (uint32_t)(half * HALF_COLS);          // this half's columns
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
        if ((tid & 127) == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
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
                    const float* dk = dws + c * 12;
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
                                  cudaStream_t st) {
    constexpr int NCH1 = 2 * CI / TC_KC, NCH2 = UP_C1 / TC_KC;
    const size_t smem = (size_t)((NCH1 + NCH2) * 2 * 32 * TC_KC + CI * UP_XH * UP_XP + CI * UP_VH * (UP_VW + 1) + CI * 12 + UP_C1) * sizeof(float);
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(up_ghost_pw2_tc_kernel<CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    const int H = 2 * p.h, W = 2 * p.w;
    dim3 grid(cdiv(W, UP_TW) * cdiv(H, UP_TH), p.B);
    up_ghost_pw2_tc_kernel<CI><<<grid, 256, smem, st>>>(p, w1_hi, w1_lo, w2_hi, w2_lo);
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
                                   void* stream) {
    using namespace ach;
    const AchUpGhostPw2& p = *pp;
    ACH_REQUIRE(p.v && p.out && p.b1 && p.w2 && p.s2 && p.b2 && p.c1 && w1_hi && w1_lo && w2_hi && w2_lo, "ach_up_ghost_pw2_tc: null arg");
    ACH_REQUIRE(p.B > 0 && p.B <= 65535 && p.h > 1 && p.w > 1, "ach_up_ghost_pw2_tc: bad dims");
    ACH_REQUIRE(ach_up_ghost_pw2_tc_supported(p.Ci, p.C1, p.N2), "ach_up_ghost_pw2_tc: (Ci=%d, C1=%d, N2=%d) not instantiated", p.Ci, p.C1, p.N2);
    ACH_REQUIRE(aligned16(w1_hi) && aligned16(w1_lo) && aligned16(w2_hi) && aligned16(w2_lo), "ach_up_ghost_pw2_tc: weight tiles must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (p.Ci) {
        case 16: return launch_up_ghost_pw2_tc<16>(p, w1_hi, w1_lo, w2_hi, w2_lo, st);
        case 24: return launch_up_ghost_pw2_tc<24>(p, w1_hi, w1_lo, w2_hi, w2_lo, st);
        default: return launch_up_ghost_pw2_tc<32>(p, w1_hi, w1_lo, w2_hi, w2_lo, st);
    }
}
