// Small HBM-bound kernels of the neck / fusion / PointNet path plus the error plumbing of the C ABI.
// All are single-pass, coalesced along the contiguous pixel axis; reductions use warp shuffles.
#include <cstdarg>

#include "common.cuh"

namespace ach {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return ACH_ERR_CUDA;
    }
    return ACH_OK;
}

// ------------------------------------------------------------------ channels-first LayerNorm
__global__ void __launch_bounds__(256) layernorm_cf_kernel(const float* __restrict__ x, long long x_bs,
                                                           const float* __restrict__ w, const float* __restrict__ bb,
                                                           float* __restrict__ out, long long out_bs, int C, int P, float eps) {
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= P) return;
    const float* xb = x + (long long)blockIdx.y * x_bs + p;
    float* ob = out + (long long)blockIdx.y * out_bs + p;
    float mean = 0.f;
    for (int c = 0; c < C; ++c) mean += xb[(long long)c * P];
    mean /= (float)C;
    float var = 0.f;
    for (int c = 0; c < C; ++c) {
        const float d = xb[(long long)c * P] - mean;
        var = fmaf(d, d, var);
    }
    const float rstd = 1.0f / sqrtf(var / (float)C + eps);
    for (int c = 0; c < C; ++c) ob[(long long)c * P] = w[c] * ((xb[(long long)c * P] - mean) * rstd) + bb[c];
}

// ------------------------------------------------------------------ channels-first LayerNorm + 2x2 space-to-depth
// EdgeNeXt downsample layers are LayerNorm -> Conv2d(k=2, s=2) (edgenext.py:29-34): a patchify conv is a GEMM
// over K = 4C once the 2x2 patch is moved into the channel axis.  One thread normalises one input pixel over its
// C channels (coalesced across the warp) and scatters it to channel c*4 + (y&1)*2 + (x&1) of the half-resolution
// map - exactly the flatten order of the conv weight (O, C, 2, 2) - so the conv itself runs on the GEMM kernels.
__global__ void __launch_bounds__(256) ln_s2d_kernel(const float* __restrict__ x, long long x_bs, const float* __restrict__ w,
                                                     const float* __restrict__ bb, float* __restrict__ out, long long out_bs, int C,
                                                     int H, int W, float eps) {
    const int P = H * W;
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= P) return;
    const int y = p / W, xx = p - y * W;
    const float* xb = x + (long long)blockIdx.y * x_bs + p;
    // statistics in ONE pass over the channels, 8 independent loads in flight at a time (the three dependent runtime-length loops of the
    // first version left 15 long-scoreboard stalls per issued instruction).  Sums are taken of (x - x[0]): no cancellation in
    // E[d^2] - E[d]^2 for LayerNorm-sized variances, same scheme as the GEMM kernels' fused LayerNorm.
    const float shift = xb[0];
    float s1 = 0.f, s2 = 0.f;
    int c = 0;
    for (; c + 8 <= C; c += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = xb[(long long)(c + j) * P];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = v[j] - shift;
            s1 += d;
            s2 = fmaf(d, d, s2);
        }
    }
    for (; c < C; ++c) {
        const float d = xb[(long long)c * P] - shift;
        s1 += d;
        s2 = fmaf(d, d, s2);
    }
    const float m1 = s1 / (float)C;
    const float mean = shift + m1;
    const float rstd = 1.0f / sqrtf(fmaxf(s2 / (float)C - m1 * m1, 0.f) + eps);
    const int Po = (H / 2) * (W / 2);
    float* ob = out + (long long)blockIdx.y * out_bs + (long long)(((y & 1) << 1) | (xx & 1)) * Po + (y >> 1) * (W / 2) + (xx >> 1);
    c = 0;
    for (; c + 8 <= C; c += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = xb[(long long)(c + j) * P];
#pragma unroll
        for (int j = 0; j < 8; ++j) ob[(long long)(c + j) * 4 * Po] = w[c + j] * ((v[j] - mean) * rstd) + bb[c + j];
    }
    for (; c < C; ++c) ob[(long long)c * 4 * Po] = w[c] * ((xb[(long long)c * P] - mean) * rstd) + bb[c];
}

// ------------------------------------------------------------------ bilinear x2, align_corners=True
// thread = 4 consecutive output columns of one (frame, channel, row): one 16-byte store, 32-bit index arithmetic (the
// first version decoded a 64-bit linear index per output element and took 1.3 ms on a 32 x 320^2 x 64 map)
template <int VEC>
__global__ void __launch_bounds__(256) upsample2x_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ out,
                                                         long long out_bs, int H, int W) {
    const int Ho = 2 * H, Wo = 2 * W;
    const int WV = Wo / VEC;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= Ho * WV) return;
    const int oy = i / WV, ox0 = (i - oy * WV) * VEC;
    const int c = blockIdx.y;
    // ATen upsample_bilinear2d, align_corners: src = dst * (in - 1) / (out - 1)
    const float sy = (Ho > 1) ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
    const float sx = (Wo > 1) ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
    const float fy = sy * (float)oy;
    const int y0 = (int)fy;
    const int y1 = y0 + (y0 < H - 1);
    const float ly = fy - (float)y0, hy = 1.f - ly;
    const float* __restrict__ r0 = x + (long long)blockIdx.z * x_bs + ((long long)c * H + y0) * W;
    const float* __restrict__ r1 = x + (long long)blockIdx.z * x_bs + ((long long)c * H + y1) * W;
    float v[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const float fx = sx * (float)(ox0 + j);
        const int x0 = (int)fx;
        const int x1 = x0 + (x0 < W - 1);
        const float lx = fx - (float)x0, hx = 1.f - lx;
        v[j] = hy * (hx * __ldg(r0 + x0) + lx * __ldg(r0 + x1)) + ly * (hx * __ldg(r1 + x0) + lx * __ldg(r1 + x1));
    }
    float* o = out + (long long)blockIdx.z * out_bs + ((long long)c * Ho + oy) * Wo + ox0;
    if constexpr (VEC == 4) {
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        o[0] = v[0];
    }
}

// ------------------------------------------------------------------ SPP max pools 5 / 9 / 13
__global__ void __launch_bounds__(256) spp_maxpool_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ o5,
                                                          float* __restrict__ o9, float* __restrict__ o13, long long out_bs,
                                                          int C, int H, int W) {
    extern __shared__ float plane[];
    const int c = blockIdx.x, b = blockIdx.y;
    const int P = H * W;
    const float* xp = x + (long long)b * x_bs + (long long)c * P;
    for (int i = threadIdx.x; i < P; i += 256) plane[i] = xp[i];
    __syncthreads();
    for (int i = threadIdx.x; i < P; i += 256) {
        const int y = i / W, xx = i - y * W;
        float m5 = -INFINITY, m9 = -INFINITY, m13 = -INFINITY;
        for (int dy = -6; dy <= 6; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
            for (int dx = -6; dx <= 6; ++dx) {
                const int xc = xx + dx;
                if (xc < 0 || xc >= W) continue;
                const float v = plane[yy * W + xc];
                m13 = fmaxf(m13, v);
                if (dy >= -4 && dy <= 4 && dx >= -4 && dx <= 4) m9 = fmaxf(m9, v);
                if (dy >= -2 && dy <= 2 && dx >= -2 && dx <= 2) m5 = fmaxf(m5, v);
            }
        }
        const long long off = (long long)b * out_bs + (long long)c * P + i;
        o5[off] = m5;
        o9[off] = m9;
        o13[off] = m13;
    }
}

// ------------------------------------------------------------------ ShuffleAttention
// One CTA per (b, source channel).  Source channel ch = g * cg + j (cg = C / G); j < cg/2 -> channel
// gate x * sigmoid(cweight * mean + cbias); else spatial gate x * sigmoid(sweight * GN(x) + sbias) with
// one-channel groups.  Destination channel after channel_shuffle(., 2): (ch % (C/2)) * 2 + ch / (C/2).
__global__ void __launch_bounds__(256) shuffle_attention_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ out,
                                                                long long out_bs, const float* __restrict__ cweight,
                                                                const float* __restrict__ cbias, const float* __restrict__ sweight,
                                                                const float* __restrict__ sbias, const float* __restrict__ gn_w,
                                                                const float* __restrict__ gn_b, int C, int P, int G, float eps) {
    __shared__ float red[8];
    const int ch = blockIdx.x, b = blockIdx.y;
    const int cg = C / G, half = cg / 2;
    const int j = ch % cg;
    const float* xp = x + (long long)b * x_bs + (long long)ch * P;
    const int dst = (ch % (C / 2)) * 2 + ch / (C / 2);
    float* op = out + (long long)b * out_bs + (long long)dst * P;
    float s = 0.f;
    for (int i = threadIdx.x; i < P; i += 256) s += xp[i];
    const float mean = block_sum_256(s, red) / (float)P;
    if (j < half) {
        const float gate = sigmoidf_(fmaf(cweight[j], mean, cbias[j]));
        for (int i = threadIdx.x; i < P; i += 256) op[i] = xp[i] * gate;
    } else {
        const int q = j - half;
        float v = 0.f;
        for (int i = threadIdx.x; i < P; i += 256) {
            const float d = xp[i] - mean;
            v = fmaf(d, d, v);
        }
        const float rstd = 1.0f / sqrtf(block_sum_256(v, red) / (float)P + eps);
        const float gw = gn_w[q], gb = gn_b[q], sw = sweight[q], sb = sbias[q];
        for (int i = threadIdx.x; i < P; i += 256) {
            const float xv = xp[i];
            const float n = fmaf((xv - mean) * rstd, gw, gb);
            op[i] = xv * sigmoidf_(fmaf(sw, n, sb));
        }
    }
}

// ------------------------------------------------------------------ plane mean, ECA fuse
__global__ void __launch_bounds__(256) plane_mean_kernel(const float* __restrict__ x, long long x_bs, const float* __restrict__ x2,
                                                         long long x2_bs, float* __restrict__ out, int C, int P) {
    __shared__ float red[8];
    const int c = blockIdx.x, b = blockIdx.y;
    const float* xp = x + (long long)b * x_bs + (long long)c * P;
    const float* yp = x2 ? x2 + (long long)b * x2_bs + (long long)c * P : nullptr;
    float s = 0.f;
    for (int i = threadIdx.x; i < P; i += 256) s += yp ? xp[i] + yp[i] : xp[i];
    const float t = block_sum_256(s, red);
    if (threadIdx.x == 0) out[(long long)b * C + c] = t / (float)P;
}

__global__ void __launch_bounds__(256) eca_fuse_kernel(const float* __restrict__ x, long long x_bs, const float* __restrict__ x2,
                                                       long long x2_bs, const float* __restrict__ mean, const float* __restrict__ w1d,
                                                       int k1d, const float* __restrict__ scale, const float* __restrict__ bias,
                                                       float* __restrict__ out, long long out_bs, int C, int P) {
    const int c = blockIdx.y, b = blockIdx.z;
    float a = 0.f;
    const int r = (k1d - 1) / 2;
    for (int t = 0; t < k1d; ++t) {
        const int cc = c + t - r;
        if (cc >= 0 && cc < C) a = fmaf(w1d[t], mean[(long long)b * C + cc], a);
    }
    const float gate = sigmoidf_(a);
    const float s = scale[c], bi = bias[c];
    const float* xp = x + (long long)b * x_bs + (long long)c * P;
    const float* yp = x2 ? x2 + (long long)b * x2_bs + (long long)c * P : nullptr;
    float* op = out + (long long)b * out_bs + (long long)c * P;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < P; i += gridDim.x * 256) {
        const float v = yp ? xp[i] + yp[i] : xp[i];
        op[i] = fmaxf(fmaf(s, v * gate, bi), 0.f);
    }
}

// ------------------------------------------------------------------ avgpool 3x3 (count_include_pad)
__global__ void __launch_bounds__(256) avgpool3_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ out,
                                                       long long out_bs, int C, int H, int W) {
    // one thread = 4 horizontally adjacent outputs of one row (32-bit index math; rows reused from registers)
    const int W4 = (W + 3) >> 2;
    const int n = C * H * W4;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int x4 = i % W4;
    const int row = i / W4;          // c * H + y
    const int y = row % H;
    const int x0 = x4 * 4;
    const float* xp = x + (long long)blockIdx.y * x_bs + (long long)(row - y) * W;   // plane base
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        const float* r = xp + yy * W;
        float v[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int xc = x0 - 1 + j;
            v[j] = (xc >= 0 && xc < W) ? r[xc] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += v[j] + v[j + 1] + v[j + 2];
    }
    float* op = out + (long long)blockIdx.y * out_bs + (long long)row * W + x0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (x0 + j < W) op[j] = acc[j] / 9.0f;
}

// channel-last variant: thread = 4 horizontally adjacent pixels of one row, all channels; per 4 channels it writes the
// 4 pixels' 16-byte groups (contiguous 64 B per thread when ceil4(C) = 4)
__global__ void __launch_bounds__(128) avgpool3_cl_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ out,
                                                          long long out_bs, int C, int H, int W) {
    const int W4 = (W + 3) >> 2;
    const int i0 = blockIdx.x * 128 + threadIdx.x;
    const bool live = i0 < H * W4;
    const int i = live ? i0 : H * W4 - 1;          // dead threads of the last CTA shadow the last strip (shuffles stay warp-wide)
    const int y = i / W4, x0 = (i - y * W4) * 4;
    const int P = H * W;
    const float* xb = x + (long long)blockIdx.y * x_bs;
    const int Q = (C + 3) >> 2;
    // vector path: every row start is 16-byte aligned; lane_l / lane_r: the neighbouring lane holds the adjacent strip of this row
    const bool vec = (W & 3) == 0 && (x_bs & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const int lane = threadIdx.x & 31;
    const bool lane_l = lane > 0 && x0 > 0, lane_r = lane < 31 && x0 + 4 < W && i0 + 1 < H * W4;
    float4* op = reinterpret_cast<float4*>(out + (long long)blockIdx.y * out_bs) + (long long)(y * W + x0) * Q;
    for (int q = 0; q < Q; ++q) {
        float o[4][4];   // [channel in quad][pixel]
        if (vec) {
            // all 12 row vectors of the quad (4 channels x 3 rows) are requested BEFORE the first one is used: ncu showed 7 long-scoreboard
            // stalls per issued instruction with the loads interleaved with their shuffles (1.8 TB/s on a pure streaming kernel)
            float4 m[4][3];
            float lf[4][3], rt[4][3];
            const bool need_l = x0 > 0 && !lane_l, need_r = x0 + 4 < W && !lane_r;   // neighbour strip not held by the adjacent lane
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = q * 4 + e;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = y + dy;
                    const bool ok = c < C && yy >= 0 && yy < H;
                    const float* r = xb + (long long)(ok ? c : 0) * P + (ok ? yy : y) * W;
                    m[e][dy + 1] = ok ? __ldg(reinterpret_cast<const float4*>(r + x0)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    lf[e][dy + 1] = (ok && need_l) ? __ldg(r + x0 - 1) : 0.f;
                    rt[e][dy + 1] = (ok && need_r) ? __ldg(r + x0 + 4) : 0.f;
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const float4 mm = m[e][dy];
                    const float lft = __shfl_up_sync(0xffffffffu, mm.w, 1), rgt = __shfl_down_sync(0xffffffffu, mm.x, 1);
                    float v[6];
                    v[1] = mm.x, v[2] = mm.y, v[3] = mm.z, v[4] = mm.w;
                    v[0] = x0 == 0 ? 0.f : (lane_l ? lft : lf[e][dy]);
                    v[5] = x0 + 4 >= W ? 0.f : (lane_r ? rgt : rt[e][dy]);
                    const int yy = y + dy - 1;
                    if (yy >= 0 && yy < H) {   // rows outside the image are skipped, as in avgpool3_kernel (same summation order)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[j] += v[j] + v[j + 1] + v[j + 2];
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) o[e][j] = acc[j] / 9.0f;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = q * 4 + e;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                if (c < C) {
                    const float* xp = xb + (long long)c * P;
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy) {
                        const int yy = y + dy;
                        if (yy < 0 || yy >= H) continue;
                        const float* r = xp + yy * W;
                        float v[6];
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            const int xc = x0 - 1 + j;
                            v[j] = (xc >= 0 && xc < W) ? __ldg(r + xc) : 0.f;
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[j] += v[j] + v[j + 1] + v[j + 2];   // same summation order as avgpool3_kernel
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) o[e][j] = acc[j] / 9.0f;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (live && x0 + j < W) op[(long long)j * Q + q] = make_float4(o[0][j], o[1][j], o[2][j], o[3][j]);
    }
}

// ------------------------------------------------------------------ fully connected on (B, K) rows
__global__ void __launch_bounds__(256) fc_kernel(const float* __restrict__ x, long long x_bs, const float* __restrict__ w,
                                                 const float* __restrict__ scale, const float* __restrict__ bias,
                                                 float* __restrict__ out, long long out_bs, int K, int O, int act) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = blockIdx.x * 8 + warp;
    const int b = blockIdx.y;
    if (o >= O) return;
    const float* xr = x + (long long)b * x_bs;
    const float* wr = w + (long long)o * K;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(wr[k], xr[k], s);
    s = warp_sum(s);
    if (lane == 0) {
        const float sc = scale ? scale[o] : 1.f;
        const float bi = bias ? bias[o] : 0.f;
        out[(long long)b * out_bs + o] = apply_act(fmaf(sc, s, bi), act);
    }
}

// ------------------------------------------------------------------ log_softmax over K, transposed output
__global__ void __launch_bounds__(256) logsoftmax_t_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ out,
                                                           long long out_bs, int K, int N) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.y;
    if (n >= N) return;
    const float* xp = x + (long long)b * x_bs + n;
    float v[32];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 32; ++k)
        if (k < K) {
            v[k] = xp[(long long)k * N];
            m = fmaxf(m, v[k]);
        }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k)
        if (k < K) s += expf(v[k] - m);
    const float lse = logf(s);
    float* op = out + (long long)b * out_bs + (long long)n * K;
#pragma unroll
    for (int k = 0; k < 32; ++k)
        if (k < K) op[k] = (v[k] - m) - lse;
}

// ------------------------------------------------------------------ copy (+ broadcast add), add, fill
__global__ void __launch_bounds__(256) copy_add_kernel(const float* __restrict__ x, long long x_bs, const float* __restrict__ post,
                                                       float* __restrict__ out, long long out_bs, long long n) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float v = x[(long long)blockIdx.y * x_bs + i];
    if (post) v += post[i];
    out[(long long)blockIdx.y * out_bs + i] = v;
}

__global__ void __launch_bounds__(256) add_kernel(const float* __restrict__ a, long long a_bs, const float* __restrict__ b2,
                                                  long long b_bs, float* __restrict__ out, long long out_bs, long long n) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    out[(long long)blockIdx.y * out_bs + i] = a[(long long)blockIdx.y * a_bs + i] + b2[(long long)blockIdx.y * b_bs + i];
}

__global__ void __launch_bounds__(256) fill_kernel(float* __restrict__ x, long long n, float v) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < n) x[i] = v;
}

}  // namespace ach

using namespace ach;

extern "C" const char* ach_last_error(void) { return g_err; }
extern "C" int ach_version(void) { return 100; }

extern "C" int ach_layernorm_cf(const float* x, long long x_bs, const float* w, const float* b, float* out, long long out_bs,
                                int B, int C, int P, float eps, void* stream) {
    ACH_REQUIRE(x && w && b && out && B > 0 && C > 0 && P > 0 && B <= 65535, "ach_layernorm_cf: bad args");
    layernorm_cf_kernel<<<dim3(cdiv(P, 256), B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, w, b, out, out_bs, C, P, eps);
    return check_launch("ach_layernorm_cf");
}

extern "C" int ach_ln_s2d(const float* x, long long x_bs, const float* w, const float* b, float* out, long long out_bs, int B,
                          int C, int H, int W, float eps, void* stream) {
    ACH_REQUIRE(x && w && b && out && B > 0 && C > 0 && B <= 65535, "ach_ln_s2d: bad args");
    ACH_REQUIRE(H % 2 == 0 && W % 2 == 0, "ach_ln_s2d: H, W must be even");
    ln_s2d_kernel<<<dim3(cdiv((long long)H * W, 256), B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, w, b, out, out_bs, C, H, W, eps);
    return check_launch("ach_ln_s2d");
}

extern "C" int ach_upsample2x(const float* x, long long x_bs, float* out, long long out_bs, int B, int C, int H, int W,
                              void* stream) {
    ACH_REQUIRE(x && out && B > 0 && C > 0 && H > 0 && W > 0 && B <= 65535, "ach_upsample2x: bad args");
    ACH_REQUIRE(C <= 65535 && (long long)4 * H * W < (1LL << 31), "ach_upsample2x: plane too large");
    const bool vec = (W % 2 == 0) && aligned16(out) && out_bs % 4 == 0;      // Wo = 2W is then a multiple of 4
    if (vec)
        upsample2x_kernel<4><<<dim3(cdiv((long long)2 * H * (2 * W / 4), 256), C, B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, out, out_bs, H, W);
    else
        upsample2x_kernel<1><<<dim3(cdiv((long long)4 * H * W, 256), C, B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, out, out_bs, H, W);
    return check_launch("ach_upsample2x");
}

extern "C" int ach_spp_maxpool(const float* x, long long x_bs, float* out5, float* out9, float* out13, long long out_bs,
                               int B, int C, int H, int W, void* stream) {
    ACH_REQUIRE(x && out5 && out9 && out13 && B > 0 && C > 0 && B <= 65535, "ach_spp_maxpool: bad args");
    ACH_REQUIRE((size_t)H * W * 4 <= 48 * 1024, "ach_spp_maxpool: plane %dx%d too large for the shared-memory path", H, W);
    spp_maxpool_kernel<<<dim3(C, B), 256, (size_t)H * W * 4, (cudaStream_t)stream>>>(x, x_bs, out5, out9, out13, out_bs, C, H, W);
    return check_launch("ach_spp_maxpool");
}

extern "C" int ach_shuffle_attention(const float* x, long long x_bs, float* out, long long out_bs, const float* cweight,
                                     const float* cbias, const float* sweight, const float* sbias, const float* gn_w,
                                     const float* gn_b, int B, int C, int P, int G, float eps, void* stream) {
    ACH_REQUIRE(x && out && cweight && cbias && sweight && sbias && gn_w && gn_b, "ach_shuffle_attention: null arg");
    ACH_REQUIRE(B > 0 && B <= 65535 && G > 0 && C % (2 * G) == 0, "ach_shuffle_attention: C=%d must be divisible by 2G=%d", C, 2 * G);
    shuffle_attention_kernel<<<dim3(C, B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, out, out_bs, cweight, cbias, sweight, sbias,
                                                                          gn_w, gn_b, C, P, G, eps);
    return check_launch("ach_shuffle_attention");
}

extern "C" int ach_plane_mean(const float* x, long long x_bs, const float* x2, long long x2_bs, float* out, int B, int C,
                              int P, void* stream) {
    ACH_REQUIRE(x && out && B > 0 && C > 0 && P > 0 && B <= 65535, "ach_plane_mean: bad args");
    plane_mean_kernel<<<dim3(C, B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, x2, x2_bs, out, C, P);
    return check_launch("ach_plane_mean");
}

extern "C" int ach_eca_fuse(const float* x, long long x_bs, const float* x2, long long x2_bs, const float* mean,
                            const float* w1d, int k1d, const float* scale, const float* bias, float* out, long long out_bs,
                            int B, int C, int P, void* stream) {
    ACH_REQUIRE(x && mean && w1d && scale && bias && out && B > 0 && C > 0 && P > 0, "ach_eca_fuse: bad args");
    ACH_REQUIRE(k1d % 2 == 1 && C <= 65535 && B <= 65535, "ach_eca_fuse: bad k/C/B");
    eca_fuse_kernel<<<dim3(min(cdiv(P, 256), 64), C, B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, x2, x2_bs, mean, w1d, k1d, scale,
                                                                                        bias, out, out_bs, C, P);
    return check_launch("ach_eca_fuse");
}

extern "C" int ach_avgpool3(const float* x, long long x_bs, float* out, long long out_bs, int B, int C, int H, int W,
                            void* stream) {
    ACH_REQUIRE(x && out && B > 0 && C > 0 && H > 0 && W > 0 && B <= 65535, "ach_avgpool3: bad args");
    const long long n = (long long)C * H * ((W + 3) / 4);
    ACH_REQUIRE(n < (1LL << 31), "ach_avgpool3: plane too large for 32-bit indexing");
    avgpool3_kernel<<<dim3(cdiv(n, 256), B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, out, out_bs, C, H, W);
    return check_launch("ach_avgpool3");
}

extern "C" int ach_avgpool3_cl(const float* x, long long x_bs, float* out, long long out_bs, int B, int C, int H, int W,
                               void* stream) {
    ACH_REQUIRE(x && out && B > 0 && C > 0 && H > 0 && W > 0 && B <= 65535, "ach_avgpool3_cl: bad args");
    ACH_REQUIRE(aligned16(out) && out_bs % 4 == 0, "ach_avgpool3_cl: out must be 16-byte aligned");
    ACH_REQUIRE((long long)H * W < (1LL << 28), "ach_avgpool3_cl: plane too large for 32-bit indexing");
    avgpool3_cl_kernel<<<dim3(cdiv((long long)H * ((W + 3) / 4), 128), B), 128, 0, (cudaStream_t)stream>>>(x, x_bs, out, out_bs, C, H, W);
    return check_launch("ach_avgpool3_cl");
}

extern "C" int ach_fc(const float* x, long long x_bs, const float* w, const float* scale, const float* bias, float* out,
                      long long out_bs, int B, int K, int O, int act, void* stream) {
    ACH_REQUIRE(x && w && out && B > 0 && K > 0 && O > 0 && B <= 65535, "ach_fc: bad args");
    fc_kernel<<<dim3(cdiv(O, 8), B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, w, scale, bias, out, out_bs, K, O, act);
    return check_launch("ach_fc");
}

extern "C" int ach_logsoftmax_t(const float* x, long long x_bs, float* out, long long out_bs, int B, int K, int N, void* stream) {
    ACH_REQUIRE(x && out && B > 0 && K > 0 && K <= 32 && N > 0 && B <= 65535, "ach_logsoftmax_t: bad args (K=%d must be <= 32)", K);
    logsoftmax_t_kernel<<<dim3(cdiv(N, 256), B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, out, out_bs, K, N);
    return check_launch("ach_logsoftmax_t");
}

extern "C" int ach_copy_add(const float* x, long long x_bs, const float* post, float* out, long long out_bs, int B, int C,
                            int P, void* stream) {
    ACH_REQUIRE(x && out && B > 0 && C > 0 && P > 0 && B <= 65535, "ach_copy_add: bad args");
    const long long n = (long long)C * P;
    copy_add_kernel<<<dim3(cdiv(n, 256), B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, post, out, out_bs, n);
    return check_launch("ach_copy_add");
}

extern "C" int ach_add(const float* a, long long a_bs, const float* b2, long long b_bs, float* out, long long out_bs, int B,
                       int C, int P, void* stream) {
    ACH_REQUIRE(a && b2 && out && B > 0 && C > 0 && P > 0 && B <= 65535, "ach_add: bad args");
    const long long n = (long long)C * P;
    add_kernel<<<dim3(cdiv(n, 256), B), 256, 0, (cudaStream_t)stream>>>(a, a_bs, b2, b_bs, out, out_bs, n);
    return check_launch("ach_add");
}

extern "C" int ach_fill(float* x, long long n, float value, void* stream) {
    ACH_REQUIRE(x && n > 0, "ach_fill: bad args");
    fill_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, value);
    return check_launch("ach_fill");
}
