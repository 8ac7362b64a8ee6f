// EdgeViT building blocks (backbone/vision/edgevit_modules/edgevit.py) that the existing
// kernels do not cover:
//   ach_subsample   GlobalSparseAttn's sampler nn.AvgPool2d(1, sr) (:66,76): kernel 1, stride sr = every sr-th pixel
//   ach_mhsa        token self-attention softmax(q k^T * scale) v per (frame, head) (:81-87), qkv channel-major
//                   (B, 3*heads*d, N) as the fused LN+qkv GEMM produces it; one CTA = 128 queries of one head, K/V of the
//                   head in shared memory, online softmax in the log2 domain (one MUFU.EX2 per key), thread = query
//   ach_dw_convT    LocalProp = depthwise ConvTranspose2d with kernel = stride = sr (:68,91): every output pixel has
//                   exactly one source pixel: out[c, y, x] = in[c, y/sr, x/sr] * w[c, y%sr, x%sr] + bias[c]
#include <algorithm>

#include "common.cuh"

namespace ach {

__global__ void __launch_bounds__(256) subsample_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ out,
                                                        long long out_bs, int H, int W, int ho, int wo, int sr) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= ho * wo) return;
    const int y = i / wo, xx = i - y * wo;
    const int c = blockIdx.y;
    out[(long long)blockIdx.z * out_bs + (long long)c * ho * wo + i] =
        __ldg(x + (long long)blockIdx.z * x_bs + ((long long)c * H + y * sr) * W + xx * sr);
}

__global__ void __launch_bounds__(256) dw_convT_kernel(const float* __restrict__ x, long long x_bs, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ out, long long out_bs,
                                                       int h, int ww, int sr) {
    const int H = h * sr, W = ww * sr;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= H * W) return;
    const int y = i / W, xx = i - y * W;
    const int c = blockIdx.y;
    const float v = __ldg(x + (long long)blockIdx.z * x_bs + ((long long)c * h + y / sr) * ww + xx / sr);
    const float wt = __ldg(w + (c * sr + y % sr) * sr + xx % sr);
    out[(long long)blockIdx.z * out_bs + (long long)c * H * W + i] = fmaf(v, wt, bias ? __ldg(bias + c) : 0.f);
}

// patch x patch space-to-depth: the im2col of a conv with kernel = stride = patch (PatchEmbed.proj, edgevit.py:184), which then
// runs as a pointwise GEMM over K = C * patch^2 on tcgen05.  Thread = one input row segment of `patch` pixels.
__global__ void __launch_bounds__(256) s2d_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ out, long long out_bs,
                                                  int H, int W, int patch) {
    const int ho = H / patch, wo = W / patch;
    const int i = blockIdx.x * 256 + threadIdx.x;          // (ky, oy, ox): consecutive threads -> consecutive ox
    if (i >= patch * ho * wo) return;
    const int ox = i % wo, oy = (i / wo) % ho, ky = i / (wo * ho);
    const int c = blockIdx.y;
    const float* __restrict__ src = x + (long long)blockIdx.z * x_bs + ((long long)c * H + oy * patch + ky) * W + ox * patch;
    float* __restrict__ dst = out + (long long)blockIdx.z * out_bs + ((long long)(c * patch + ky) * patch) * ho * wo + oy * wo + ox;
    for (int kx = 0; kx < patch; ++kx) dst[(long long)kx * ho * wo] = __ldg(src + kx);
}

template <int D>
__global__ void __launch_bounds__(128) mhsa_kernel(const float* __restrict__ qkv, long long qkv_bs, float* __restrict__ out,
                                                   long long out_bs, int heads, int d, int N, float scale2) {
    extern __shared__ float smem[];
    float* ks = smem;            // [N][D] (zero padded past d)
    float* vs = smem + N * D;    // [N][D]
    const int head = blockIdx.x, qt = blockIdx.y, b = blockIdx.z;
    const int inner = heads * d;
    const float* qb = qkv + (long long)b * qkv_bs + (long long)(head * d) * N;
    const float* kb = qb + (long long)inner * N;
    const float* vb = kb + (long long)inner * N;
    for (int i = threadIdx.x; i < N * D; i += 128) {
        const int dd = i / N, n = i - dd * N;           // consecutive threads -> consecutive tokens of one channel
        ks[n * D + dd] = dd < d ? __ldg(kb + (long long)dd * N + n) : 0.f;
        vs[n * D + dd] = dd < d ? __ldg(vb + (long long)dd * N + n) : 0.f;
    }
    __syncthreads();
    const int n = qt * 128 + threadIdx.x;
    if (n >= N) return;
    float q[D], acc[D];
#pragma unroll
    for (int dd = 0; dd < D; ++dd) {
        q[dd] = dd < d ? __ldg(qb + (long long)dd * N + n) * scale2 : 0.f;     // log2 units: exp(x) = 2^(x * log2 e)
        acc[dd] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < N; ++j) {
        const float4* kj = reinterpret_cast<const float4*>(ks + j * D);
        float s = 0.f;
#pragma unroll
        for (int dd = 0; dd < D / 4; ++dd) {
            const float4 kk = kj[dd];
            s = fmaf(q[4 * dd], kk.x, s);
            s = fmaf(q[4 * dd + 1], kk.y, s);
            s = fmaf(q[4 * dd + 2], kk.z, s);
            s = fmaf(q[4 * dd + 3], kk.w, s);
        }
        if (s > m) {                      // rescale only when the running max moves
            const float c = ex2_approx(m - s);
            l *= c;
#pragma unroll
            for (int dd = 0; dd < D; ++dd) acc[dd] *= c;
            m = s;
        }
        const float e = ex2_approx(s - m);
        l += e;
        const float4* vj = reinterpret_cast<const float4*>(vs + j * D);
#pragma unroll
        for (int dd = 0; dd < D / 4; ++dd) {
            const float4 vv = vj[dd];
            acc[4 * dd] = fmaf(e, vv.x, acc[4 * dd]);
            acc[4 * dd + 1] = fmaf(e, vv.y, acc[4 * dd + 1]);
            acc[4 * dd + 2] = fmaf(e, vv.z, acc[4 * dd + 2]);
            acc[4 * dd + 3] = fmaf(e, vv.w, acc[4 * dd + 3]);
        }
    }
    const float inv = 1.0f / l;
    float* ob = out + (long long)b * out_bs + (long long)(head * d) * N + n;
#pragma unroll
    for (int dd = 0; dd < D; ++dd)
        if (dd < d) ob[(long long)dd * N] = acc[dd] * inv;
}

template <int D>
static int launch_mhsa(const float* qkv, long long qkv_bs, float* out, long long out_bs, int B, int heads, int d, int N, float scale,
                       cudaStream_t st) {
    const size_t smem = (size_t)2 * N * D * sizeof(float);
    ACH_REQUIRE(smem <= 200 * 1024, "ach_mhsa: %d tokens x %d dims do not fit shared memory", N, D);
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(mhsa_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    mhsa_kernel<D><<<dim3(heads, cdiv(N, 128), B), 128, smem, st>>>(qkv, qkv_bs, out, out_bs, heads, d, N, scale * 1.4426950408889634f);
    return check_launch("ach_mhsa");
}

}  // namespace ach

extern "C" int ach_subsample(const float* x, long long x_bs, float* out, long long out_bs, int B, int C, int H, int W, int sr, void* stream) {
    using namespace ach;
    ACH_REQUIRE(x && out && B > 0 && B <= 65535 && C > 0 && C <= 65535 && H > 0 && W > 0 && sr >= 1, "ach_subsample: bad args");
    const int ho = (H - 1) / sr + 1, wo = (W - 1) / sr + 1;
    subsample_kernel<<<dim3(cdiv((long long)ho * wo, 256), C, B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, out, out_bs, H, W, ho, wo, sr);
    return check_launch("ach_subsample");
}

extern "C" int ach_s2d(const float* x, long long x_bs, float* out, long long out_bs, int B, int C, int H, int W, int patch, void* stream) {
    using namespace ach;
    ACH_REQUIRE(x && out && B > 0 && B <= 65535 && C > 0 && C <= 65535 && patch >= 1 && H % patch == 0 && W % patch == 0, "ach_s2d: bad args");
    s2d_kernel<<<dim3(cdiv((long long)H * W / patch, 256), C, B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, out, out_bs, H, W, patch);
    return check_launch("ach_s2d");
}

extern "C" int ach_dw_convT(const float* x, long long x_bs, const float* w, const float* bias, float* out, long long out_bs, int B, int C,
                            int h, int w_in, int sr, void* stream) {
    using namespace ach;
    ACH_REQUIRE(x && w && out && B > 0 && B <= 65535 && C > 0 && C <= 65535 && h > 0 && w_in > 0 && sr >= 1, "ach_dw_convT: bad args");
    dw_convT_kernel<<<dim3(cdiv((long long)h * sr * w_in * sr, 256), C, B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, w, bias, out, out_bs, h,
                                                                                                          w_in, sr);
    return check_launch("ach_dw_convT");
}

extern "C" int ach_mhsa(const float* qkv, long long qkv_bs, float* out, long long out_bs, int B, int heads, int dim_head, int N, float scale,
                        void* stream) {
    using namespace ach;
    ACH_REQUIRE(qkv && out && B > 0 && B <= 65535 && heads > 0 && dim_head > 0 && N > 0, "ach_mhsa: bad args");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dim_head <= 24) return launch_mhsa<24>(qkv, qkv_bs, out, out_bs, B, heads, dim_head, N, scale, st);
    if (dim_head <= 32) return launch_mhsa<32>(qkv, qkv_bs, out, out_bs, B, heads, dim_head, N, scale, st);
    if (dim_head <= 48) return launch_mhsa<48>(qkv, qkv_bs, out, out_bs, B, heads, dim_head, N, scale, st);
    set_error("ach_mhsa: dim_head=%d not instantiated (<= 48)", dim_head);
    return ACH_ERR_INVALID;
}

// ------------------------------------------------------------------------------------------------
// EfficientFormerV2 ("ImageEncoder", backbone='ef') attention: Attention4D (ImageEncoder.py:131-160) and
// Attention4DDownsample (:267-289).  All heads of a query tile are processed together because the talking-head 1x1
// convolutions mix the heads of every (query, key) score before and after the softmax:
//   S[h][q][k]  = scale * sum_c q[h*kd + c][q] k[h*kd + c][k] + ab[h][q][k]
//   S           = th1 S + b1           (optional)        P = softmax_k(S)        P = th2 P + b2   (optional)
//   out[h*d + j][q] = sum_k P[h][q][k] v[h*d + j][k]  (+ add[h*d + j][q], then GELU if `gelu`)
// One CTA per (frame, tile of TQ queries); the (heads x TQ x Nk) score block lives in shared memory.
namespace ach {

constexpr int EFA_TQ = 4;      // queries per CTA
constexpr int EFA_MAXH = 8;

__global__ void __launch_bounds__(256) ef_attention_kernel(const float* __restrict__ q, long long q_bs, const float* __restrict__ k,
                                                           long long k_bs, const float* __restrict__ v, long long v_bs,
                                                           const float* __restrict__ ab, const float* __restrict__ th1,
                                                           const float* __restrict__ th2, const float* __restrict__ add, long long add_bs,
                                                           float* __restrict__ out, long long out_bs, int heads, int kd, int d, int Nq,
                                                           int Nk, float scale, int gelu) {
    extern __shared__ float smem[];
    float* S = smem;                              // [heads][TQ][Nk]
    float* qs = S + heads * EFA_TQ * Nk;          // [heads*kd][TQ]
    float* ths = qs + heads * kd * EFA_TQ;        // th1 (h*h + h) | th2 (h*h + h)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * EFA_TQ, b = blockIdx.y;
    const float* qb = q + (long long)b * q_bs;
    const float* kb = k + (long long)b * k_bs;
    const float* vb = v + (long long)b * v_bs;
    for (int i = tid; i < heads * kd * EFA_TQ; i += 256) {
        const int c = i / EFA_TQ, t = i - c * EFA_TQ;
        qs[i] = (q0 + t < Nq) ? __ldg(qb + (long long)c * Nq + q0 + t) * scale : 0.f;
    }
    const int nth = heads * heads + heads;
    for (int i = tid; i < 2 * nth; i += 256) ths[i] = (i < nth) ? (th1 ? th1[i] : 0.f) : (th2 ? th2[i - nth] : 0.f);
    __syncthreads();
    // ---- scores: thread per (head, key), TQ accumulators
    for (int i = tid; i < heads * Nk; i += 256) {
        const int h = i / Nk, kk = i - h * Nk;
        float acc[EFA_TQ];
#pragma unroll
        for (int t = 0; t < EFA_TQ; ++t) acc[t] = 0.f;
        for (int c = 0; c < kd; ++c) {
            const float kv = __ldg(kb + (long long)(h * kd + c) * Nk + kk);
            const float* qc = qs + (h * kd + c) * EFA_TQ;
#pragma unroll
            for (int t = 0; t < EFA_TQ; ++t) acc[t] = fmaf(qc[t], kv, acc[t]);
        }
#pragma unroll
        for (int t = 0; t < EFA_TQ; ++t) {
            const int qq = min(q0 + t, Nq - 1);
            S[(h * EFA_TQ + t) * Nk + kk] = acc[t] + __ldg(ab + ((long long)h * Nq + qq) * Nk + kk);
        }
    }
    __syncthreads();
    // ---- talking head 1 (mixes the heads of every (query, key) score)
    if (th1) {
        for (int i = tid; i < EFA_TQ * Nk; i += 256) {
            float s[EFA_MAXH], o[EFA_MAXH];
            for (int h = 0; h < heads; ++h) s[h] = S[h * EFA_TQ * Nk + i];
            for (int ho = 0; ho < heads; ++ho) {
                float a = ths[heads * heads + ho];
                for (int h = 0; h < heads; ++h) a = fmaf(ths[ho * heads + h], s[h], a);
                o[ho] = a;
            }
            for (int h = 0; h < heads; ++h) S[h * EFA_TQ * Nk + i] = o[h];
        }
        __syncthreads();
    }
    // ---- softmax over the keys: one warp per (head, query) row
    for (int r = warp; r < heads * EFA_TQ; r += 8) {
        float* row = S + r * Nk;
        float m = -INFINITY;
        for (int j = lane; j < Nk; j += 32) m = fmaxf(m, row[j]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < Nk; j += 32) {
            const float e = expf(row[j] - m);
            row[j] = e;
            s += e;
        }
        s = warp_sum(s);
        const float inv = 1.0f / s;
        for (int j = lane; j < Nk; j += 32) row[j] *= inv;
    }
    __syncthreads();
    if (th2) {
        const float* t2 = ths + nth;
        for (int i = tid; i < EFA_TQ * Nk; i += 256) {
            float s[EFA_MAXH], o[EFA_MAXH];
            for (int h = 0; h < heads; ++h) s[h] = S[h * EFA_TQ * Nk + i];
            for (int ho = 0; ho < heads; ++ho) {
                float a = t2[heads * heads + ho];
                for (int h = 0; h < heads; ++h) a = fmaf(t2[ho * heads + h], s[h], a);
                o[ho] = a;
            }
            for (int h = 0; h < heads; ++h) S[h * EFA_TQ * Nk + i] = o[h];
        }
        __syncthreads();
    }
    // ---- out = P v: one warp per value channel, lanes over the keys
    for (int r = warp; r < heads * d; r += 8) {
        const int h = r / d;
        const float* vr = vb + (long long)r * Nk;
        float acc[EFA_TQ];
#pragma unroll
        for (int t = 0; t < EFA_TQ; ++t) acc[t] = 0.f;
        for (int j = lane; j < Nk; j += 32) {
            const float vv = __ldg(vr + j);
#pragma unroll
            for (int t = 0; t < EFA_TQ; ++t) acc[t] = fmaf(S[(h * EFA_TQ + t) * Nk + j], vv, acc[t]);
        }
#pragma unroll
        for (int t = 0; t < EFA_TQ; ++t) acc[t] = warp_sum(acc[t]);
        if (lane < EFA_TQ && q0 + lane < Nq) {
            float y = acc[0];
#pragma unroll
            for (int t = 1; t < EFA_TQ; ++t) y = (lane == t) ? acc[t] : y;
            if (add) y += __ldg(add + (long long)b * add_bs + (long long)r * Nq + q0 + lane);
            if (gelu) y = apply_act(y, ACT_GELU);
            out[(long long)b * out_bs + (long long)r * Nq + q0 + lane] = y;
        }
    }
}

template <int TQ>
__global__ void __launch_bounds__(256) ef_attention_v2_kernel(const float* __restrict__ q, long long q_bs, const float* __restrict__ k,
                                                           long long k_bs, const float* __restrict__ v, long long v_bs,
                                                           const float* __restrict__ ab, const float* __restrict__ th1,
                                                           const float* __restrict__ th2, const float* __restrict__ add, long long add_bs,
                                                           float* __restrict__ out, long long out_bs, int heads, int kd, int d, int Nq,
                                                           int Nk, float scale, int gelu) {
    extern __shared__ float smem[];
    float* S = smem;                              // [heads][TQ][Nk]
    float* ths = S + heads * TQ * Nk;         // th1 (h*h + h) | th2 (h*h + h)
    float* qs = ths + 2 * (heads * heads + heads);
    qs += (4 - ((qs - smem) & 3)) & 3;        // [heads*kd][TQ]; the same (16-byte aligned) region later holds V_h^T: q is dead after the scores
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * TQ, b = blockIdx.y;
    const float* qb = q + (long long)b * q_bs;
    const float* kb = k + (long long)b * k_bs;
    const float* vb = v + (long long)b * v_bs;
    for (int i = tid; i < heads * kd * TQ; i += 256) {
        const int c = i / TQ, t = i - c * TQ;
        qs[i] = (q0 + t < Nq) ? __ldg(qb + (long long)c * Nq + q0 + t) * scale : 0.f;
    }
    const int nth = heads * heads + heads;
    for (int i = tid; i < 2 * nth; i += 256) ths[i] = (i < nth) ? (th1 ? th1[i] : 0.f) : (th2 ? th2[i - nth] : 0.f);
    __syncthreads();
    // ---- scores: thread per (head, key), TQ accumulators
    for (int i = tid; i < heads * Nk; i += 256) {
        const int h = i / Nk, kk = i - h * Nk;
        float acc[TQ];
#pragma unroll
        for (int t = 0; t < TQ; ++t) acc[t] = 0.f;
        for (int c = 0; c < kd; ++c) {
            const float kv = __ldg(kb + (long long)(h * kd + c) * Nk + kk);
            const float* qc = qs + (h * kd + c) * TQ;
#pragma unroll
            for (int t = 0; t < TQ; ++t) acc[t] = fmaf(qc[t], kv, acc[t]);
        }
#pragma unroll
        for (int t = 0; t < TQ; ++t) {
            const int qq = min(q0 + t, Nq - 1);
            S[(h * TQ + t) * Nk + kk] = acc[t] + __ldg(ab + ((long long)h * Nq + qq) * Nk + kk);
        }
    }
    __syncthreads();
    // ---- talking head 1 (mixes the heads of every (query, key) score)
    if (th1) {
        for (int i = tid; i < TQ * Nk; i += 256) {
            float s[EFA_MAXH], o[EFA_MAXH];
            for (int h = 0; h < heads; ++h) s[h] = S[h * TQ * Nk + i];
            for (int ho = 0; ho < heads; ++ho) {
                float a = ths[heads * heads + ho];
                for (int h = 0; h < heads; ++h) a = fmaf(ths[ho * heads + h], s[h], a);
                o[ho] = a;
            }
            for (int h = 0; h < heads; ++h) S[h * TQ * Nk + i] = o[h];
        }
        __syncthreads();
    }
    // ---- softmax over the keys: one warp per (head, query) row
    for (int r = warp; r < heads * TQ; r += 8) {
        float* row = S + r * Nk;
        float m = -INFINITY;
        for (int j = lane; j < Nk; j += 32) m = fmaxf(m, row[j]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < Nk; j += 32) {
            const float e = expf(row[j] - m);
            row[j] = e;
            s += e;
        }
        s = warp_sum(s);
        const float inv = 1.0f / s;
        for (int j = lane; j < Nk; j += 32) row[j] *= inv;
    }
    __syncthreads();
    if (th2) {
        const float* t2 = ths + nth;
        for (int i = tid; i < TQ * Nk; i += 256) {
            float s[EFA_MAXH], o[EFA_MAXH];
            for (int h = 0; h < heads; ++h) s[h] = S[h * TQ * Nk + i];
            for (int ho = 0; ho < heads; ++ho) {
                float a = t2[heads * heads + ho];
                for (int h = 0; h < heads; ++h) a = fmaf(t2[ho * heads + h], s[h], a);
                o[ho] = a;
            }
            for (int h = 0; h < heads; ++h) S[h * TQ * Nk + i] = o[h];
        }
        __syncthreads();
    }
    // ---- out = P v, one head at a time: V_h^T staged in shared memory ([key][d], 16-byte aligned rows), every thread
    // a 4 (value channels) x 4 (queries) register tile - per key one LDS.128 of V and four broadcast loads of P feed 16 FMAs
    // (v1 reduced every output over a warp: 20 shuffles per 4 outputs and V re-read from L2 by every 4-query CTA)
    float* vs = qs;                               // q tile is dead (several barriers ago): 105 KB per CTA -> two CTAs per SM
    const int vp = d + 4;                         // row pitch (floats)
    const int jt = tid % (d / 4), tt = tid / (d / 4);
    const bool active = tt < TQ / 4;
    for (int h = 0; h < heads; ++h) {
        __syncthreads();
        for (int i = tid; i < d * Nk; i += 256) {
            const int r = i / Nk, kk = i - r * Nk;
            vs[kk * vp + r] = __ldg(vb + (long long)(h * d + r) * Nk + kk);
        }
        __syncthreads();
        if (!active) continue;
        float acc[4][4];
#pragma unroll
        for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[a_][t] = 0.f;
        const float* sp = S + (h * TQ + 4 * tt) * Nk;
        for (int kk = 0; kk < Nk; ++kk) {
            const float4 vv = *reinterpret_cast<const float4*>(vs + kk * vp + 4 * jt);
            const float p0 = sp[kk], p1 = sp[Nk + kk], p2 = sp[2 * Nk + kk], p3 = sp[3 * Nk + kk];
            acc[0][0] = fmaf(vv.x, p0, acc[0][0]); acc[0][1] = fmaf(vv.x, p1, acc[0][1]); acc[0][2] = fmaf(vv.x, p2, acc[0][2]); acc[0][3] = fmaf(vv.x, p3, acc[0][3]);
            acc[1][0] = fmaf(vv.y, p0, acc[1][0]); acc[1][1] = fmaf(vv.y, p1, acc[1][1]); acc[1][2] = fmaf(vv.y, p2, acc[1][2]); acc[1][3] = fmaf(vv.y, p3, acc[1][3]);
            acc[2][0] = fmaf(vv.z, p0, acc[2][0]); acc[2][1] = fmaf(vv.z, p1, acc[2][1]); acc[2][2] = fmaf(vv.z, p2, acc[2][2]); acc[2][3] = fmaf(vv.z, p3, acc[2][3]);
            acc[3][0] = fmaf(vv.w, p0, acc[3][0]); acc[3][1] = fmaf(vv.w, p1, acc[3][1]); acc[3][2] = fmaf(vv.w, p2, acc[3][2]); acc[3][3] = fmaf(vv.w, p3, acc[3][3]);
        }
#pragma unroll
        for (int a_ = 0; a_ < 4; ++a_) {
            const int r = h * d + 4 * jt + a_;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int qq = q0 + 4 * tt + t;
                if (qq < Nq) {
                    float y = acc[a_][t];
                    if (add) y += __ldg(add + (long long)b * add_bs + (long long)r * Nq + qq);
                    if (gelu) y = apply_act(y, ACT_GELU);
                    out[(long long)b * out_bs + (long long)r * Nq + qq] = y;
                }
            }
        }
    }
}

// bilinear x2, align_corners=False (nn.Upsample(scale_factor=2, mode='bilinear'), ImageEncoder.py:80), optional GELU
__global__ void __launch_bounds__(256) upsample2x_hp_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ out,
                                                            long long out_bs, int H, int W, int gelu) {
    const int Ho = 2 * H, Wo = 2 * W;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= Ho * Wo) return;
    const int oy = i / Wo, ox = i - oy * Wo;
    const int c = blockIdx.y;
    // ATen area_pixel_compute_source_index (align_corners=False): src = max((dst + 0.5) * 0.5 - 0.5, 0)
    const float fy = fmaxf(((float)oy + 0.5f) * 0.5f - 0.5f, 0.f), fx = fmaxf(((float)ox + 0.5f) * 0.5f - 0.5f, 0.f);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < H - 1), x1 = x0 + (x0 < W - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float* xp = x + (long long)blockIdx.z * x_bs + (long long)c * H * W;
    float v = hy * (hx * __ldg(xp + y0 * W + x0) + lx * __ldg(xp + y0 * W + x1)) + ly * (hx * __ldg(xp + y1 * W + x0) + lx * __ldg(xp + y1 * W + x1));
    if (gelu) v = apply_act(v, ACT_GELU);
    out[(long long)blockIdx.z * out_bs + (long long)c * Ho * Wo + i] = v;
}

}  // namespace ach

extern "C" int ach_ef_attention(const float* q, long long q_bs, const float* k, long long k_bs, const float* v, long long v_bs,
                                const float* ab, const float* th1, const float* th2, const float* add, long long add_bs, float* out,
                                long long out_bs, int B, int heads, int key_dim, int d, int Nq, int Nk, float scale, int gelu, void* stream) {
    using namespace ach;
    ACH_REQUIRE(q && k && v && ab && out, "ach_ef_attention: null arg");
    ACH_REQUIRE(B > 0 && B <= 65535 && heads > 0 && heads <= EFA_MAXH && key_dim > 0 && d > 0 && Nq > 0 && Nk > 0, "ach_ef_attention: bad dims");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // v2 (register-tiled P.V with V_h^T in shared memory) when the tiles fit; v1 (4 queries per CTA, warp reductions) otherwise
    for (int tq : {16}) {   // measured: the 8-query variant (Nk = 400, d = 64) keeps only 32 threads busy in P.V and loses to v1 (1.36 vs 0.54 ms)
        if ((tq == 16 && Nk > 128) || d % 4 != 0 || 256 / (d / 4) < tq / 4) continue;
        const size_t uni = (size_t)std::max(heads * key_dim * tq, Nk * (d + 4));
        const size_t smem2 = ((size_t)heads * tq * Nk + 2 * (heads * heads + heads) + 4 + uni) * sizeof(float);
        if (smem2 > 220 * 1024) continue;
        static PerDeviceOnce attr2_once;
        if (attr2_once.first()) {
            cudaFuncSetAttribute(ef_attention_v2_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
            cudaFuncSetAttribute(ef_attention_v2_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        }
        if (tq == 16)
            ef_attention_v2_kernel<16><<<dim3(cdiv(Nq, 16), B), 256, smem2, st>>>(q, q_bs, k, k_bs, v, v_bs, ab, th1, th2, add, add_bs, out, out_bs,
                                                                                 heads, key_dim, d, Nq, Nk, scale, gelu);
        else
            ef_attention_v2_kernel<8><<<dim3(cdiv(Nq, 8), B), 256, smem2, st>>>(q, q_bs, k, k_bs, v, v_bs, ab, th1, th2, add, add_bs, out, out_bs,
                                                                               heads, key_dim, d, Nq, Nk, scale, gelu);
        return check_launch("ach_ef_attention");
    }
    const size_t smem = (size_t)(heads * EFA_TQ * Nk + heads * key_dim * EFA_TQ + 2 * (heads * heads + heads)) * sizeof(float);
    ACH_REQUIRE(smem <= 160 * 1024, "ach_ef_attention: %d keys do not fit shared memory", Nk);
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(ef_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    }
    ef_attention_kernel<<<dim3(cdiv(Nq, EFA_TQ), B), 256, smem, st>>>(q, q_bs, k, k_bs, v, v_bs, ab, th1, th2, add, add_bs, out,
                                                                      out_bs, heads, key_dim, d, Nq, Nk, scale, gelu);
    return check_launch("ach_ef_attention");
}

extern "C" int ach_upsample2x_hp(const float* x, long long x_bs, float* out, long long out_bs, int B, int C, int H, int W, int gelu,
                                 void* stream) {
    using namespace ach;
    ACH_REQUIRE(x && out && B > 0 && B <= 65535 && C > 0 && C <= 65535 && H > 0 && W > 0, "ach_upsample2x_hp: bad args");
    upsample2x_hp_kernel<<<dim3(cdiv((long long)4 * H * W, 256), C, B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, out, out_bs, H, W, gelu);
    return check_launch("ach_upsample2x_hp");
}
