// XCA (cross-covariance attention, EdgeNeXt SDTA encoder, sdta_encoder.py:162-185) core.
// qkv arrives channel-major (B, 3C, N) straight from the fused LN+qkv GEMM, which is already the
// (B, h, d, N) layout the reference reaches through three permutes.  One CTA per (head, frame):
//   1. L2 norms of the d q-rows and d k-rows over the N tokens        (warp per row, shuffle reduce)
//   2. Gram G = q k^T (d x d), N streamed through shared memory in 64-token chunks
//   3. attn = softmax_j(G_ij / (|q_i| |k_j|) * temperature[h])         (warp per row)
//   4. the attention matrix is folded into the output projection:
//        wt_eff[b][h*d + j][o] = sum_i proj_wt[h*d + i][o] * attn[i][j]
//      so "attn @ v" and "proj" become ONE per-frame-weight pointwise GEMM over v (ach_pw_conv) and the
//      (B, h, d, N) attention output never exists in memory.
#include "common.cuh"

namespace ach {

constexpr int XCA_MAX_D = 64;
constexpr int XCA_CHUNK = 64;
constexpr int XCA_MAX_PAIRS = (XCA_MAX_D * XCA_MAX_D + 255) / 256;  // Gram entries per thread (a multiple of 4: 2 x 2 register tiles)
static_assert(XCA_MAX_PAIRS % 4 == 0, "2 x 2 Gram tiles");

__global__ void __launch_bounds__(256) xca_fold_kernel(const float* __restrict__ qkv, long long qkv_bs,
                                                       const float* __restrict__ temperature,
                                                       const float* __restrict__ proj_wt, int ldw, float* __restrict__ wt_eff,
                                                       long long wt_eff_bs, int C, int heads, int N) {
    extern __shared__ float smem[];
    const int d = C / heads;
    float* qs = smem;                             // [d][XCA_CHUNK + 1]
    float* ks = qs + d * (XCA_CHUNK + 1);         // [d][XCA_CHUNK + 1]
    float* nrm = ks + d * (XCA_CHUNK + 1);        // [2d]
    float* attn = nrm + 2 * d;                    // [d][d]
    float* pw = attn + d * d;                     // [d][ldw]: this head's rows of the projection weight

    const int h = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* q = qkv + (long long)b * qkv_bs + (long long)(h * d) * N;
    const float* k = q + (long long)C * N;

    // this head's projection rows -> shared memory (all loads in flight at once; step 4 used to read them from L2 inside a
    // d-long dependent FMA loop: ~1 300 L2 round trips per thread and half of the kernel's 92 us at C = 176)
    for (int i = tid; i < d * ldw; i += 256) pw[i] = __ldg(proj_wt + (long long)(h * d) * ldw + i);

    // 1. row norms (F.normalize: x / max(||x||, 1e-12))
    for (int r = warp; r < 2 * d; r += 8) {
        const float* row = (r < d) ? q + (long long)r * N : k + (long long)(r - d) * N;
        float s = 0.f;
        for (int n = lane; n < N; n += 32) {
            const float v = row[n];
            s = fmaf(v, v, s);
        }
        s = warp_sum(s);
        if (lane == 0) nrm[r] = fmaxf(sqrtf(s), 1e-12f);
    }

    // 2. Gram matrix.  Even d: 2 x 2 register tiles (rows 2ti, 2ti+1 of q against rows 2tj, 2tj+1 of k: 4 shared loads per 4 FMAs; the
    // one-entry-per-thread form issued 2 loads per FMA and ncu showed the kernel LSU-bound, 52 % LSU / 60 % l1tex);
    // odd d: entries (i, j) = e / d, e % d round-robin.  Either way every entry is the same sequential sum over the tokens.
    float g[XCA_MAX_PAIRS];
#pragma unroll
    for (int e = 0; e < XCA_MAX_PAIRS; ++e) g[e] = 0.f;
    const int npairs = d * d;
    const int dh = d >> 1, ntiles = dh * dh;
    const bool tiled = (d & 1) == 0 && ntiles >= 128;   // small heads (d = 12: 36 tiles) keep one entry per thread - measured 0.066 vs 0.086 ms
    for (int n0 = 0; n0 < N; n0 += XCA_CHUNK) {
        const int nn = min(XCA_CHUNK, N - n0);
        __syncthreads();
        for (int i = tid; i < d * XCA_CHUNK; i += 256) {
            const int r = i / XCA_CHUNK, c = i - r * XCA_CHUNK;
            const bool ok = c < nn;
            qs[r * (XCA_CHUNK + 1) + c] = ok ? q[(long long)r * N + n0 + c] : 0.f;
            ks[r * (XCA_CHUNK + 1) + c] = ok ? k[(long long)r * N + n0 + c] : 0.f;
        }
        __syncthreads();
        if (tiled) {
#pragma unroll
            for (int e = 0; e < XCA_MAX_PAIRS / 4; ++e) {
                const int idx = tid + e * 256;
                if (idx < ntiles) {
                    const int ti = idx / dh, tj = idx - ti * dh;
                    const float* q0 = qs + (2 * ti) * (XCA_CHUNK + 1);
                    const float* k0 = ks + (2 * tj) * (XCA_CHUNK + 1);
                    float s00 = g[4 * e], s01 = g[4 * e + 1], s10 = g[4 * e + 2], s11 = g[4 * e + 3];
#pragma unroll 16
                    for (int c = 0; c < XCA_CHUNK; ++c) {
                        const float a0 = q0[c], a1 = q0[XCA_CHUNK + 1 + c], b0 = k0[c], b1 = k0[XCA_CHUNK + 1 + c];
                        s00 = fmaf(a0, b0, s00);
                        s01 = fmaf(a0, b1, s01);
                        s10 = fmaf(a1, b0, s10);
                        s11 = fmaf(a1, b1, s11);
                    }
                    g[4 * e] = s00, g[4 * e + 1] = s01, g[4 * e + 2] = s10, g[4 * e + 3] = s11;
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < XCA_MAX_PAIRS; ++e) {
                const int idx = tid + e * 256;
                if (idx < npairs) {
                    const float* qi = qs + (idx / d) * (XCA_CHUNK + 1);
                    const float* kj = ks + (idx % d) * (XCA_CHUNK + 1);
                    float s = g[e];
#pragma unroll 16
                    for (int c = 0; c < XCA_CHUNK; ++c) s = fmaf(qi[c], kj[c], s);
                    g[e] = s;
                }
            }
        }
    }
    const float temp = temperature[h];
    if (tiled) {
#pragma unroll
        for (int e = 0; e < XCA_MAX_PAIRS / 4; ++e) {
            const int idx = tid + e * 256;
            if (idx < ntiles) {
                const int ti = idx / dh, tj = idx - ti * dh;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = 2 * ti + (u >> 1), j = 2 * tj + (u & 1);
                    attn[i * d + j] = g[4 * e + u] / (nrm[i] * nrm[d + j]) * temp;
                }
            }
        }
    } else {
#pragma unroll
        for (int e = 0; e < XCA_MAX_PAIRS; ++e) {
            const int idx = tid + e * 256;
            if (idx < npairs) attn[idx] = g[e] / (nrm[idx / d] * nrm[d + idx % d]) * temp;
        }
    }
    __syncthreads();

    // 3. softmax over j for each row i
    for (int i = warp; i < d; i += 8) {
        float* row = attn + i * d;
        float m = -INFINITY;
        for (int j = lane; j < d; j += 32) m = fmaxf(m, row[j]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < d; j += 32) {
            const float e = expf(row[j] - m);
            row[j] = e;
            s += e;
        }
        s = warp_sum(s);
        for (int j = lane; j < d; j += 32) row[j] = row[j] / s;
    }
    __syncthreads();

    // 4. fold into the projection: rows h*d + j of the per-frame K-major weight
    float* wo = wt_eff + (long long)b * wt_eff_bs;
    if ((d & 3) == 0 && (d >> 2) * ldw >= 256) {   // (small heads keep one output per thread: d = 12 would leave 144 threads busy)
        // one output column o (coalesced) and FOUR consecutive j per thread: a projection value is loaded once for four FMAs, the four
        // attention values are one 16-byte broadcast (2 loads per 4 FMAs instead of 2 per FMA; same sum order over i per output)
        const int dq = d >> 2;
        for (int idx = tid; idx < dq * ldw; idx += 256) {
            const int jq = idx / ldw, o = idx - jq * ldw;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            if (o < C) {
                for (int i = 0; i < d; ++i) {
                    const float wv = pw[i * ldw + o];
                    const float4 a = *reinterpret_cast<const float4*>(attn + i * d + 4 * jq);
                    s0 = fmaf(wv, a.x, s0);
                    s1 = fmaf(wv, a.y, s1);
                    s2 = fmaf(wv, a.z, s2);
                    s3 = fmaf(wv, a.w, s3);
                }
            }
            float* wrow = wo + (long long)(h * d + 4 * jq) * ldw + o;
            wrow[0] = s0, wrow[ldw] = s1, wrow[2 * ldw] = s2, wrow[3 * ldw] = s3;
        }
        return;
    }
    for (int idx = tid; idx < d * ldw; idx += 256) {
        const int j = idx / ldw, o = idx - j * ldw;
        float s = 0.f;
        if (o < C)
            for (int i = 0; i < d; ++i) s = fmaf(pw[i * ldw + o], attn[i * d + j], s);
        wo[(long long)(h * d + j) * ldw + o] = s;
    }
}

}  // namespace ach

extern "C" int ach_xca_fold(const float* qkv, long long qkv_bs, const float* temperature, const float* proj_wt, int ldw,
                            float* wt_eff, long long wt_eff_bs, int B, int C, int heads, int N, void* stream) {
    using namespace ach;
    ACH_REQUIRE(qkv && temperature && proj_wt && wt_eff, "ach_xca_fold: null arg");
    ACH_REQUIRE(B > 0 && B <= 65535 && heads > 0 && C % heads == 0 && N > 0, "ach_xca_fold: bad dims");
    const int d = C / heads;
    ACH_REQUIRE(d <= XCA_MAX_D, "ach_xca_fold: head dim %d > %d", d, XCA_MAX_D);
    ACH_REQUIRE(ldw >= C, "ach_xca_fold: ldw < C");
    const size_t smem = (size_t)(2 * d * (XCA_CHUNK + 1) + 2 * d + d * d + d * ldw) * sizeof(float);
    ACH_REQUIRE(smem <= 160 * 1024, "ach_xca_fold: d=%d, ldw=%d do not fit shared memory", d, ldw);
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(xca_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    }
    xca_fold_kernel<<<dim3(heads, B), 256, smem, (cudaStream_t)stream>>>(qkv, qkv_bs, temperature, proj_wt, ldw, wt_eff, wt_eff_bs,
                                                                        C, heads, N);
    return check_launch("ach_xca_fold");
}
