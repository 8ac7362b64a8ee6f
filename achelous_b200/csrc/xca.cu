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
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ach {

constexpr int XCA_MAX_D = 64;
constexpr int XCA_CHUNK = 64;
constexpr int XCA_ST = XCA_CHUNK + 4;   // row stride of a staged chunk: rows stay 16-byte aligned, 8 consecutive rows hit 8 different 16-byte bank groups
constexpr int XCA_MAX_PAIRS = (XCA_MAX_D * XCA_MAX_D + 255) / 256;  // Gram entries per thread (a multiple of 4: 2 x 2 register tiles)
static_assert(XCA_MAX_PAIRS % 4 == 0, "2 x 2 Gram tiles");

__host__ __device__ inline int xca_r4(int n) { return (n + 3) & ~3; }

// v2 (round 2): ncu on v1 - one CTA per (head, frame), 256 CTAs: issue slots 24 % busy, 5.3 long-scoreboard and 2.3 barrier stalls per
// issue - the token loop staged every 64-token chunk with synchronous loads between two barriers and then read q / k again for the norms.
// Now a CLUSTER of S CTAs shares one (head, frame): CTA r takes chunks r, r + S, ... (double-buffered 16-byte cp.async), accumulates
// partial Gram entries and partial squared norms from the SAME staged chunk (16-byte shared loads: 2 per 4 FMAs), the partials are summed
// over the cluster through distributed shared memory in rank order (every CTA gets bit-identical totals), each CTA computes the small
// softmax redundantly and folds its own column slice of the projection.
template <int S>
__global__ void __launch_bounds__(256) xca_fold_kernel(const float* __restrict__ qkv, long long qkv_bs,
                                                       const float* __restrict__ temperature,
                                                       const float* __restrict__ proj_wt, int ldw, float* __restrict__ wt_eff,
                                                       long long wt_eff_bs, int C, int heads, int N, int ow) {
    extern __shared__ __align__(16) float smem[];
    const int d = C / heads;
    const int npairs = d * d;
    float* stage = smem;                              // [2 buffers][q rows 0..d-1, k rows d..2d-1][XCA_ST]
    float* part = stage + 4 * d * XCA_ST;             // [d*d] partial Gram (row-major), then [2d] partial squared norms
    float* attn = part + xca_r4(npairs + 2 * d);      // [d][d]
    float* nrm = attn + xca_r4(npairs);               // [2d]
    float* pw = nrm + xca_r4(2 * d);                  // [d][ow]: this CTA's column slice of this head's projection rows

    const int r = (S > 1) ? (int)(blockIdx.x % S) : 0, h = (int)blockIdx.x / S, b = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* q = qkv + (long long)b * qkv_bs + (long long)(h * d) * N;
    const float* k = q + (long long)C * N;
    const int o0 = r * ow, cols = max(0, min(ow, ldw - o0));

    const int nchunks = (N + XCA_CHUNK - 1) / XCA_CHUNK;
    auto issue = [&](int ci, int buf) {
        const int n0 = ci * XCA_CHUNK, nn = min(XCA_CHUNK, N - n0);
        float* dst = stage + buf * (2 * d * XCA_ST);
        for (int i = tid; i < 2 * d * (XCA_CHUNK / 4); i += 256) {
            const int row = i / (XCA_CHUNK / 4), c4 = (i - row * (XCA_CHUNK / 4)) * 4;
            const bool ok = c4 < nn;                  // N % 4 == 0: a 16-byte piece is entirely inside or outside
            const float* src = (row < d ? q + (long long)row * N : k + (long long)(row - d) * N) + (ok ? n0 + c4 : 0);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + row * XCA_ST + c4)),
                         "l"(src), "r"(ok ? 16u : 0u)
                         : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (r < nchunks) issue(r, 0);

    // this CTA's columns of the head's projection rows -> shared memory (all loads in flight, under the first chunk's copy)
    for (int i = tid; i < d * cols; i += 256) {
        const int row = i / cols, o = i - row * cols;
        pw[row * ow + o] = __ldg(proj_wt + (long long)(h * d + row) * ldw + o0 + o);
    }

    // 1 + 2. Gram matrix and squared row norms of this CTA's chunks.  Even d >= 16: 2 x 2 register tiles; otherwise one entry per thread.
    float g[XCA_MAX_PAIRS];
#pragma unroll
    for (int e = 0; e < XCA_MAX_PAIRS; ++e) g[e] = 0.f;
    float nacc = 0.f;
    const int dh = d >> 1, ntiles = dh * dh;
    const bool tiled = (d & 1) == 0 && ntiles >= 128;
    const int tn = 255 - tid;                        // norm work goes to the threads the Gram entries use last: (row, half chunk) = (tn / 2, tn % 2)
    const bool has_norm = tn < 4 * d;
    int buf = 0;
    for (int ci = r; ci < nchunks; ci += S, buf ^= 1) {
        if (ci + S < nchunks) {
            issue(ci + S, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const float* sq = stage + buf * (2 * d * XCA_ST);
        const float* sk = sq + d * XCA_ST;
        if (tiled) {
#pragma unroll
            for (int e = 0; e < XCA_MAX_PAIRS / 4; ++e) {
                const int idx = tid + e * 256;
                if (idx < ntiles) {
                    const int ti = idx / dh, tj = idx - ti * dh;
                    const float4* q0 = reinterpret_cast<const float4*>(sq + (2 * ti) * XCA_ST);
                    const float4* k0 = reinterpret_cast<const float4*>(sk + (2 * tj) * XCA_ST);
                    float s00 = g[4 * e], s01 = g[4 * e + 1], s10 = g[4 * e + 2], s11 = g[4 * e + 3];
#pragma unroll 4
                    for (int c = 0; c < XCA_CHUNK / 4; ++c) {
                        const float4 a0 = q0[c], a1 = q0[XCA_ST / 4 + c], b0 = k0[c], b1 = k0[XCA_ST / 4 + c];
                        s00 = fmaf(a0.x, b0.x, s00), s01 = fmaf(a0.x, b1.x, s01), s10 = fmaf(a1.x, b0.x, s10), s11 = fmaf(a1.x, b1.x, s11);
                        s00 = fmaf(a0.y, b0.y, s00), s01 = fmaf(a0.y, b1.y, s01), s10 = fmaf(a1.y, b0.y, s10), s11 = fmaf(a1.y, b1.y, s11);
                        s00 = fmaf(a0.z, b0.z, s00), s01 = fmaf(a0.z, b1.z, s01), s10 = fmaf(a1.z, b0.z, s10), s11 = fmaf(a1.z, b1.z, s11);
                        s00 = fmaf(a0.w, b0.w, s00), s01 = fmaf(a0.w, b1.w, s01), s10 = fmaf(a1.w, b0.w, s10), s11 = fmaf(a1.w, b1.w, s11);
                    }
                    g[4 * e] = s00, g[4 * e + 1] = s01, g[4 * e + 2] = s10, g[4 * e + 3] = s11;
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < XCA_MAX_PAIRS; ++e) {
                const int idx = tid + e * 256;
                if (idx < npairs) {
                    const float4* qi = reinterpret_cast<const float4*>(sq + (idx / d) * XCA_ST);
                    const float4* kj = reinterpret_cast<const float4*>(sk + (idx % d) * XCA_ST);
                    float s = g[e];
#pragma unroll 4
                    for (int c = 0; c < XCA_CHUNK / 4; ++c) {
                        const float4 a = qi[c], bb = kj[c];
                        s = fmaf(a.x, bb.x, s), s = fmaf(a.y, bb.y, s), s = fmaf(a.z, bb.z, s), s = fmaf(a.w, bb.w, s);
                    }
                    g[e] = s;
                }
            }
        }
        if (has_norm) {
            const float4* row = reinterpret_cast<const float4*>(sq + (tn >> 1) * XCA_ST + (tn & 1) * (XCA_CHUNK / 2));
#pragma unroll
            for (int c = 0; c < XCA_CHUNK / 8; ++c) {
                const float4 a = row[c];
                nacc = fmaf(a.x, a.x, nacc), nacc = fmaf(a.y, a.y, nacc), nacc = fmaf(a.z, a.z, nacc), nacc = fmaf(a.w, a.w, nacc);
            }
        }
        __syncthreads();   // the buffer is refilled by the copy issued in the next iteration
    }
    if (tiled) {
#pragma unroll
        for (int e = 0; e < XCA_MAX_PAIRS / 4; ++e) {
            const int idx = tid + e * 256;
            if (idx < ntiles) {
                const int ti = idx / dh, tj = idx - ti * dh;
#pragma unroll
                for (int u = 0; u < 4; ++u) part[(2 * ti + (u >> 1)) * d + 2 * tj + (u & 1)] = g[4 * e + u];
            }
        }
    } else {
#pragma unroll
        for (int e = 0; e < XCA_MAX_PAIRS; ++e) {
            const int idx = tid + e * 256;
            if (idx < npairs) part[idx] = g[e];
        }
    }
    {
        const float other = __shfl_xor_sync(0xffffffffu, nacc, 1);   // the two half-chunk sums of a row sit in neighbouring lanes
        if (has_norm && (tn & 1) == 0) part[npairs + (tn >> 1)] = nacc + other;
    }

    // totals over the cluster, in rank order (identical in every CTA)
    if constexpr (S > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();
        const float* parts[S];
#pragma unroll
        for (int rr = 0; rr < S; ++rr) parts[rr] = cluster.map_shared_rank(part, rr);
        for (int idx = tid; idx < npairs + 2 * d; idx += 256) {
            float s = parts[0][idx];
#pragma unroll
            for (int rr = 1; rr < S; ++rr) s += parts[rr][idx];
            if (idx < npairs) attn[idx] = s;
            else nrm[idx - npairs] = fmaxf(sqrtf(s), 1e-12f);   // F.normalize: x / max(||x||, 1e-12)
        }
        cluster.sync();   // nobody leaves (or overwrites `part`) while a peer still reads it
    } else {
        __syncthreads();
        for (int idx = tid; idx < npairs + 2 * d; idx += 256) {
            const float s = part[idx];
            if (idx < npairs) attn[idx] = s;
            else nrm[idx - npairs] = fmaxf(sqrtf(s), 1e-12f);
        }
        __syncthreads();
    }
    const float temp = temperature[h];
    for (int idx = tid; idx < npairs; idx += 256) attn[idx] = attn[idx] / (nrm[idx / d] * nrm[d + idx % d]) * temp;
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

    // 4. fold into the projection: rows h*d + j of the per-frame K-major weight, this CTA's columns [o0, o0 + cols)
    float* wo = wt_eff + (long long)b * wt_eff_bs;
    if ((d & 3) == 0 && (d >> 2) * cols >= 256) {   // (small slices keep one output per thread: all 256 threads busy)
        // one output column o (coalesced) and FOUR consecutive j per thread: a projection value is loaded once for four FMAs, the four
        // attention values are one 16-byte broadcast (2 loads per 4 FMAs instead of 2 per FMA; same sum order over i per output)
        const int dq = d >> 2;
        for (int idx = tid; idx < dq * cols; idx += 256) {
            const int jq = idx / cols, o = idx - jq * cols;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            if (o0 + o < C) {
                for (int i = 0; i < d; ++i) {
                    const float wv = pw[i * ow + o];
                    const float4 a = *reinterpret_cast<const float4*>(attn + i * d + 4 * jq);
                    s0 = fmaf(wv, a.x, s0);
                    s1 = fmaf(wv, a.y, s1);
                    s2 = fmaf(wv, a.z, s2);
                    s3 = fmaf(wv, a.w, s3);
                }
            }
            float* wrow = wo + (long long)(h * d + 4 * jq) * ldw + o0 + o;
            wrow[0] = s0, wrow[ldw] = s1, wrow[2 * ldw] = s2, wrow[3 * ldw] = s3;
        }
        return;
    }
    for (int idx = tid; idx < d * cols; idx += 256) {
        const int j = idx / cols, o = idx - j * cols;
        float s = 0.f;
        if (o0 + o < C)
            for (int i = 0; i < d; ++i) s = fmaf(pw[i * ow + o], attn[i * d + j], s);
        wo[(long long)(h * d + j) * ldw + o0 + o] = s;
    }
}

template <int S>
static int launch_xca(const float* qkv, long long qkv_bs, const float* temperature, const float* proj_wt, int ldw, float* wt_eff,
                      long long wt_eff_bs, int B, int C, int heads, int N, cudaStream_t st) {
    const int d = C / heads;
    const int ow = ((ldw / 4 + S - 1) / S) * 4;   // column slice per CTA (multiple of 4)
    const size_t smem = (size_t)(4 * d * XCA_ST + xca_r4(d * d + 2 * d) + xca_r4(d * d) + xca_r4(2 * d) + d * ow) * sizeof(float);
    ACH_REQUIRE(smem <= 160 * 1024, "ach_xca_fold: d=%d, ldw=%d do not fit shared memory", d, ldw);
    static PerDeviceOnce attr_once;
    if (attr_once.first()) cudaFuncSetAttribute(xca_fold_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(heads * S, B);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = S > 1 ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, xca_fold_kernel<S>, qkv, qkv_bs, temperature, proj_wt, ldw, wt_eff, wt_eff_bs, C, heads, N, ow);
    if (e != cudaSuccess) {
        set_error("ach_xca_fold: launch failed: %s", cudaGetErrorString(e));
        return ACH_ERR_CUDA;
    }
    return check_launch("ach_xca_fold");
}

}  // namespace ach

extern "C" int ach_xca_fold(const float* qkv, long long qkv_bs, const float* temperature, const float* proj_wt, int ldw,
                            float* wt_eff, long long wt_eff_bs, int B, int C, int heads, int N, void* stream) {
    using namespace ach;
    ACH_REQUIRE(qkv && temperature && proj_wt && wt_eff, "ach_xca_fold: null arg");
    ACH_REQUIRE(B > 0 && B <= 65535 && heads > 0 && C % heads == 0 && N > 0, "ach_xca_fold: bad dims");
    const int d = C / heads;
    ACH_REQUIRE(d <= XCA_MAX_D, "ach_xca_fold: head dim %d > %d", d, XCA_MAX_D);
    ACH_REQUIRE(ldw >= C && ldw % 4 == 0, "ach_xca_fold: ldw < C or ldw %% 4 != 0");
    ACH_REQUIRE(N % 4 == 0 && aligned16(qkv) && qkv_bs % 4 == 0, "ach_xca_fold: needs N %% 4 == 0 and a 16-byte aligned qkv view");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nchunks = (N + XCA_CHUNK - 1) / XCA_CHUNK;
    // tokens of one (head, frame) over a cluster of S CTAs: 4 x 64 x 4 = 1024 CTAs of 6-7 chunks at 40 x 40, 2 chunks at 20 x 20
    if (nchunks >= 4) return launch_xca<4>(qkv, qkv_bs, temperature, proj_wt, ldw, wt_eff, wt_eff_bs, B, C, heads, N, st);
    if (nchunks >= 2) return launch_xca<2>(qkv, qkv_bs, temperature, proj_wt, ldw, wt_eff, wt_eff_bs, B, C, heads, N, st);
    return launch_xca<1>(qkv, qkv_bs, temperature, proj_wt, ldw, wt_eff, wt_eff_bs, B, C, heads, N, st);
}
