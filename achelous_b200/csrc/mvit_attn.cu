// MobileViT token self-attention (mobilevit.py:48-73 inside MobileViTBlock :134-165).
//
// The reference unfolds the (B, D, H, W) map into 2x2 patch-position groups
//   'b d (h ph) (w pw) -> b (ph pw) (h w) d'          (mobilevit.py:156)
// runs softmax(Q K^T / sqrt(d)) V per (group, head) over the N = H*W/4 tokens of a group, and folds back.
// Here nothing is unfolded: qkv stays channel-major (B, 3*heads*d, H*W) as produced by the fused LN+qkv GEMM,
// a token is a pixel, and its group is (y & 1, x & 1).  One CTA per (frame, group, head) stages the group's
// K and V (N x d each) in shared memory; each thread owns one query and streams over the keys with an online
// softmax (running max / sum), so the N x N score tensor (10 MB/frame/layer in the reference) never exists.
// The result is written back channel-major at the query's pixel, ready for the out-projection GEMM.
#include <cstdlib>

#include "common.cuh"

namespace ach {

template <int D>
__global__ void __launch_bounds__(256) mvit_attn_kernel(const float* __restrict__ qkv, long long qkv_bs, float* __restrict__ out,
                                                        long long out_bs, int heads, int H, int W, float scale) {
    extern __shared__ float smem[];
    const int hw2 = W / 2;
    const int N = (H / 2) * hw2;
    float* ks = smem;            // [N][D]
    float* vs = smem + N * D;    // [N][D]
    const int head = blockIdx.x % heads;
    const int g = blockIdx.x / heads;        // patch position: ph = g >> 1, pw = g & 1
    const int b = blockIdx.y;
    const int ph = g >> 1, pw = g & 1;
    const int P = H * W;
    const int inner = heads * D;
    const float* qb = qkv + (long long)b * qkv_bs + (long long)(head * D) * P;
    const float* kb = qb + (long long)inner * P;
    const float* vb = kb + (long long)inner * P;

    for (int i = threadIdx.x; i < N * D; i += 256) {
        const int dd = i / N, n = i - dd * N;           // consecutive threads -> consecutive tokens of one channel
        const int pix = (2 * (n / hw2) + ph) * W + 2 * (n % hw2) + pw;
        ks[n * D + dd] = kb[(long long)dd * P + pix];
        vs[n * D + dd] = vb[(long long)dd * P + pix];
    }
    __syncthreads();

    float* ob = out + (long long)b * out_bs + (long long)(head * D) * P;
    const float scale2 = scale * 1.4426950408889634f;   // one MUFU.EX2 per key instead of expf (the exponentials were half the work at d = 8)
    for (int n = threadIdx.x; n < N; n += 256) {
        const int pix = (2 * (n / hw2) + ph) * W + 2 * (n % hw2) + pw;
        float q[D];
#pragma unroll
        for (int dd = 0; dd < D; ++dd) q[dd] = qb[(long long)dd * P + pix] * scale2;   // log2 units: exp(x) = 2^(x * log2 e)
        float m = -INFINITY, l = 0.f;
        float acc[D];
#pragma unroll
        for (int dd = 0; dd < D; ++dd) acc[dd] = 0.f;
        for (int j = 0; j < N; ++j) {
            const float* kj = ks + j * D;
            float s = 0.f;
#pragma unroll
            for (int dd = 0; dd < D; ++dd) s = fmaf(q[dd], kj[dd], s);
            if (s > m) {                      // rescale only when the running max moves
                const float c = ex2_approx(m - s);
                l *= c;
#pragma unroll
                for (int dd = 0; dd < D; ++dd) acc[dd] *= c;
                m = s;
            }
            const float e = ex2_approx(s - m);
            l += e;
            const float* vj = vs + j * D;
#pragma unroll
            for (int dd = 0; dd < D; ++dd) acc[dd] = fmaf(e, vj[dd], acc[dd]);
        }
        const float inv = 1.0f / l;
#pragma unroll
        for (int dd = 0; dd < D; ++dd) ob[(long long)dd * P + pix] = acc[dd] * inv;
    }
}

}  // namespace ach

int mvit_attention_tc_launch(const float* qkv, long long qkv_bs, float* out, long long out_bs, int B, int heads, int H, int W, float scale,
                             cudaStream_t st);   // mvit_attn_tc.cu

extern "C" int ach_mvit_attention(const float* qkv, long long qkv_bs, float* out, long long out_bs, int B, int heads, int dim_head,
                                  int H, int W, void* stream) {
    using namespace ach;
    ACH_REQUIRE(qkv && out && B > 0 && B <= 65535 && heads > 0, "ach_mvit_attention: bad args");
    ACH_REQUIRE(dim_head == 8, "ach_mvit_attention: dim_head=%d not instantiated (MobileViT uses 8)", dim_head);
    ACH_REQUIRE(H % 2 == 0 && W % 2 == 0, "ach_mvit_attention: H, W must be even (2x2 patches)");
    // The tensor-core variant is a separate entry point (ach_mvit_attention_tc below).  Measured on B200 at 40x40 (400 tokens per
    // group, B = 64): tcgen05 0.375 ms vs 0.221 ms for this file's CUDA-core kernel - with d = 8 each 16-key step of P.V is a full
    // TMEM round trip (ld -> 2^x -> split -> st -> barrier -> MMA) for 256 MACs per query, and TMEM (256 columns per CTA)
    // caps the SM at two such chains; the plan therefore calls this CUDA-core kernel.
    const int N = (H / 2) * (W / 2);
    const size_t smem = (size_t)2 * N * dim_head * sizeof(float);
    ACH_REQUIRE(smem <= 200 * 1024, "ach_mvit_attention: %d tokens per group do not fit shared memory", N);
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(mvit_attn_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    const float scale = 1.0f / sqrtf((float)dim_head);
    mvit_attn_kernel<8><<<dim3(4 * heads, B), 256, smem, (cudaStream_t)stream>>>(qkv, qkv_bs, out, out_bs, heads, H, W, scale);
    return check_launch("ach_mvit_attention");
}

// tcgen05 variant of the same contract (mvit_attn_tc.cu): parity-tested, not on the default plan (see above).  Shapes whose
// token count does not fit its shared-memory layout are rejected with ACH_ERR_INVALID - there is no silent fallback.
extern "C" int ach_mvit_attention_tc(const float* qkv, long long qkv_bs, float* out, long long out_bs, int B, int heads, int dim_head,
                                     int H, int W, void* stream) {
    using namespace ach;
    ACH_REQUIRE(qkv && out && B > 0 && B <= 65535 && heads > 0, "ach_mvit_attention_tc: bad args");
    ACH_REQUIRE(dim_head == 8, "ach_mvit_attention_tc: dim_head=%d not instantiated (MobileViT uses 8)", dim_head);
    ACH_REQUIRE(H % 2 == 0 && W % 2 == 0, "ach_mvit_attention_tc: H, W must be even (2x2 patches)");
    const int rc = mvit_attention_tc_launch(qkv, qkv_bs, out, out_bs, B, heads, H, W, 1.0f / sqrtf((float)dim_head), (cudaStream_t)stream);
    ACH_REQUIRE(rc >= 0, "ach_mvit_attention_tc: %d x %d tokens do not fit the tensor-core kernel's shared-memory layout", H, W);
    return rc;
}
