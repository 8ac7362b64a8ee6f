// On-device input pre-processing (reference: achelous.py:200-246, utils/utils.py:20-54).
//
//   ach_pre_resize_h       Pillow's 8-bit horizontal resampling pass (Resample.c ImagingResampleHorizontal_8bpc): 22-bit
//                          fixed-point coefficients, int32 accumulator started at 1 << 21, >> 22 with saturation.
//   ach_pre_resize_v_norm  the vertical pass fused with the letterbox paste (grey 128 border), the HWC -> CHW transpose and
//                          preprocess_input (/255 in fp32, then -mean and /std each computed in fp64 and rounded to fp32,
//                          which is what numpy's in-place ops with float64 operands do) - the uint8 letterboxed image never
//                          exists in memory.
//   ach_pre_radar          per-sample min-max normalisation + 1e-13 (preprocess_input_radar), fp32 or fp64 input.
//   ach_pre_points         row gather by index, column L2 norms over the sampled rows in fp64 (sklearn normalize(axis=0)),
//                          division, fp32 output transposed to (B, C, N).
// The coefficient tables are tiny (out_size x ksize) and are computed on the host in float64 exactly as Pillow does
// (achelous_b200/utils/preprocess.py); all pixel arithmetic is integer and therefore bit-exact.
#include "common.cuh"

namespace ach {

constexpr int PRE_BITS = 32 - 8 - 2;   // Pillow PRECISION_BITS

__device__ __forceinline__ int clip8(int acc) { return min(max(acc >> PRE_BITS, 0), 255); }

// one thread per (row, output column); src row r0 + row of image b; 3 interleaved channels
__global__ void __launch_bounds__(256) pre_resize_h_kernel(const uint8_t* __restrict__ src, long long src_bs, int iw, int r0, int rows, int nw,
                                                           const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                           uint8_t* __restrict__ tmp, long long tmp_bs) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y, b = blockIdx.z;
    if (xx >= nw) return;
    const int xmin = bounds[2 * xx], cnt = bounds[2 * xx + 1];
    const int* __restrict__ k = kk + (long long)xx * ksize;
    const uint8_t* __restrict__ s = src + (long long)b * src_bs + ((long long)(r0 + row) * iw + xmin) * 3;
    int a0 = 1 << (PRE_BITS - 1), a1 = a0, a2 = a0;
    for (int x = 0; x < cnt; ++x) {
        const int w = k[x];
        a0 += (int)s[3 * x + 0] * w;
        a1 += (int)s[3 * x + 1] * w;
        a2 += (int)s[3 * x + 2] * w;
    }
    uint8_t* __restrict__ o = tmp + (long long)b * tmp_bs + ((long long)row * nw + xx) * 3;
    o[0] = (uint8_t)clip8(a0);
    o[1] = (uint8_t)clip8(a1);
    o[2] = (uint8_t)clip8(a2);
}

__device__ __forceinline__ float norm_px(int v, double mean, double std_) {
    const float a = __fdiv_rn((float)v, 255.0f);              // image /= 255.0 (float32, IEEE division)
    const float b = (float)__dsub_rn((double)a, mean);       // image -= float64 array: computed in fp64, stored fp32
    return (float)__ddiv_rn((double)b, std_);                // image /= float64 array
}

// one thread per network-input pixel (Y, X); tmp holds rows [t0, t0 + trows) of the horizontally resampled image
__global__ void __launch_bounds__(256) pre_resize_v_norm_kernel(const uint8_t* __restrict__ tmp, long long tmp_bs, int nw, int nh,
                                                                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                                int identity, float* __restrict__ out, long long out_bs, int H, int W,
                                                                int x_off, int y_off) {
    const int X = blockIdx.x * blockDim.x + threadIdx.x;
    const int Y = blockIdx.y, b = blockIdx.z;
    if (X >= W) return;
    const int xx = X - x_off, yy = Y - y_off;
    int v0 = 128, v1 = 128, v2 = 128;   // Image.new('RGB', size, (128, 128, 128))
    if (xx >= 0 && xx < nw && yy >= 0 && yy < nh) {
        const uint8_t* __restrict__ t = tmp + (long long)b * tmp_bs;
        if (identity) {
            const uint8_t* s = t + ((long long)yy * nw + xx) * 3;
            v0 = s[0], v1 = s[1], v2 = s[2];
        } else {
            const int ymin = bounds[2 * yy], cnt = bounds[2 * yy + 1];
            const int* __restrict__ k = kk + (long long)yy * ksize;
            int a0 = 1 << (PRE_BITS - 1), a1 = a0, a2 = a0;
            for (int j = 0; j < cnt; ++j) {
                const uint8_t* s = t + ((long long)(ymin + j) * nw + xx) * 3;
                const int w = k[j];
                a0 += (int)s[0] * w;
                a1 += (int)s[1] * w;
                a2 += (int)s[2] * w;
            }
            v0 = clip8(a0), v1 = clip8(a1), v2 = clip8(a2);
        }
    }
    float* __restrict__ o = out + (long long)b * out_bs + (long long)Y * W + X;
    const long long plane = (long long)H * W;
    o[0] = norm_px(v0, 0.485, 0.229);
    o[plane] = norm_px(v1, 0.456, 0.224);
    o[2 * plane] = norm_px(v2, 0.406, 0.225);
}

// ---- radar map: one CTA of 1024 threads per sample (two sweeps over n elements)
template <typename T>
__global__ void __launch_bounds__(1024) pre_radar_kernel(const T* __restrict__ src, long long src_bs, long long n, float* __restrict__ out,
                                                         long long out_bs) {
    __shared__ T s_min[32], s_max[32];
    const T* __restrict__ x = src + (long long)blockIdx.x * src_bs;
    T lo = x[0], hi = x[0];
    for (long long i = threadIdx.x; i < n; i += 1024) {
        const T v = x[i];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = lo, s_max[threadIdx.x >> 5] = hi;
    __syncthreads();
    lo = s_min[0], hi = s_max[0];
#pragma unroll
    for (int i = 1; i < 32; ++i) {
        lo = s_min[i] < lo ? s_min[i] : lo;
        hi = s_max[i] > hi ? s_max[i] : hi;
    }
    float* __restrict__ o = out + (long long)blockIdx.x * out_bs;
    if constexpr (sizeof(T) == 8) {
        const double range = __dsub_rn(hi, lo);
        for (long long i = threadIdx.x; i < n; i += 1024)
            o[i] = (float)__dadd_rn(__ddiv_rn(__dsub_rn(x[i], lo), range), 0.0000000000001);
    } else {
        const float range = __fsub_rn(hi, lo);
        for (long long i = threadIdx.x; i < n; i += 1024)
            o[i] = __fadd_rn(__fdiv_rn(__fsub_rn(x[i], lo), range), (float)0.0000000000001);   // weak Python scalar -> float32
    }
}

// ---- points: one CTA of 256 threads per (column c, sample b)
__global__ void __launch_bounds__(256) pre_points_kernel(const double* __restrict__ feat, int n_rows, int C, const int* __restrict__ idx,
                                                         int N, float* __restrict__ out) {
    __shared__ double red[256];
    const int c = blockIdx.x, b = blockIdx.y;
    const int* __restrict__ id = idx + (long long)b * N;
    double ss = 0.0;
    for (int i = threadIdx.x; i < N; i += 256) {
        const double v = feat[(long long)id[i] * C + c];
        ss = __dadd_rn(ss, __dmul_rn(v, v));
    }
    red[threadIdx.x] = ss;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = __dadd_rn(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    double norm = sqrt(red[0]);
    if (norm == 0.0) norm = 1.0;
    float* __restrict__ o = out + ((long long)b * C + c) * N;
    for (int i = threadIdx.x; i < N; i += 256) o[i] = (float)__ddiv_rn(feat[(long long)id[i] * C + c], norm);
    (void)n_rows;
}

}  // namespace ach

extern "C" int ach_pre_resize_h(const unsigned char* src, long long src_bs, int B, int iw, int r0, int rows, int nw, const int* bounds,
                                const int* kk, int ksize, unsigned char* tmp, long long tmp_bs, void* stream) {
    using namespace ach;
    ACH_REQUIRE(src && bounds && kk && tmp, "ach_pre_resize_h: null arg");
    ACH_REQUIRE(B > 0 && B <= 65535 && iw > 0 && r0 >= 0 && rows > 0 && rows <= 65535 && nw > 0 && ksize > 0, "ach_pre_resize_h: bad dims");
    dim3 grid(cdiv(nw, 256), rows, B);
    pre_resize_h_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_bs, iw, r0, rows, nw, bounds, kk, ksize, tmp, tmp_bs);
    return check_launch("ach_pre_resize_h");
}

extern "C" int ach_pre_resize_v_norm(const unsigned char* tmp, long long tmp_bs, int B, int nw, int nh, const int* bounds, const int* kk,
                                     int ksize, int identity, float* out, long long out_bs, int H, int W, int x_off, int y_off,
                                     void* stream) {
    using namespace ach;
    ACH_REQUIRE(tmp && out && (identity || (bounds && kk && ksize > 0)), "ach_pre_resize_v_norm: null arg");
    ACH_REQUIRE(B > 0 && B <= 65535 && nw > 0 && nh > 0 && H > 0 && H <= 65535 && W > 0, "ach_pre_resize_v_norm: bad dims");
    ACH_REQUIRE(x_off >= 0 && y_off >= 0 && x_off + nw <= W && y_off + nh <= H, "ach_pre_resize_v_norm: the resized image does not fit the input");
    dim3 grid(cdiv(W, 256), H, B);
    pre_resize_v_norm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tmp, tmp_bs, nw, nh, bounds, kk, ksize, identity, out, out_bs, H, W,
                                                                     x_off, y_off);
    return check_launch("ach_pre_resize_v_norm");
}

extern "C" int ach_pre_radar(const void* src, long long src_bs, int is_f64, int B, long long n, float* out, long long out_bs, void* stream) {
    using namespace ach;
    ACH_REQUIRE(src && out && B > 0 && n > 0, "ach_pre_radar: bad args");
    if (is_f64)
        pre_radar_kernel<double><<<B, 1024, 0, (cudaStream_t)stream>>>(static_cast<const double*>(src), src_bs, n, out, out_bs);
    else
        pre_radar_kernel<float><<<B, 1024, 0, (cudaStream_t)stream>>>(static_cast<const float*>(src), src_bs, n, out, out_bs);
    return check_launch("ach_pre_radar");
}

extern "C" int ach_pre_points(const double* feat, int n_rows, int C, const int* idx, int B, int N, float* out, void* stream) {
    using namespace ach;
    ACH_REQUIRE(feat && idx && out && n_rows > 0 && C > 0 && C <= 65535 && B > 0 && B <= 65535 && N > 0, "ach_pre_points: bad args");
    pre_points_kernel<<<dim3(C, B), 256, 0, (cudaStream_t)stream>>>(feat, n_rows, C, idx, N, out);
    return check_launch("ach_pre_points");
}
