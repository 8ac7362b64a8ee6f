// RCBlock body (RCNet radar encoder): offset + modulator 3x3 convs, modulated deformable 3x3 conv
// (DCNv2, torchvision.ops.deform_conv2d semantics), 1x1 conv + BN + ReLU and the residual add, in ONE
// kernel - the reference runs im2col + addmm per frame plus five more launches (RadarEncoder.py:65-72,
// dcn.py:49-63).  One thread per output pixel; all weights of the block live in shared memory
// (K-major, output-contiguous -> float4 broadcasts); the 27 offset/modulator values, the C deformable
// accumulators and the 1x1 outputs stay in registers.  The pooled input is tiny (<= 36 channels) and is
// gathered through L1/L2 (bilinear taps land anywhere, so no static tile helps).
#include "common.cuh"

namespace ach {

template <int C>
__global__ void __launch_bounds__(128) rc_deform_kernel(const AchRcDeform p) {
    constexpr int CP = (C + 3) & ~3;  // accumulators padded to float4 granularity
    extern __shared__ __align__(16) float smem[];
    float* s_om = smem;                  // [C*9][28]
    float* s_reg = s_om + C * 9 * 28;    // [C*9][CP]
    float* s_w1 = s_reg + C * 9 * CP;    // [C][CP]
    float* s_bom = s_w1 + C * CP;        // [28]
    for (int i = threadIdx.x; i < C * 9 * 28; i += 128) s_om[i] = p.w_om[i];
    for (int i = threadIdx.x; i < C * 9 * CP; i += 128) {
        const int r = i / CP, o = i - r * CP;
        s_reg[i] = (o < C) ? p.w_reg[r * C + o] : 0.f;
    }
    for (int i = threadIdx.x; i < C * CP; i += 128) {
        const int r = i / CP, o = i - r * CP;
        s_w1[i] = (o < C) ? p.w1[r * C + o] : 0.f;
    }
    if (threadIdx.x < 28) s_bom[threadIdx.x] = (threadIdx.x < 27) ? p.b_om[threadIdx.x] : 0.f;
    __syncthreads();

    const int H = p.H, W = p.W;
    const int P = H * W;
    const int pix = blockIdx.x * 128 + threadIdx.x;
    if (pix >= P) return;
    const int b = blockIdx.y;
    const int y = pix / W, x = pix - y * W;
    const float* __restrict__ pooled = p.pooled + (long long)b * p.pooled_bs;

    // ---- offset (18) + modulator (9) 3x3 convolutions over the pooled map (zero padding)
    float om[28];
#pragma unroll
    for (int i = 0; i < 28; ++i) om[i] = s_bom[i];
    for (int c = 0; c < C; ++c) {
        const float* pc = pooled + (long long)c * P;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
            float v = 0.f;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(pc + yy * W + xx);
            const float4* w4 = reinterpret_cast<const float4*>(s_om + (c * 9 + t) * 28);
#pragma unroll
            for (int i = 0; i < 7; ++i) fma4_bcast(om + 4 * i, v, w4[i]);
        }
    }

    // ---- modulated deformable conv: tap k = i*3 + j, offsets (dy, dx) = om[2k], om[2k+1], mask = 2*sigmoid(om[18+k])
    float acc[CP];
#pragma unroll
    for (int i = 0; i < CP; ++i) acc[i] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float py = (float)(y - 1 + t / 3) + om[2 * t];
        const float px = (float)(x - 1 + t % 3) + om[2 * t + 1];
        const float m = 2.0f * sigmoidf_(om[18 + t]);
        const float fy = floorf(py), fx = floorf(px);
        const int y0 = (int)fy, x0 = (int)fx;
        const float ly = py - fy, lx = px - fx;
        const float hy = 1.f - ly, hx = 1.f - lx;
        const bool in = (py > -1.f) && (py < (float)H) && (px > -1.f) && (px < (float)W);
        const bool y0ok = in && y0 >= 0, y1ok = in && (y0 + 1 <= H - 1);
        const bool x0ok = x0 >= 0, x1ok = (x0 + 1 <= W - 1);
        const float w00 = (y0ok && x0ok) ? hy * hx : 0.f;
        const float w01 = (y0ok && x1ok) ? hy * lx : 0.f;
        const float w10 = (y1ok && x0ok) ? ly * hx : 0.f;
        const float w11 = (y1ok && x1ok) ? ly * lx : 0.f;
        const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
        const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
        const int i00 = yc0 * W + xc0, i01 = yc0 * W + xc1, i10 = yc1 * W + xc0, i11 = yc1 * W + xc1;
        for (int c = 0; c < C; ++c) {
            const float* pc = pooled + (long long)c * P;
            const float v = m * (w00 * __ldg(pc + i00) + w01 * __ldg(pc + i01) + w10 * __ldg(pc + i10) + w11 * __ldg(pc + i11));
            const float4* w4 = reinterpret_cast<const float4*>(s_reg + (c * 9 + t) * CP);
#pragma unroll
            for (int i = 0; i < CP / 4; ++i) fma4_bcast(acc + 4 * i, v, w4[i]);
        }
    }

    // ---- 1x1 conv + folded BN + ReLU + residual
    float z[CP];
#pragma unroll
    for (int i = 0; i < CP; ++i) z[i] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float4* w4 = reinterpret_cast<const float4*>(s_w1 + c * CP);
#pragma unroll
        for (int i = 0; i < CP / 4; ++i) fma4_bcast(z + 4 * i, acc[c], w4[i]);
    }
    const float* __restrict__ xr = p.x + (long long)b * p.x_bs + pix;
    float* __restrict__ orow = p.out + (long long)b * p.out_bs + pix;
#pragma unroll
    for (int o = 0; o < C; ++o)
        orow[(long long)o * P] = xr[(long long)o * P] + fmaxf(fmaf(p.scale[o], z[o], p.bias[o]), 0.f);
}

// ---- channel-last pooled map (fast path): pooled[b][pixel][CP].  The 3x3 window of the offset/modulator convs is 9
// coalesced 16-byte loads per 4 channels (planar: 36 scalar loads), and each bilinear corner of the deformable conv is one
// 16-byte gather per 4 channels instead of 4 scattered 4-byte gathers - the planar kernel spent ~45 % of its issue
// slots on gather loads and their 64-bit address arithmetic (ncu: lsu pipe 42-47 %).  Arithmetic per element is
// unchanged (same bilinear expression, accumulation order tap-major then channel).
template <int C>
__global__ void __launch_bounds__(128) rc_deform_cl_kernel(const AchRcDeform p) {
    constexpr int CP = (C + 3) & ~3;
    constexpr int Q = CP / 4;
    constexpr int QU = Q <= 2 ? Q : 1;   // wider blocks keep the channel-group loop rolled (registers: 176 -> ~100 at C = 12)
    extern __shared__ __align__(16) float smem[];
    float* s_om = smem;                  // [C*9][28]
    float* s_reg = s_om + C * 9 * 28;    // [C*9][CP]
    float* s_w1 = s_reg + C * 9 * CP;    // [C][CP]
    float* s_bom = s_w1 + C * CP;        // [28]
    for (int i = threadIdx.x; i < C * 9 * 28; i += 128) s_om[i] = p.w_om[i];
    for (int i = threadIdx.x; i < C * 9 * CP; i += 128) {
        const int r = i / CP, o = i - r * CP;
        s_reg[i] = (o < C) ? p.w_reg[r * C + o] : 0.f;
    }
    for (int i = threadIdx.x; i < C * CP; i += 128) {
        const int r = i / CP, o = i - r * CP;
        s_w1[i] = (o < C) ? p.w1[r * C + o] : 0.f;
    }
    if (threadIdx.x < 28) s_bom[threadIdx.x] = (threadIdx.x < 27) ? p.b_om[threadIdx.x] : 0.f;
    __syncthreads();

    const int H = p.H, W = p.W;
    const int P = H * W;
    const int pix = blockIdx.x * 128 + threadIdx.x;
    if (pix >= P) return;
    const int b = blockIdx.y;
    const int y = pix / W, x = pix - y * W;
    const float4* __restrict__ pooled = reinterpret_cast<const float4*>(p.pooled + (long long)b * p.pooled_bs);

    // ---- offset (18) + modulator (9) 3x3 convolutions over the pooled map (zero padding)
    float om[28];
#pragma unroll
    for (int i = 0; i < 28; ++i) om[i] = s_bom[i];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
        const float4* src = pooled + (ok ? (yy * W + xx) * Q : 0);
#pragma unroll(QU)
        for (int q = 0; q < Q; ++q) {
            const float4 v4 = ok ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (q * 4 + e < C) {
                    const float4* w4 = reinterpret_cast<const float4*>(s_om + ((q * 4 + e) * 9 + t) * 28);
#pragma unroll
                    for (int i = 0; i < 7; ++i) fma4_bcast(om + 4 * i, v[e], w4[i]);
                }
            }
        }
    }

    // ---- modulated deformable conv: tap k = i*3 + j, offsets (dy, dx) = om[2k], om[2k+1], mask = 2*sigmoid(om[18+k])
    float acc[CP];
#pragma unroll
    for (int i = 0; i < CP; ++i) acc[i] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float py = (float)(y - 1 + t / 3) + om[2 * t];
        const float px = (float)(x - 1 + t % 3) + om[2 * t + 1];
        const float m = 2.0f * sigmoidf_(om[18 + t]);
        const float fy = floorf(py), fx = floorf(px);
        const int y0 = (int)fy, x0 = (int)fx;
        const float ly = py - fy, lx = px - fx;
        const float hy = 1.f - ly, hx = 1.f - lx;
        const bool in = (py > -1.f) && (py < (float)H) && (px > -1.f) && (px < (float)W);
        const bool y0ok = in && y0 >= 0, y1ok = in && (y0 + 1 <= H - 1);
        const bool x0ok = x0 >= 0, x1ok = (x0 + 1 <= W - 1);
        const float w00 = (y0ok && x0ok) ? hy * hx : 0.f;
        const float w01 = (y0ok && x1ok) ? hy * lx : 0.f;
        const float w10 = (y1ok && x0ok) ? ly * hx : 0.f;
        const float w11 = (y1ok && x1ok) ? ly * lx : 0.f;
        const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
        const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
        const float4* g00 = pooled + (yc0 * W + xc0) * Q;
        const float4* g01 = pooled + (yc0 * W + xc1) * Q;
        const float4* g10 = pooled + (yc1 * W + xc0) * Q;
        const float4* g11 = pooled + (yc1 * W + xc1) * Q;
#pragma unroll(QU)
        for (int q = 0; q < Q; ++q) {
            const float4 a4 = __ldg(g00 + q), b4 = __ldg(g01 + q), c4 = __ldg(g10 + q), d4 = __ldg(g11 + q);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
            const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (q * 4 + e < C) {
                    const float v = m * (w00 * a[e] + w01 * bb[e] + w10 * cc[e] + w11 * dd[e]);
                    const float4* w4 = reinterpret_cast<const float4*>(s_reg + ((q * 4 + e) * 9 + t) * CP);
#pragma unroll
                    for (int i = 0; i < CP / 4; ++i) fma4_bcast(acc + 4 * i, v, w4[i]);
                }
            }
        }
    }

    // ---- 1x1 conv + folded BN + ReLU + residual
    float z[CP];
#pragma unroll
    for (int i = 0; i < CP; ++i) z[i] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float4* w4 = reinterpret_cast<const float4*>(s_w1 + c * CP);
#pragma unroll
        for (int i = 0; i < CP / 4; ++i) fma4_bcast(z + 4 * i, acc[c], w4[i]);
    }
    const float* __restrict__ xr = p.x + (long long)b * p.x_bs + pix;
    float* __restrict__ orow = p.out + (long long)b * p.out_bs + pix;
#pragma unroll
    for (int o = 0; o < C; ++o)
        orow[(long long)o * P] = xr[(long long)o * P] + fmaxf(fmaf(p.scale[o], z[o], p.bias[o]), 0.f);
}

template <int C>
static int launch_rc(const AchRcDeform& p, cudaStream_t st) {
    constexpr int CP = (C + 3) & ~3;
    const size_t smem = (size_t)(C * 9 * 28 + C * 9 * CP + C * CP + 28) * sizeof(float);
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(rc_deform_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(rc_deform_cl_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    const dim3 grid(cdiv((long long)p.H * p.W, 128), p.B);
    if (p.pooled_cl) rc_deform_cl_kernel<C><<<grid, 128, smem, st>>>(p);
    else rc_deform_kernel<C><<<grid, 128, smem, st>>>(p);
    return check_launch("ach_rc_deform");
}

}  // namespace ach

extern "C" int ach_rc_deform(const AchRcDeform* pp, void* stream) {
    using namespace ach;
    const AchRcDeform& p = *pp;
    ACH_REQUIRE(p.x && p.pooled && p.w_om && p.b_om && p.w_reg && p.w1 && p.scale && p.bias && p.out, "ach_rc_deform: null arg");
    ACH_REQUIRE(p.B > 0 && p.B <= 65535 && p.H > 0 && p.W > 0, "ach_rc_deform: bad dims");
    ACH_REQUIRE(!p.pooled_cl || (aligned16(p.pooled) && p.pooled_bs % 4 == 0), "ach_rc_deform: channel-last pooled map must be 16-byte aligned");
    ACH_REQUIRE((long long)p.H * p.W * ((p.C + 3) / 4) < (1LL << 31), "ach_rc_deform: plane too large for 32-bit indexing");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (p.C) {
        case 3: return launch_rc<3>(p, st);
        case 8: return launch_rc<8>(p, st);
        case 12: return launch_rc<12>(p, st);
        case 16: return launch_rc<16>(p, st);
        case 24: return launch_rc<24>(p, st);
        case 30: return launch_rc<30>(p, st);
        case 36: return launch_rc<36>(p, st);
        default: break;
    }
    set_error("ach_rc_deform: C=%d not instantiated (supported: 3, 8, 12, 16, 24, 30, 36)", p.C);
    return ACH_ERR_INVALID;
}
