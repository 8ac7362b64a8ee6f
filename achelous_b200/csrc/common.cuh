// Shared device/host helpers for the achelous_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>

#include "../../include/achelous_b200.h"

namespace ach {

// ---- error plumbing: the C ABI never throws; it returns a status and keeps a thread-local message
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define ACH_REQUIRE(cond, ...)                  \
    do {                                        \
        if (!(cond)) {                          \
            ::ach::set_error(__VA_ARGS__);      \
            return ACH_ERR_INVALID;             \
        }                                       \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- activations (exact forms used by the reference; no fast-math)
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2, ACT_GELU = 3, ACT_SIGMOID = 4 };

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// MUFU.RCP / MUFU.EX2 without the range-fixup code the C library wraps around them (9 and 6 instructions); callers
// guarantee a normal-range argument (rcp) or accept flush-to-zero on underflow (ex2).  Relative error <= 2^-22.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// erf with |error| <= 1.5e-7 (Abramowitz & Stegun 7.1.26): 1 rcp + 1 ex2 + 7 fma instead of the ~40-instruction
// branchy libdevice erff; GELU inherits an absolute error <= 0.75e-7 * |x| - below fp32 rounding of the GEMM feeding it.
__device__ __forceinline__ float erf_as(float x) {
    const float ax = fabsf(x);
    const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));   // argument >= 1
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    poly *= t;
    const float e = ex2_approx(-ax * ax * 1.4426950408889634f);   // underflow -> 0 -> erf = +-1
    return copysignf(fmaf(-poly, e, 1.0f), x);
}

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case ACT_RELU: return fmaxf(v, 0.0f);
        // x * sigmoid(x) on MUFU.EX2 + MUFU.RCP (relative error <= 2^-21; the denominator is >= 1, x -> -inf gives x / inf -> -0 like the
        // exact form): the libm expf + IEEE division of sigmoidf_ cost ~20 instructions per value and ncu showed the MobileViT 1x1
        // layers (K = 16 ... 64, SiLU) bound by exactly this epilogue arithmetic
        case ACT_SILU: return v * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * v));
        case ACT_GELU: {
            const float h = 0.5f * v;
            return fmaf(h, erf_as(v * 0.70710678118654752440f), h);
        }
        case ACT_SIGMOID: return sigmoidf_(v);
        default: return v;
    }
}

// Blackwell packed fp32 FMA (FFMA2): two independent fp32 FMAs per issued instruction; bit-identical to two fmaf().
// acc[0..3] += x * w[0..3] costs 2 issue slots instead of 4 in the issue-bound direct-convolution inner loops.
__device__ __forceinline__ void fma4_bcast(float* acc, float x, const float4& w) {
    const float2 xx = make_float2(x, x);
    float2 a01 = make_float2(acc[0], acc[1]), a23 = make_float2(acc[2], acc[3]);
    a01 = __ffma2_rn(xx, make_float2(w.x, w.y), a01);
    a23 = __ffma2_rn(xx, make_float2(w.z, w.w), a23);
    acc[0] = a01.x; acc[1] = a01.y; acc[2] = a23.x; acc[3] = a23.y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum for blockDim.x == 256 (8 warps); `red` is >= 8 floats of shared memory
__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    return t;
}

// float atomic max that is correct for any sign (memory must be initialised, e.g. to -inf)
__device__ __forceinline__ void atomic_max_float(float* addr, float val) {
    if (val >= 0.f)
        atomicMax(reinterpret_cast<int*>(addr), __float_as_int(val));
    else
        atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(val));
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// cudaFuncSetAttribute (opt-in dynamic shared memory) and occupancy figures are PER DEVICE; nn.DataParallel drives several GPUs
// from one process (one host thread per GPU), so "done once" flags are kept per device ordinal.
constexpr int ACH_MAX_DEVICES = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev & (ACH_MAX_DEVICES - 1);
}
struct PerDeviceOnce {
    std::atomic<unsigned long long> done{0};
    bool first() {
        const unsigned long long bit = 1ull << current_device();
        return !(done.fetch_or(bit) & bit);
    }
};

}  // namespace ach
