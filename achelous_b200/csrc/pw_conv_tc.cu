// Pointwise convolution / Linear on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate.
//
//   OUT[O][P] = W[O][K] * X[K][P]  per frame, same prologue/epilogue contract as ach_pw_conv (AchPwConv).
//
// The SIMT kernel is instruction-issue bound (ncu: fma pipe ~35 %, issue slots ~70 %): one FFMA does 32 MACs
// per issue slot, one tcgen05.mma (M=128, N=128, K=8, kind::tf32) does 131072.  To keep the reference's fp32
// accuracy (north-star tolerance 1e-3, observed ~1e-6) every operand is split into two TF32 terms,
// x = x_hi + x_lo, w = w_hi + w_lo (x_hi = round-to-tf32(x), x_lo = x - x_hi), and three MMAs accumulate
// x_hi w_hi + x_lo w_hi + x_hi w_lo in the fp32 TMEM accumulator ("3xTF32"; the dropped x_lo w_lo term is
// ~2^-22 relative).
//
// Mapping: UMMA M = 128 pixels (TMEM lanes), N = NT <= 128 outputs (TMEM columns), K = 8 per instruction.
// Both operands are K-major, no swizzle ("interleaved" 8 x 16 B core matrices):
//   A = X^T tile [128 px][32 k]: smem [k-core (4 k)][m-core (8 px)][8 px rows][4 k].  Thread t owns pixel t:
//       it gathers 4 consecutive k of its pixel with 4 warp-coalesced 128-byte loads and writes them as ONE
//       16-byte shared store at float offset kcore*512 + t*4 - consecutive threads hit consecutive 16-byte
//       slots, so the transposition costs no bank conflicts (descriptor: LBO = 2048 B between k cores,
//       SBO = 128 B between pixel cores).  An MN-major A descriptor would avoid the transposition but was
//       measured to yield zeros for kind::tf32 on this part (tools/probe/tc_probe.cu, modes 1-2).
//   B = W tile [NT outputs][32 k], pre-packed on the device by ach_pack_pw_tc into exactly the shared-memory
//       image [k-core][n-core][8 rows][4 k] so the kernel copies it linearly
//       (descriptor: LBO = NT/8 * 128 B between k cores, SBO = 128 B between n cores).
// This file holds the C-ABI entry (argument validation) and the device-side weight packing; the kernel itself - warp-specialised,
// asynchronous, A operand in tensor memory - is pw_conv_tc_ws.cu.  The LayerNorm prologue rides along as
//   LN(x) . w = rstd * (x . w - mean * sum_k w)      (wsum = row sums of the folded weights, from the host)
// so the activations are read exactly once.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace ach {

// ---- device-side packing of K-major [K][ldw] weights into hi/lo UMMA tiles
__global__ void __launch_bounds__(256) pack_pw_tc_kernel(const float* __restrict__ wt, int K, int O, int ldw, int NT, int n_kchunks,
                                                         float* __restrict__ hi, float* __restrict__ lo, long long total) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int blk_elems = NT * TC_KC;
    const long long blk = i / blk_elems;
    int r = (int)(i - blk * blk_elems);
    const int e = r & 3;  r >>= 2;          // k within core
    const int row = r & 7; r >>= 3;         // n within core
    const int ncore = r % (NT / 8);
    const int kcore = r / (NT / 8);
    const int o_tile = (int)(blk / n_kchunks), c = (int)(blk % n_kchunks);
    const int o = o_tile * NT + ncore * 8 + row;
    const int k = c * TC_KC + kcore * 4 + e;
    float w = 0.f;
    if (o < O && k < K) w = wt[(long long)k * ldw + o];
    const float h = to_tf32(w);
    hi[i] = h;
    lo[i] = w - h;
}

static int tc_tile_n(int O) { return O <= 32 ? 32 : (O <= 64 ? 64 : 128); }

int pw_conv_tc_ws_launch(const AchPwConv& p, const float* w_hi, const float* w_lo, const float* wsum, cudaStream_t st);   // pw_conv_tc_ws.cu

}  // namespace ach

extern "C" long long ach_pack_pw_tc_elems(int K, int O) {
    using namespace ach;
    const int NT = tc_tile_n(O);
    return (long long)cdiv(O, NT) * cdiv(K, TC_KC) * NT * TC_KC;
}

extern "C" int ach_pack_pw_tc(const float* wt, int K, int O, int ldw, float* w_hi, float* w_lo, void* stream) {
    using namespace ach;
    ACH_REQUIRE(wt && w_hi && w_lo && K > 0 && O > 0 && ldw >= O, "ach_pack_pw_tc: bad args");
    const int NT = tc_tile_n(O);
    const long long total = ach_pack_pw_tc_elems(K, O);
    pack_pw_tc_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(wt, K, O, ldw, NT, cdiv(K, TC_KC), w_hi, w_lo, total);
    return check_launch("ach_pack_pw_tc");
}

/* explicit tile width (the fused MLP kernel wants W1 in 32-column tiles and W2 in one C-column tile) */
extern "C" long long ach_pack_pw_tc_nt_elems(int K, int O, int NT) {
    using namespace ach;
    if (NT <= 0 || NT % 8 != 0) return -1;
    return (long long)cdiv(O, NT) * cdiv(K, TC_KC) * NT * TC_KC;
}

extern "C" int ach_pack_pw_tc_nt(const float* wt, int K, int O, int ldw, int NT, float* w_hi, float* w_lo, void* stream) {
    using namespace ach;
    ACH_REQUIRE(wt && w_hi && w_lo && K > 0 && O > 0 && ldw >= O, "ach_pack_pw_tc_nt: bad args");
    ACH_REQUIRE(NT >= 16 && NT <= 256 && NT % 16 == 0, "ach_pack_pw_tc_nt: NT=%d must be a multiple of 16 in [16, 256]", NT);
    const long long total = ach_pack_pw_tc_nt_elems(K, O, NT);
    pack_pw_tc_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(wt, K, O, ldw, NT, cdiv(K, TC_KC), w_hi, w_lo, total);
    return check_launch("ach_pack_pw_tc_nt");
}

extern "C" int ach_pw_conv_tc(const AchPwConv* pp, const float* w_hi, const float* w_lo, const float* wsum, void* stream) {
    using namespace ach;
    const AchPwConv& p = *pp;
    ACH_REQUIRE(p.x0 && p.out && w_hi && w_lo, "ach_pw_conv_tc: null x0/out/weights");
    ACH_REQUIRE(p.B > 0 && p.O > 0 && p.P > 0 && p.c0 > 0 && p.c1 >= 0, "ach_pw_conv_tc: bad dims");
    ACH_REQUIRE((p.c1 == 0) == (p.x1 == nullptr), "ach_pw_conv_tc: x1/c1 mismatch");
    ACH_REQUIRE(p.P % 4 == 0, "ach_pw_conv_tc: P=%d must be a multiple of 4", p.P);
    ACH_REQUIRE(aligned16(p.x0) && aligned16(p.x1) && aligned16(w_hi) && aligned16(w_lo), "ach_pw_conv_tc: views must be 16-byte aligned");
    ACH_REQUIRE(p.x0_bs % 4 == 0 && p.x1_bs % 4 == 0, "ach_pw_conv_tc: batch strides must be multiples of 4 elements");
    ACH_REQUIRE(p.wt_bs == 0, "ach_pw_conv_tc: per-frame weights use ach_pw_conv");
    ACH_REQUIRE(!(p.reduce_max && p.res), "ach_pw_conv_tc: reduce_max excludes a residual");
    ACH_REQUIRE(p.B <= 65535, "ach_pw_conv_tc: B too large");
    ACH_REQUIRE(!p.ln || wsum, "ach_pw_conv_tc: the LayerNorm prologue needs wsum (row sums of the folded weights)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return pw_conv_tc_ws_launch(p, w_hi, w_lo, wsum, st);   // warp-specialised kernel, pw_conv_tc_ws.cu
}
