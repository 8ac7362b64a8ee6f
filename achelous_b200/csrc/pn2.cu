// PointNet++ building blocks (builder-defined network, oracle/pn2.py; the reference ships no PN2 code):
// farthest-point sampling, ball query + grouping, group max, 3-NN inverse-distance interpolation.
// Index-producing arithmetic uses explicit round-to-nearest intrinsics in the oracle's order
// ((dx*dx + dy*dy) + dz*dz, no FMA contraction) so sampled / grouped / neighbour indices are bit-identical
// to the CPU oracle.  The per-point MLPs run on the pointwise GEMM kernels.
#include "common.cuh"

namespace ach {

__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---- farthest point sampling: one CTA per frame, points + running min-distance in shared memory.
// Start index 0; arg-max ties -> lowest index (warp shuffle reduction on (value, index) pairs).
__global__ void __launch_bounds__(256) pn2_fps_kernel(const float* __restrict__ xyz, long long xyz_bs, int N, int npoint,
                                                      int* __restrict__ idx_out, float* __restrict__ new_xyz, long long new_bs) {
    extern __shared__ float smem[];
    float* sx = smem; float* sy = sx + N; float* sz = sy + N; float* sd = sz + N;
    __shared__ float red_v[8];
    __shared__ int red_i[8];
    __shared__ int s_far;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* xb = xyz + (long long)b * xyz_bs;
    for (int i = tid; i < N; i += 256) {
        sx[i] = xb[i]; sy[i] = xb[N + i]; sz[i] = xb[2 * N + i];
        sd[i] = 1e10f;
    }
    if (tid == 0) s_far = 0;
    __syncthreads();
    for (int it = 0; it < npoint; ++it) {
        const int far = s_far;
        const float cx = sx[far], cy = sy[far], cz = sz[far];
        if (tid == 0) {
            idx_out[(long long)b * npoint + it] = far;
            float* nb = new_xyz + (long long)b * new_bs;
            nb[it] = cx; nb[npoint + it] = cy; nb[2 * npoint + it] = cz;
        }
        float bv = -1.f;
        int bi = 0x7fffffff;
        for (int i = tid; i < N; i += 256) {
            const float d = fminf(sd[i], sqdist3(cx, cy, cz, sx[i], sy[i], sz[i]));   // oracle: sqdist(c, xyz) = c - x
            sd[i] = d;
            if (d > bv) { bv = d; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        __syncthreads();   // everyone has read s_far
        if ((tid & 31) == 0) { red_v[tid >> 5] = bv; red_i[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            float v = red_v[0];
            int ii = red_i[0];
            for (int w = 1; w < 8; ++w)
                if (red_v[w] > v || (red_v[w] == v && red_i[w] < ii)) { v = red_v[w]; ii = red_i[w]; }
            s_far = ii;
        }
        __syncthreads();
    }
}

// ---- ball query + grouping: one warp per centroid.  Lanes test 32 candidate points at a time in index order,
// a ballot keeps the first `nsample` hits (d2 <= r2), missing slots repeat the first hit.  The group is written as
// out[b][ch][j * nsample + s]: ch < 3 -> xyz[idx] - centroid, else the point's features.
__global__ void __launch_bounds__(256) pn2_group_kernel(const float* __restrict__ xyz, long long xyz_bs, const float* __restrict__ pts,
                                                        long long pts_bs, int C, const float* __restrict__ new_xyz, long long new_bs,
                                                        int N, int S, int nsample, float r2, float* __restrict__ out, long long out_bs,
                                                        int* __restrict__ idx_out) {
    __shared__ int s_idx[8][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + warp;
    const int b = blockIdx.y;
    if (j >= S) return;
    const float* xb = xyz + (long long)b * xyz_bs;
    const float* nb = new_xyz + (long long)b * new_bs;
    const float cx = nb[j], cy = nb[S + j], cz = nb[2 * S + j];
    int found = 0;
    for (int base = 0; base < N && found < nsample; base += 32) {
        const int i = base + lane;
        bool hit = false;
        if (i < N) hit = !(sqdist3(cx, cy, cz, xb[i], xb[N + i], xb[2 * N + i]) > r2);
        unsigned m = __ballot_sync(0xffffffffu, hit);
        while (m && found < nsample) {
            const int bit = __ffs(m) - 1;
            if (lane == 0) s_idx[warp][found] = base + bit;
            ++found;
            m &= m - 1;
        }
    }
    __syncwarp();
    // the centroid itself is a point of the cloud (d2 = 0), so found >= 1
    const int first = s_idx[warp][0];
    for (int s = found + lane; s < nsample; s += 32) s_idx[warp][s] = first;
    __syncwarp();
    const long long P = (long long)S * nsample;
    float* ob = out + (long long)b * out_bs + (long long)j * nsample;
    const float* pb = pts + (long long)b * pts_bs;
    for (int s = lane; s < nsample; s += 32) {
        const int id = s_idx[warp][s];
        if (idx_out) idx_out[((long long)b * S + j) * nsample + s] = id;
        ob[s] = __fsub_rn(xb[id], cx);
        ob[P + s] = __fsub_rn(xb[N + id], cy);
        ob[2 * P + s] = __fsub_rn(xb[2 * N + id], cz);
        for (int c = 0; c < C; ++c) ob[(long long)(3 + c) * P + s] = pb[(long long)c * N + id];
    }
}

__global__ void __launch_bounds__(256) pn2_group_max_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ out,
                                                            long long out_bs, int C, int S, int nsample) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= C * S) return;
    const float* xp = x + (long long)blockIdx.y * x_bs + (long long)i * nsample;   // (c, j) -> c*S*ns + j*ns
    float m = xp[0];
    for (int s = 1; s < nsample; ++s) m = fmaxf(m, xp[s]);
    out[(long long)blockIdx.y * out_bs + i] = m;
}

// ---- feature propagation: 3 nearest sources (ties -> lowest index), weights 1/(d2 + 1e-8) normalised,
// out[b][c][i] = (p[c][i0]*w0 + p[c][i1]*w1) + p[c][i2]*w2
__global__ void __launch_bounds__(128) pn2_interp3_kernel(const float* __restrict__ xyz1, long long xyz1_bs, const float* __restrict__ xyz2,
                                                          long long xyz2_bs, const float* __restrict__ pts2, long long pts2_bs, int C2,
                                                          int N1, int S, float* __restrict__ out, long long out_bs) {
    extern __shared__ float s2[];   // [3][S]
    const int b = blockIdx.y;
    const float* x2 = xyz2 + (long long)b * xyz2_bs;
    for (int i = threadIdx.x; i < 3 * S; i += 128) s2[i] = x2[i];
    __syncthreads();
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= N1) return;
    const float* x1 = xyz1 + (long long)b * xyz1_bs;
    const float px = x1[i], py = x1[N1 + i], pz = x1[2 * N1 + i];
    float d0 = INFINITY, d1 = INFINITY, d2v = INFINITY;
    int i0 = 0, i1 = 0, i2 = 0;
    for (int s = 0; s < S; ++s) {
        const float d = sqdist3(px, py, pz, s2[s], s2[S + s], s2[2 * S + s]);
        if (d < d0) { d2v = d1; i2 = i1; d1 = d0; i1 = i0; d0 = d; i0 = s; }
        else if (d < d1) { d2v = d1; i2 = i1; d1 = d; i1 = s; }
        else if (d < d2v) { d2v = d; i2 = s; }
    }
    const float r0 = __fdiv_rn(1.0f, __fadd_rn(d0, 1e-8f)), r1 = __fdiv_rn(1.0f, __fadd_rn(d1, 1e-8f)), r2 = __fdiv_rn(1.0f, __fadd_rn(d2v, 1e-8f));
    const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
    const float w0 = __fdiv_rn(r0, norm), w1 = __fdiv_rn(r1, norm), w2 = __fdiv_rn(r2, norm);
    const float* pb = pts2 + (long long)b * pts2_bs;
    float* ob = out + (long long)b * out_bs + i;
    for (int c = 0; c < C2; ++c) {
        const float* pc = pb + (long long)c * S;
        ob[(long long)c * N1] = __fadd_rn(__fadd_rn(__fmul_rn(pc[i0], w0), __fmul_rn(pc[i1], w1)), __fmul_rn(pc[i2], w2));
    }
}

}  // namespace ach

extern "C" int ach_pn2_fps(const float* xyz, long long xyz_bs, int B, int N, int npoint, int* idx_out, float* new_xyz,
                           long long new_bs, void* stream) {
    using namespace ach;
    ACH_REQUIRE(xyz && idx_out && new_xyz && B > 0 && B <= 65535 && N > 0 && npoint > 0 && npoint <= N, "ach_pn2_fps: bad args");
    ACH_REQUIRE((size_t)N * 16 <= 96 * 1024, "ach_pn2_fps: N=%d too large for the shared-memory path", N);
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        cudaFuncSetAttribute(pn2_fps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    }
    pn2_fps_kernel<<<B, 256, (size_t)N * 16, (cudaStream_t)stream>>>(xyz, xyz_bs, N, npoint, idx_out, new_xyz, new_bs);
    return check_launch("ach_pn2_fps");
}

extern "C" int ach_pn2_group(const float* xyz, long long xyz_bs, const float* pts, long long pts_bs, int C, const float* new_xyz,
                             long long new_bs, int B, int N, int S, int nsample, float radius, float* out, long long out_bs,
                             int* idx_out, void* stream) {
    using namespace ach;
    ACH_REQUIRE(xyz && pts && new_xyz && out && B > 0 && B <= 65535 && N > 0 && S > 0, "ach_pn2_group: bad args");
    ACH_REQUIRE(nsample > 0 && nsample <= 64, "ach_pn2_group: nsample=%d must be in [1, 64]", nsample);
    const float r2 = __builtin_powif(radius, 2);   // float32(radius) ** 2, as the oracle
    pn2_group_kernel<<<dim3(cdiv(S, 8), B), 256, 0, (cudaStream_t)stream>>>(xyz, xyz_bs, pts, pts_bs, C, new_xyz, new_bs, N, S, nsample, r2,
                                                                          out, out_bs, idx_out);
    return check_launch("ach_pn2_group");
}

extern "C" int ach_pn2_group_max(const float* x, long long x_bs, float* out, long long out_bs, int B, int C, int S, int nsample,
                                 void* stream) {
    using namespace ach;
    ACH_REQUIRE(x && out && B > 0 && B <= 65535 && C > 0 && S > 0 && nsample > 0, "ach_pn2_group_max: bad args");
    pn2_group_max_kernel<<<dim3(cdiv((long long)C * S, 256), B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, out, out_bs, C, S, nsample);
    return check_launch("ach_pn2_group_max");
}

extern "C" int ach_pn2_interp3(const float* xyz1, long long xyz1_bs, const float* xyz2, long long xyz2_bs, const float* pts2,
                               long long pts2_bs, int B, int C2, int N1, int S, float* out, long long out_bs, void* stream) {
    using namespace ach;
    ACH_REQUIRE(xyz1 && xyz2 && pts2 && out && B > 0 && B <= 65535 && C2 > 0 && N1 > 0 && S >= 3, "ach_pn2_interp3: bad args (S >= 3)");
    ACH_REQUIRE((size_t)S * 12 <= 48 * 1024, "ach_pn2_interp3: S too large");
    pn2_interp3_kernel<<<dim3(cdiv(N1, 128), B), 128, (size_t)S * 12, (cudaStream_t)stream>>>(xyz1, xyz1_bs, xyz2, xyz2_bs, pts2, pts2_bs, C2,
                                                                                            N1, S, out, out_bs);
    return check_launch("ach_pn2_interp3");
}
