// Probe: which cp.async.bulk.tensor tile shapes work for a (W, H, C, B) fp32 plane stack on this part.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../achelous_b200/csrc/tma_common.cuh"
using namespace ach;

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int n, int c0, int c1, int c2, int c3) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) unsigned long long mbar;
    if (threadIdx.x == 0) {
        tma_mbar_init(tma_smem_u32(&mbar), 1);
        tma_mbar_expect_tx(tma_smem_u32(&mbar), (uint32_t)n * 4u);
        if (RANK == 4)
            tma_load_4d(tma_smem_u32(sm), &tm, c0, c1, c2, c3, tma_smem_u32(&mbar));
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                             tma_smem_u32(sm)), "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(tma_smem_u32(&mbar)) : "memory");
    }
    __syncthreads();
    tma_mbar_wait(tma_smem_u32(&mbar), 0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = sm[i];
}

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    int idx = -1;
    const int W = 160, H = 160, C = 16, B = 2;
    std::vector<float> h((size_t)W * H * C * B);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&o, 1 << 20);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    TmaEncodeTiledFn enc = tma_encode_fn();
    struct Case { int rank, bw, bh, bc; CUtensorMapL2promotion l2; const char* name; };
    Case cases[] = {{4, 24, 24, 16, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "4d 24x24x16 l2-128"}, {4, 24, 24, 16, CU_TENSOR_MAP_L2_PROMOTION_NONE, "4d 24x24x16 l2-none"},
                    {4, 32, 24, 16, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "4d 32x24x16"}, {4, 24, 24, 8, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "4d 24x24x8"},
                    {4, 24, 24, 4, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "4d 24x24x4"}, {4, 24, 12, 16, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "4d 24x12x16"},
                    {3, 24, 24, 16, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "3d 24x24x16 (C*B folded)"}, {3, 24, 24, 8, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "3d 24x24x8"},
                    {3, 32, 24, 16, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "3d 32x24x16"},
                    {4, 36, 33, 3, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "4d 36x33x3"}, {4, 36, 32, 3, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "4d 36x32x3"},
                    {4, 32, 33, 3, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "4d 32x33x3"}, {4, 36, 33, 4, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "4d 36x33x4"},
                    {4, 36, 33, 11, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "4d 36x33x11"}, {4, 20, 18, 16, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "4d 20x18x16"}};
    for (const Case& cs : cases) {
        if (++idx != only && only >= 0) continue;
        alignas(64) CUtensorMap tm;
        memset(&tm, 0, sizeof(tm));
        const cuuint64_t dims4[4] = {W, H, C, B}, dims3[3] = {W, H, (cuuint64_t)C * B};
        const cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
        const cuuint32_t box[4] = {(cuuint32_t)cs.bw, (cuuint32_t)cs.bh, (cuuint32_t)cs.bc, 1};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, cs.rank, d, cs.rank == 4 ? dims4 : dims3, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, cs.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int n = cs.bw * cs.bh * cs.bc;
        if (r != CUDA_SUCCESS) { printf("%-28s encode failed (%d)\n", cs.name, (int)r); continue; }
        cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        const int c0 = argc > 2 ? atoi(argv[2]) : 94, c1 = argc > 3 ? atoi(argv[3]) : -2, c2 = 0, c3 = 1;
        if (cs.rank == 4) probe<4><<<1, 128, n * 4>>>(tm, o, n, c0, c1, c2, c3);
        else probe<3><<<1, 128, n * 4>>>(tm, o, n, c0, c1, C * c3, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-28s %d bytes: LAUNCH ERROR %s\n", cs.name, n * 4, cudaGetErrorString(e)); return 1; }
        std::vector<float> got(n);
        cudaMemcpy(got.data(), o, n * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int c = 0; c < cs.bc; ++c)
            for (int y = 0; y < cs.bh; ++y)
                for (int x = 0; x < cs.bw; ++x) {
                    const int gy = c1 + y, gx = c0 + x;
                    const float want = (gy < 0 || gy >= H || gx < 0 || gx >= W) ? 0.f : h[(((size_t)c3 * C + c) * H + gy) * W + gx];
                    bad += got[(c * cs.bh + y) * cs.bw + x] != want;
                }
        printf("%-28s %d bytes: ok, %d mismatches\n", cs.name, n * 4, bad);
    }
    return 0;
}
