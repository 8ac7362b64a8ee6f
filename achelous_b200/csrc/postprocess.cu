// Detection post-process on device: YOLOX box decode (utils/utils_bbox.py:33-85) and class-aware
// greedy NMS (utils/utils_bbox.py:87-130 -> torchvision batched_nms, coordinate-trick branch).
//
// ach_nms runs one CTA per image and is a single launch for the whole batch - the reference loops over
// images in Python and calls torchvision per image.  Inside the CTA:
//   A. class max / first-argmax, score = obj * cls, threshold, ORDER-PRESERVING compaction (ballot scan)
//   B. max coordinate over the candidates -> per-class box shift (the "coordinate trick")
//   C. stable descending rank by counting (ties -> lower candidate index first, = torch stable sort)
//   D. greedy sweep in rank order; every surviving box suppresses in parallel, one barrier per KEPT box
//   E. order-preserving compaction of the survivors into (x1,y1,x2,y2,obj,cls_conf,cls) rows.
// IoU arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction) so kept indices are
// bit-identical to the fp32 CPU oracle.
#include "common.cuh"

namespace ach {

struct DecodeLevels {
    const float* ptr[4];
    long long bs[4];
    int h[4], w[4], a0[4];
    int n;
};

__global__ void __launch_bounds__(256) decode_kernel(DecodeLevels L, float* __restrict__ out, int A, int CH, float input_h,
                                                     float input_w) {
    const int a = blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.y;
    if (a >= A) return;
    int l = 0;
    while (l + 1 < L.n && a >= L.a0[l + 1]) ++l;
    const int cell = a - L.a0[l];
    const int h = L.h[l], w = L.w[l];
    const int gy = cell / w, gx = cell - gy * w;
    const long long plane = (long long)h * w;
    const float* src = L.ptr[l] + (long long)b * L.bs[l] + cell;
    float* dst = out + ((long long)b * A + a) * CH;
    const float stride = input_h / (float)h;  // Python float input_shape[0] / h, then cast to fp32
    // xy = (v + grid) * stride ; wh = exp(v) * stride ; then / input size.  No FMA contraction possible here.
    dst[0] = __fdiv_rn(__fmul_rn(__fadd_rn(src[0], (float)gx), stride), input_w);
    dst[1] = __fdiv_rn(__fmul_rn(__fadd_rn(src[plane], (float)gy), stride), input_h);
    dst[2] = __fdiv_rn(__fmul_rn(expf(src[2 * plane]), stride), input_w);
    dst[3] = __fdiv_rn(__fmul_rn(expf(src[3 * plane]), stride), input_h);
    for (int c = 4; c < CH; ++c) dst[c] = sigmoidf_(src[(long long)c * plane]);
}

constexpr int NMS_T = 512;
constexpr int NMS_MAX_A = 16384;  // shared `removed` bytes

// exclusive block scan of a 0/1 flag (ballot based); returns position, adds the block total to *running
__device__ __forceinline__ int block_scan_flag(bool flag, int* warp_tot, int* running) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    const int pre = __popc(m & ((1u << lane) - 1u));
    __syncthreads();  // previous use of warp_tot / running is complete
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int base = *running;
    for (int i = 0; i < warp; ++i) base += warp_tot[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < NMS_T / 32; ++i) t += warp_tot[i];
        *running += t;
    }
    return base + pre;
}

__global__ void __launch_bounds__(NMS_T) nms_kernel(const float* __restrict__ decoded, int A, int K, float conf_thres,
                                                    float nms_thres, float* __restrict__ kept, long long kept_bs, int max_keep,
                                                    int* __restrict__ kept_idx, int* __restrict__ counts, long long counts_bs,
                                                    char* __restrict__ workspace, long long ws_per_img) {
    extern __shared__ unsigned char removed[];  // [A]
    __shared__ int warp_tot[NMS_T / 32];
    __shared__ int running;
    __shared__ float red[NMS_T / 32];
    __shared__ float s_maxc;

    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int CH = 5 + K;
    const float* pred = decoded + (long long)b * A * CH;
    char* ws = workspace + (long long)b * ws_per_img;
    const int Ap = (A + 3) & ~3;
    float4* c_box = reinterpret_cast<float4*>(ws);            // [Ap] unshifted boxes, candidate (anchor) order
    float4* s_box = c_box + Ap;                               // [Ap] class-shifted boxes, rank order
    float* c_score = reinterpret_cast<float*>(s_box + Ap);    // [Ap] candidate scores
    float* s_area = c_score + Ap;                             // [Ap] areas, rank order
    int* c_anchor = reinterpret_cast<int*>(s_area + Ap);      // [Ap]
    int* c_cls = c_anchor + Ap;                               // [Ap]
    int* order = c_cls + Ap;                                  // [Ap] rank -> candidate slot

    if (tid == 0) running = 0;
    __syncthreads();

    // ---- A. candidates
    float local_max = -INFINITY;
    for (int a0 = 0; a0 < A; a0 += NMS_T) {
        const int a = a0 + tid;
        bool flag = false;
        float score = 0.f;
        int cls = 0;
        float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
        if (a < A) {
            const float* r = pred + (long long)a * CH;
            float conf = r[5];
            for (int c = 1; c < K; ++c) {
                const float v = r[5 + c];
                if (v > conf) {
                    conf = v;
                    cls = c;
                }
            }
            score = __fmul_rn(r[4], conf);
            flag = score >= conf_thres;
            if (flag) {
                const float hw = __fmul_rn(r[2], 0.5f), hh = __fmul_rn(r[3], 0.5f);
                x1 = __fsub_rn(r[0], hw);
                y1 = __fsub_rn(r[1], hh);
                x2 = __fadd_rn(r[0], hw);
                y2 = __fadd_rn(r[1], hh);
                local_max = fmaxf(local_max, fmaxf(fmaxf(x1, y1), fmaxf(x2, y2)));
            }
        }
        const int pos = block_scan_flag(flag, warp_tot, &running);
        if (flag) {
            c_score[pos] = score;
            c_anchor[pos] = a;
            c_cls[pos] = cls;
            c_box[pos] = make_float4(x1, y1, x2, y2);
        }
    }
    __syncthreads();
    const int n = running;
    local_max = warp_max(local_max);
    if ((tid & 31) == 0) red[tid >> 5] = local_max;
    __syncthreads();
    if (tid == 0) {
        float m = -INFINITY;
        for (int i = 0; i < NMS_T / 32; ++i) m = fmaxf(m, red[i]);
        s_maxc = m;
    }
    __syncthreads();
    if (n == 0) {
        if (tid == 0) counts[(long long)b * counts_bs] = 0;
        if (max_keep < A)   // compact rows: deterministic zeros behind the kept rows
            for (int i = tid; i < max_keep * 7; i += NMS_T) kept[(long long)b * kept_bs + i] = 0.f;
        return;
    }
    const float shift = __fadd_rn(s_maxc, 1.0f);

    // ---- C. stable descending rank
    for (int s = tid; s < n; s += NMS_T) {
        const float sc = c_score[s];
        int rank = 0;
        for (int t = 0; t < n; ++t) {
            const float st = c_score[t];
            rank += (st > sc) || (st == sc && t < s);
        }
        order[rank] = s;
    }
    __syncthreads();
    // ---- B. class-shifted boxes + areas in rank order
    for (int r = tid; r < n; r += NMS_T) {
        const int s = order[r];
        const float off = __fmul_rn((float)c_cls[s], shift);
        const float4 u = c_box[s];
        const float4 v = make_float4(__fadd_rn(u.x, off), __fadd_rn(u.y, off), __fadd_rn(u.z, off), __fadd_rn(u.w, off));
        s_box[r] = v;
        s_area[r] = __fmul_rn(__fsub_rn(v.z, v.x), __fsub_rn(v.w, v.y));
        removed[r] = 0;
    }
    __syncthreads();

    // ---- D. greedy sweep in rank order, 32 boxes (one warp's worth) per round.  Invariant at the start of a round: every box of
    // the chunk has already been tested against every box kept in earlier chunks.  Warp 0 then resolves the chunk sequentially
    // (box k, if it survives, suppresses the later boxes of the chunk: shuffles, no block barrier); all threads apply the chunk's
    // survivors to the boxes behind the chunk.  Same decisions as the one-box-at-a-time sweep with n/32 instead of n barriers
    // (the per-box version cost 0.71 ms for 64 images, one block barrier per kept box).
    __shared__ float4 k_box[32];
    __shared__ float k_area[32];
    __shared__ int k_n;
    for (int c0 = 0; c0 < n; c0 += 32) {
        if (tid < 32) {
            const int r = c0 + tid;
            const bool valid = r < n;
            bool dead = valid ? removed[r] != 0 : true;
            const float4 bx = valid ? s_box[r] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float ar = valid ? s_area[r] : 0.f;
            for (int k = 0; k < 32; ++k) {
                const unsigned alive = ~__ballot_sync(0xffffffffu, dead);
                if (!((alive >> k) & 1u)) continue;            // warp-uniform
                const float4 bk = make_float4(__shfl_sync(0xffffffffu, bx.x, k), __shfl_sync(0xffffffffu, bx.y, k),
                                              __shfl_sync(0xffffffffu, bx.z, k), __shfl_sync(0xffffffffu, bx.w, k));
                const float ak = __shfl_sync(0xffffffffu, ar, k);
                if (tid > k && !dead) {
                    const float w = fmaxf(__fsub_rn(fminf(bk.z, bx.z), fmaxf(bk.x, bx.x)), 0.f);
                    const float h = fmaxf(__fsub_rn(fminf(bk.w, bx.w), fmaxf(bk.y, bx.y)), 0.f);
                    const float inter = __fmul_rn(w, h);
                    if (inter > 0.f) {   // disjoint pair (every cross-class pair after the shift): IoU = 0 never exceeds a threshold >= 0
                        const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ak, ar), inter));
                        if (iou > nms_thres) dead = true;
                    }
                }
            }
            if (valid) removed[r] = dead ? 1 : 0;
            const unsigned alive = ~__ballot_sync(0xffffffffu, dead);
            if (!dead) {
                const int slot = __popc(alive & ((1u << tid) - 1u));
                k_box[slot] = bx;
                k_area[slot] = ar;
            }
            if (tid == 0) k_n = __popc(alive);
        }
        __syncthreads();
        const int kn = k_n;
        for (int j = c0 + 32 + tid; j < n; j += NMS_T) {
            if (removed[j]) continue;
            const float4 bj = s_box[j];
            const float aj = s_area[j];
            for (int k = 0; k < kn; ++k) {
                const float4 bi = k_box[k];
                const float w = fmaxf(__fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)), 0.f);
                const float h = fmaxf(__fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)), 0.f);
                const float inter = __fmul_rn(w, h);
                if (inter > 0.f) {
                    const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(k_area[k], aj), inter));
                    if (iou > nms_thres) {
                        removed[j] = 1;
                        break;
                    }
                }
            }
        }
        __syncthreads();
    }

    // ---- E. survivors, rank order
    if (tid == 0) running = 0;
    __syncthreads();
    for (int r0 = 0; r0 < n; r0 += NMS_T) {
        const int r = r0 + tid;
        const bool flag = (r < n) && !removed[r];
        const int pos = block_scan_flag(flag, warp_tot, &running);
        if (flag && pos < max_keep) {
            const int s = order[r];
            const int a = c_anchor[s];
            const float* row = pred + (long long)a * CH;
            const float hw = __fmul_rn(row[2], 0.5f), hh = __fmul_rn(row[3], 0.5f);
            float* o = kept + (long long)b * kept_bs + (long long)pos * 7;
            o[0] = __fsub_rn(row[0], hw);
            o[1] = __fsub_rn(row[1], hh);
            o[2] = __fadd_rn(row[0], hw);
            o[3] = __fadd_rn(row[1], hh);
            o[4] = row[4];
            o[5] = row[5 + c_cls[s]];
            o[6] = (float)c_cls[s];
            if (kept_idx) kept_idx[(long long)b * A + pos] = a;
        }
    }
    __syncthreads();
    if (tid == 0) counts[(long long)b * counts_bs] = running;   // the TRUE number of survivors (may exceed max_keep)
    if (max_keep < A)
        for (int i = min(running, max_keep) * 7 + tid; i < max_keep * 7; i += NMS_T) kept[(long long)b * kept_bs + i] = 0.f;
}

static long long nms_ws_per_img(int A) {
    const long long Ap = (A + 3) & ~3LL;
    return Ap * (16 + 16 + 4 + 4 + 4 + 4 + 4);
}

}  // namespace ach

extern "C" int ach_decode_outputs(const float* const* levels, const long long* level_bs, const int* hs, const int* ws,
                                  int n_levels, float* out, int B, int K, float input_h, float input_w, void* stream) {
    using namespace ach;
    ACH_REQUIRE(levels && level_bs && hs && ws && out, "ach_decode_outputs: null arg");
    ACH_REQUIRE(n_levels >= 1 && n_levels <= 4 && B > 0 && B <= 65535 && K >= 1, "ach_decode_outputs: bad dims");
    DecodeLevels L;
    int A = 0;
    for (int i = 0; i < n_levels; ++i) {
        ACH_REQUIRE(levels[i] && hs[i] > 0 && ws[i] > 0, "ach_decode_outputs: bad level %d", i);
        L.ptr[i] = levels[i];
        L.bs[i] = level_bs[i];
        L.h[i] = hs[i];
        L.w[i] = ws[i];
        L.a0[i] = A;
        A += hs[i] * ws[i];
    }
    L.n = n_levels;
    decode_kernel<<<dim3(cdiv(A, 256), B), 256, 0, (cudaStream_t)stream>>>(L, out, A, 5 + K, input_h, input_w);
    return check_launch("ach_decode_outputs");
}

extern "C" long long ach_nms_workspace_bytes(int B, int A) { return (long long)B * ach::nms_ws_per_img(A); }

extern "C" int ach_nms(const float* decoded, int B, int A, int K, float conf_thres, float nms_thres, float* kept,
                       int* kept_idx, int* counts, void* workspace, long long workspace_bytes, void* stream) {
    using namespace ach;
    ACH_REQUIRE(decoded && kept && kept_idx && counts && workspace, "ach_nms: null arg");
    ACH_REQUIRE(B > 0 && A > 0 && K >= 1, "ach_nms: bad dims");
    ACH_REQUIRE(nms_thres >= 0.f, "ach_nms: nms_thres must be >= 0");
    ACH_REQUIRE(A <= NMS_MAX_A, "ach_nms: A=%d anchors per image exceeds the supported %d", A, NMS_MAX_A);
    ACH_REQUIRE(workspace_bytes >= ach_nms_workspace_bytes(B, A), "ach_nms: workspace too small");
    ACH_REQUIRE(aligned16(workspace), "ach_nms: workspace must be 16-byte aligned");
    nms_kernel<<<B, NMS_T, (size_t)((A + 15) & ~15), (cudaStream_t)stream>>>(decoded, A, K, conf_thres, nms_thres, kept, 7LL * A, A, kept_idx,
                                                                           counts, 1, static_cast<char*>(workspace), nms_ws_per_img(A));
    return check_launch("ach_nms");
}

// Same NMS writing at most max_keep rows per image (score-descending, so the cap keeps the best ones) at a caller-chosen
// row-block stride, with the true survivor count next to them: the detection part of the compact output record.
extern "C" int ach_nms_rows(const float* decoded, int B, int A, int K, float conf_thres, float nms_thres, float* kept, long long kept_bs,
                            int max_keep, int* counts, long long counts_bs, void* workspace, long long workspace_bytes, void* stream) {
    using namespace ach;
    ACH_REQUIRE(decoded && kept && counts && workspace, "ach_nms_rows: null arg");
    ACH_REQUIRE(B > 0 && A > 0 && K >= 1 && max_keep > 0 && max_keep <= A, "ach_nms_rows: bad dims");
    ACH_REQUIRE(nms_thres >= 0.f, "ach_nms_rows: nms_thres must be >= 0");
    ACH_REQUIRE(kept_bs >= 7LL * max_keep && counts_bs >= 1, "ach_nms_rows: strides smaller than one image's record");
    ACH_REQUIRE(A <= NMS_MAX_A, "ach_nms_rows: A=%d anchors per image exceeds the supported %d", A, NMS_MAX_A);
    ACH_REQUIRE(workspace_bytes >= ach_nms_workspace_bytes(B, A), "ach_nms_rows: workspace too small");
    ACH_REQUIRE(aligned16(workspace), "ach_nms_rows: workspace must be 16-byte aligned");
    nms_kernel<<<B, NMS_T, (size_t)((A + 15) & ~15), (cudaStream_t)stream>>>(decoded, A, K, conf_thres, nms_thres, kept, kept_bs, max_keep,
                                                                           nullptr, counts, counts_bs, static_cast<char*>(workspace),
                                                                           nms_ws_per_img(A));
    return check_launch("ach_nms_rows");
}

// ------------------------------------------------------------------------------------------------
// Segmentation post-process on device (reference achelous.py:283-318 does it on the host):
//   softmax over classes -> letterbox crop -> cv2.resize(INTER_LINEAR) to the original image size -> argmax.
// ach_seg_softmax writes class probabilities at network resolution; ach_seg_resize_argmax evaluates, per
// destination pixel, OpenCV's half-pixel-centre bilinear rule (fx = (dx + 0.5) * scale - 0.5, indices clamped with
// the weight forced to 0 at the borders, horizontal pass then vertical pass in fp32) on the cropped window and writes
// the first-max class index as uint8: 1 byte per original pixel leaves the GPU instead of 4*K bytes per network pixel.
namespace ach {

__global__ void __launch_bounds__(256) seg_softmax_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ out,
                                                          long long out_bs, int K, int P) {
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= P) return;
    const float* xp = x + (long long)blockIdx.y * x_bs + p;
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) m = fmaxf(m, xp[(long long)k * P]);
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += expf(xp[(long long)k * P] - m);
    float* op = out + (long long)blockIdx.y * out_bs + p;
    for (int k = 0; k < K; ++k) op[(long long)k * P] = expf(xp[(long long)k * P] - m) / s;
}

__device__ __forceinline__ void cv_linear_coord(int d, double scale, int ssize, int& s0, int& s1, float& a0, float& a1) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
    s0 = s;
    s1 = min(s + 1, ssize - 1);
    a0 = 1.f - f;
    a1 = f;
}

__global__ void __launch_bounds__(256) seg_resize_argmax_kernel(const float* __restrict__ prob, long long prob_bs, int K, int H, int W,
                                                                int y_off, int x_off, int nh, int nw, unsigned char* __restrict__ out,
                                                                int OH, int OW) {
    const int ox = blockIdx.x * 256 + threadIdx.x;
    const int oy = blockIdx.y, b = blockIdx.z;
    if (ox >= OW) return;
    int sx0, sx1, sy0, sy1;
    float ax0, ax1, ay0, ay1;
    cv_linear_coord(ox, (double)nw / (double)OW, nw, sx0, sx1, ax0, ax1);
    cv_linear_coord(oy, (double)nh / (double)OH, nh, sy0, sy1, ay0, ay1);
    const long long P = (long long)H * W;
    const float* pb = prob + (long long)b * prob_bs;
    const long long r0 = (long long)(y_off + sy0) * W + x_off, r1 = (long long)(y_off + sy1) * W + x_off;
    float best = -INFINITY;
    int arg = 0;
    for (int k = 0; k < K; ++k) {
        const float* pk = pb + (long long)k * P;
        const float h0 = __fadd_rn(__fmul_rn(pk[r0 + sx0], ax0), __fmul_rn(pk[r0 + sx1], ax1));   // horizontal pass, row 0
        const float h1 = __fadd_rn(__fmul_rn(pk[r1 + sx0], ax0), __fmul_rn(pk[r1 + sx1], ax1));   // horizontal pass, row 1
        const float v = __fadd_rn(__fmul_rn(h0, ay0), __fmul_rn(h1, ay1));                         // vertical pass
        if (v > best) { best = v; arg = k; }
    }
    out[((long long)b * OH + oy) * OW + ox] = (unsigned char)arg;
}

// argmax over the class planes at network resolution, 4 pixels per thread (one 16-byte load per class, one 4-byte store);
// first maximum like torch.argmax; classes whose keep_mask bit is clear map to 0 (achelous.py:297)
__global__ void __launch_bounds__(256) seg_argmax_u8_kernel(const float* __restrict__ x, long long x_bs, int K, int P, unsigned keep_mask,
                                                            unsigned char* __restrict__ out, long long out_bs) {
    const int p4 = blockIdx.x * 256 + threadIdx.x;
    if (p4 * 4 >= P) return;
    const float4* xp = reinterpret_cast<const float4*>(x + (long long)blockIdx.y * x_bs) + p4;
    float4 best = xp[0];
    int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int k = 1; k < K; ++k) {
        const float4 v = xp[(long long)k * (P / 4)];
        if (v.x > best.x) { best.x = v.x; a0 = k; }
        if (v.y > best.y) { best.y = v.y; a1 = k; }
        if (v.z > best.z) { best.z = v.z; a2 = k; }
        if (v.w > best.w) { best.w = v.w; a3 = k; }
    }
    a0 = ((keep_mask >> a0) & 1u) ? a0 : 0;
    a1 = ((keep_mask >> a1) & 1u) ? a1 : 0;
    a2 = ((keep_mask >> a2) & 1u) ? a2 : 0;
    a3 = ((keep_mask >> a3) & 1u) ? a3 : 0;
    reinterpret_cast<unsigned*>(out + (long long)blockIdx.y * out_bs)[p4] = (unsigned)a0 | ((unsigned)a1 << 8) | ((unsigned)a2 << 16) | ((unsigned)a3 << 24);
}

// softmax + crop + cv2 INTER_LINEAR resize + argmax in ONE pass over the logits: the probabilities of the four source pixels are
// rebuilt in registers with exactly seg_softmax_kernel's arithmetic (max, sum of expf in class order, expf / sum), so the class
// map equals the two-kernel sequence bit for bit while the (K, H, W) fp32 probability map never exists in HBM.
template <int KMAX>
__global__ void __launch_bounds__(256) seg_softmax_resize_argmax_kernel(const float* __restrict__ logits, long long bs, int K, int H, int W,
                                                                        int y_off, int x_off, int nh, int nw, unsigned char* __restrict__ out,
                                                                        long long out_bs, int OH, int OW, unsigned keep_mask) {
    const int ox = blockIdx.x * 256 + threadIdx.x;
    const int oy = blockIdx.y, b = blockIdx.z;
    if (ox >= OW) return;
    int sx0, sx1, sy0, sy1;
    float ax0, ax1, ay0, ay1;
    cv_linear_coord(ox, (double)nw / (double)OW, nw, sx0, sx1, ax0, ax1);
    cv_linear_coord(oy, (double)nh / (double)OH, nh, sy0, sy1, ay0, ay1);
    const long long P = (long long)H * W;
    const float* pb = logits + (long long)b * bs;
    const long long off[4] = {(long long)(y_off + sy0) * W + x_off + sx0, (long long)(y_off + sy0) * W + x_off + sx1,
                              (long long)(y_off + sy1) * W + x_off + sx0, (long long)(y_off + sy1) * W + x_off + sx1};
    float pr[4][KMAX];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) {
                pr[c][k] = __ldg(pb + (long long)k * P + off[c]);
                m = fmaxf(m, pr[c][k]);
            }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) s += expf(pr[c][k] - m);
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) pr[c][k] = expf(pr[c][k] - m) / s;
    }
    float best = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
        if (k < K) {
            const float h0 = __fadd_rn(__fmul_rn(pr[0][k], ax0), __fmul_rn(pr[1][k], ax1));
            const float h1 = __fadd_rn(__fmul_rn(pr[2][k], ax0), __fmul_rn(pr[3][k], ax1));
            const float v = __fadd_rn(__fmul_rn(h0, ay0), __fmul_rn(h1, ay1));
            if (v > best) { best = v; arg = k; }
        }
    out[(long long)b * out_bs + (long long)oy * OW + ox] = (unsigned char)(((keep_mask >> arg) & 1u) ? arg : 0);
}

// per point: log_softmax over the classes exactly as logsoftmax_t_kernel (misc.cu) computes it, then the first maximum of those
// values - the class torch.argmax returns on the raw path's output (achelous.py:262) - as one byte
__global__ void __launch_bounds__(256) logsoftmax_argmax_t_kernel(const float* __restrict__ x, long long x_bs, unsigned char* __restrict__ out,
                                                                  long long out_bs, int K, int N) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= N) return;
    const float* xp = x + (long long)blockIdx.y * x_bs + n;
    float v[32];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 32; ++k)
        if (k < K) {
            v[k] = xp[(long long)k * N];
            m = fmaxf(m, v[k]);
        }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k)
        if (k < K) s += expf(v[k] - m);
    const float lse = logf(s);
    float best = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k)
        if (k < K) {
            const float lp = (v[k] - m) - lse;
            if (lp > best) { best = lp; arg = k; }
        }
    out[(long long)blockIdx.y * out_bs + n] = (unsigned char)arg;
}

}  // namespace ach

extern "C" int ach_seg_argmax_u8(const float* x, long long x_bs, int B, int K, int P, unsigned keep_mask, unsigned char* out,
                                 long long out_bs, void* stream) {
    using namespace ach;
    ACH_REQUIRE(x && out && B > 0 && B <= 65535 && K > 0 && K <= 32 && P > 0, "ach_seg_argmax_u8: bad args");
    ACH_REQUIRE(P % 4 == 0 && x_bs % 4 == 0 && out_bs % 4 == 0 && aligned16(x) && (reinterpret_cast<uintptr_t>(out) & 3u) == 0,
                "ach_seg_argmax_u8: P and the batch strides must be multiples of 4, x 16-byte and out 4-byte aligned");
    seg_argmax_u8_kernel<<<dim3(cdiv(P / 4, 256), B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, K, P, keep_mask, out, out_bs);
    return check_launch("ach_seg_argmax_u8");
}

extern "C" int ach_seg_softmax_resize_argmax(const float* logits, long long bs, int B, int K, int H, int W, int y_off, int x_off, int nh,
                                             int nw, unsigned char* out, long long out_bs, int OH, int OW, unsigned keep_mask, void* stream) {
    using namespace ach;
    ACH_REQUIRE(logits && out && B > 0 && B <= 65535 && K > 0 && K <= 16, "ach_seg_softmax_resize_argmax: bad args (K=%d must be <= 16)", K);
    ACH_REQUIRE(nh > 0 && nw > 0 && y_off >= 0 && x_off >= 0 && y_off + nh <= H && x_off + nw <= W,
                "ach_seg_softmax_resize_argmax: crop window outside the map");
    ACH_REQUIRE(OH > 0 && OH <= 65535 && OW > 0 && out_bs >= (long long)OH * OW, "ach_seg_softmax_resize_argmax: bad output size");
    const dim3 grid(cdiv(OW, 256), OH, B);
    if (K <= 2)
        seg_softmax_resize_argmax_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(logits, bs, K, H, W, y_off, x_off, nh, nw, out, out_bs, OH, OW, keep_mask);
    else if (K <= 9)
        seg_softmax_resize_argmax_kernel<9><<<grid, 256, 0, (cudaStream_t)stream>>>(logits, bs, K, H, W, y_off, x_off, nh, nw, out, out_bs, OH, OW, keep_mask);
    else
        seg_softmax_resize_argmax_kernel<16><<<grid, 256, 0, (cudaStream_t)stream>>>(logits, bs, K, H, W, y_off, x_off, nh, nw, out, out_bs, OH, OW, keep_mask);
    return check_launch("ach_seg_softmax_resize_argmax");
}

extern "C" int ach_logsoftmax_argmax_t(const float* x, long long x_bs, unsigned char* out, long long out_bs, int B, int K, int N, void* stream) {
    using namespace ach;
    ACH_REQUIRE(x && out && B > 0 && K > 0 && K <= 32 && N > 0 && B <= 65535, "ach_logsoftmax_argmax_t: bad args (K=%d must be <= 32)", K);
    logsoftmax_argmax_t_kernel<<<dim3(cdiv(N, 256), B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, out, out_bs, K, N);
    return check_launch("ach_logsoftmax_argmax_t");
}

extern "C" int ach_seg_softmax(const float* x, long long x_bs, float* out, long long out_bs, int B, int K, int P, void* stream) {
    using namespace ach;
    ACH_REQUIRE(x && out && B > 0 && B <= 65535 && K > 0 && P > 0, "ach_seg_softmax: bad args");
    seg_softmax_kernel<<<dim3(cdiv(P, 256), B), 256, 0, (cudaStream_t)stream>>>(x, x_bs, out, out_bs, K, P);
    return check_launch("ach_seg_softmax");
}

extern "C" int ach_seg_resize_argmax(const float* prob, long long prob_bs, int B, int K, int H, int W, int y_off, int x_off, int nh,
                                     int nw, unsigned char* out, int OH, int OW, void* stream) {
    using namespace ach;
    ACH_REQUIRE(prob && out && B > 0 && B <= 65535 && K > 0 && K <= 255, "ach_seg_resize_argmax: bad args");
    ACH_REQUIRE(nh > 0 && nw > 0 && y_off >= 0 && x_off >= 0 && y_off + nh <= H && x_off + nw <= W, "ach_seg_resize_argmax: crop window outside the map");
    ACH_REQUIRE(OH > 0 && OH <= 65535 && OW > 0, "ach_seg_resize_argmax: bad output size");
    seg_resize_argmax_kernel<<<dim3(cdiv(OW, 256), OH, B), 256, 0, (cudaStream_t)stream>>>(prob, prob_bs, K, H, W, y_off, x_off, nh, nw, out, OH, OW);
    return check_launch("ach_seg_resize_argmax");
}
