/*
 * achelous_b200 - C ABI of the sm_100a kernels behind the Achelous 5-task forward.
 *
 * The reference (GuanRunwei/Achelous) has no FFI: its hot path is a torch.nn.Module
 * (nets/Achelous.py:49-53) whose native work is done by ATen/cuDNN/cuBLAS and
 * torchvision._C.  This header is the boundary a maintainer binds instead of those
 * library calls (ctypes stub in INTEGRATION.md; achelous_b200/_lib.py is that stub).
 *
 * Conventions
 *   - every tensor is fp32; activations are "views" (pointer, batch stride in elements) over
 *     channel-major planes: element (b, c, p) lives at ptr[b * bs + c * P + p], P = H * W
 *     (exactly torch NCHW / (B, C, N) layout, so channel slices and concatenations are free);
 *   - every entry point is stream-ordered on `stream` (a cudaStream_t), allocates nothing,
 *     is re-entrant and CUDA-graph capturable;
 *   - return value: 0 = ok, otherwise an ACH_ERR_* code; ach_last_error() returns the
 *     thread-local message.  Nothing throws across this boundary.
 *   - "folded" scale/bias: eval-mode BatchNorm (and conv bias) folded by the host into
 *     y = scale[o] * acc + bias[o]; NULL scale means 1, NULL bias means 0.
 */
#ifndef ACHELOUS_B200_H
#define ACHELOUS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ACH_API __attribute__((visibility("default")))
#else
#define ACH_API
#endif

#define ACH_OK 0
#define ACH_ERR_INVALID 1
#define ACH_ERR_CUDA 2

#define ACH_ACT_NONE 0
#define ACH_ACT_RELU 1
#define ACH_ACT_SILU 2
#define ACH_ACT_GELU 3
#define ACH_ACT_SIGMOID 4

ACH_API const char* ach_last_error(void);
ACH_API int ach_version(void);

/* ---------------------------------------------------------------------------------------------
 * Pointwise (1x1) convolution / Linear / Conv1d(k=1) / bmm as one GEMM per frame:
 *   out[b, o, p] = res[b, o, p] + gamma[o] * act(scale[o] * (sum_k wt[b][k, o] * xn[b, k, p] + pbias[b, o]) + bias[o])
 * x = concat(x0 (c0 channels), x1 (c1 channels)) along channels (x1 may be NULL);
 * xn = x, or LayerNorm over channels without affine when ln != 0 (affine folded into wt/bias by the host).
 * wt is K-major [K][ldw] (ldw % 4 == 0, zero padded); wt_bs != 0 selects per-frame weights.
 * reduce_max != 0: out is (B, O) = max over p (must be pre-filled with -inf; res/gamma unused).
 * Replaces: nn.Conv2d 1x1 (+BN+act) ghost_conv.py:13-17, normal_conv.py:36-49, spp.py:27-35,
 *   decouplehead.py:37-57; nn.Linear in conv_encoder.py:11-13, sdta_encoder.py:25,33,155-159
 *   (with F.layer_norm layers.py:20); Conv1d in pointnet_utils.py:13-15,92-94 and
 *   pointnet_sem_seg.py:18-21; torch.bmm pointnet_utils.py:110,119; torch.max(x, 2) :36,76,127.
 * Requires P % 4 == 0 and 16-byte aligned views.
 */
typedef struct AchPwConv {
    const float* x0;
    const float* x1;
    const float* wt;
    const float* scale;
    const float* bias;
    const float* pbias;   /* (B, O) added before scale, or NULL */
    const float* res;     /* residual view (B, O, P) or NULL */
    const float* gamma;   /* per-output layer-scale applied before the residual add, or NULL */
    float* out;
    long long x0_bs, x1_bs, wt_bs, res_bs, out_bs;
    int c0, c1, ldw;
    int B, O, P;
    int ln;
    float ln_eps;
    int act;
    int reduce_max;
} AchPwConv;
ACH_API int ach_pw_conv(const AchPwConv* p, void* stream);

/* Tensor-core (tcgen05 + TMEM, 3xTF32 split = fp32-accurate) version of ach_pw_conv for shared weights
 * (wt_bs == 0) without reduce_max.  Same AchPwConv contract except that the weights come pre-packed:
 * ach_pack_pw_tc converts K-major wt [K][ldw] into two UMMA shared-memory tile images (tf32 "hi" part and
 * fp32 residual "lo" part), each ach_pack_pw_tc_elems(K, O) floats.  AchPwConv.wt / ldw are ignored.
 * With the LayerNorm prologue (ln != 0) the kernel multiplies the RAW activations and applies
 * LN(x).w = rstd * (x.w - mean * wsum[o]) in the epilogue: wsum (O floats) = sum_k wt[k][o], else NULL. */
ACH_API long long ach_pack_pw_tc_elems(int K, int O);
ACH_API int ach_pack_pw_tc(const float* wt, int K, int O, int ldw, float* w_hi, float* w_lo, void* stream);
ACH_API int ach_pw_conv_tc(const AchPwConv* p, const float* w_hi, const float* w_lo, const float* wsum, void* stream);

/* Fused inverted-bottleneck MLP of the EdgeNeXt encoders on tcgen05 (3xTF32), one launch per block:
 *   out[b, :, p] = res[b, :, p] + gamma * (W2 . gelu(LN(x[b, :, p]) . W1 + b1) + b2)
 * x, res, out are (B, C, P) views (x 16-byte aligned, P % 4 == 0); the 4C-wide hidden activations exist only in tensor memory.
 * LayerNorm over channels without affine (its affine is folded into W1 / b1 by the host; wsum1[n] = sum_k W1[k][n] of the
 * folded weights, as for ach_pw_conv_tc).  Weights pre-packed by ach_pack_pw_tc_nt:
 *   w1_hi/lo <- K-major [C][>= 4C] with NT = 32,   w2_hi/lo <- K-major [4C][>= C] with NT = C.
 * Same arithmetic as ach_pw_conv_tc(ln, GELU) followed by ach_pw_conv_tc(bias, gamma, res).
 * Replaces conv_encoder.py:23-31 and sdta_encoder.py:64-73 (norm -> pwconv1 -> act -> pwconv2 -> gamma -> + input).
 * C in {32, 48, 64, 96} (ach_mlp_tc_supported); other widths use the two GEMM launches. */
typedef struct AchMlp {
    const float* x;
    const float* res;
    const float* b1;      /* [4C] */
    const float* b2;      /* [C] */
    const float* gamma;   /* [C] */
    float* out;
    long long x_bs, res_bs, out_bs;
    int B, C, P;
    float ln_eps;
} AchMlp;
ACH_API long long ach_pack_pw_tc_nt_elems(int K, int O, int NT);
ACH_API int ach_pack_pw_tc_nt(const float* wt, int K, int O, int ldw, int NT, float* w_hi, float* w_lo, void* stream);
ACH_API int ach_mlp_tc_supported(int C);
ACH_API int ach_mlp_tc(const AchMlp* p, const float* w1_hi, const float* w1_lo, const float* w2_hi, const float* w2_lo,
                       const float* wsum1, void* stream);

/* Depthwise k x k convolution (k in {3,5,7,9}, stride 1 or 2, pad k/2):
 *   out[b,c] = act(scale[c] * dw(x[b,c] + xadd[b,c]) + bias[c]) + post[c]      (post broadcast over b)
 * w is [C][k*k].  Replaces nn.Conv2d(groups=C): conv_encoder.py:10, sdta_encoder.py:23 (cascade
 * `sp = conv(sp + spx[i])` :46-50), ghost_conv.py:19-23,48-61, normal_conv.py:26-27, mobilevit.py:104,117. */
typedef struct AchDwConv {
    const float* x;
    const float* xadd;
    const float* w;
    const float* scale;
    const float* bias;
    const float* post;    /* (C, Ho*Wo) or NULL */
    float* out;
    long long x_bs, xadd_bs, out_bs;
    int B, C, H, W, Ho, Wo, k, stride, act;
} AchDwConv;
ACH_API int ach_dw_conv(const AchDwConv* p, void* stream);

/* Dense spatial convolution for the small-channel layers (patchify stem / downsample, RCNet
 * 3x3 stride-2, MobileViT 3x3): w packed [Cin][k*k][ldo] (ldo % 4 == 0, zero padded).
 *   y = act(scale * conv(x) + bias);  ln_out != 0: channels-first LayerNorm over the O outputs
 *   (biased variance, affine ln_w/ln_b) applied after bias - requires O <= 32.
 * Replaces nn.Conv2d in edgenext.py:24-34 (+LayerNorm layers.py:21-26), RadarEncoder.py:63,
 * mobilevit.py:15-21. */
typedef struct AchConvDense {
    const float* x;
    const float* w;
    const float* scale;
    const float* bias;
    const float* ln_w;
    const float* ln_b;
    float* out;
    long long x_bs, out_bs;
    int B, Cin, H, W, O, ldo, Ho, Wo, k, stride, pad, act, ln_out;
    float ln_eps;
} AchConvDense;
ACH_API int ach_conv_dense(const AchConvDense* p, void* stream);

/* Dense 3x3 convolution, stride 1, pad 1, as an implicit GEMM on tcgen05 (3xTF32, fp32-accurate):
 *   out[b, o, y, x] = act(scale[o] * sum_{c, i, j} w[o, c, i, j] * x[b, c, y-1+i, x-1+j] + bias[o]) (+ res[b, o, y, x])
 * Replaces BaseConv(ksize=3) of the CSP Bottlenecks (neck/cspdualfpn.py:42-56 incl. the "+ x" shortcut, normal_conv.py:36-49)
 * and conv_nxn_bn of the MobileViT blocks (mobilevit.py:14-21).  Weights: ach_pack_pw_tc tiles of a K-major matrix
 * [ach_conv3x3_tc_k(Cin)][ldw >= O] whose rows are ordered k = (g * 9 + tap) * 16 + c for input channel 16 g + c, tap = 3 i + j
 * (zero rows past Cin). */
typedef struct AchConv3x3Tc {
    const float* x;
    const float* scale;
    const float* bias;
    const float* res;
    float* out;
    long long x_bs, res_bs, out_bs;
    int B, Cin, H, W, O, act;
} AchConv3x3Tc;
ACH_API int ach_conv3x3_tc_k(int Cin);
ACH_API int ach_conv3x3_tc(const AchConv3x3Tc* p, const float* w_hi, const float* w_lo, void* stream);

/* Channels-first LayerNorm over C for every (b, p): layers.py:21-26 (biased variance). */
ACH_API int ach_layernorm_cf(const float* x, long long x_bs, const float* w, const float* b, float* out, long long out_bs,
                     int B, int C, int P, float eps, void* stream);

/* Channels-first LayerNorm (layers.py:21-26) fused with a 2x2 space-to-depth: out (B, 4C, H/2, W/2) with
 * out[b, c*4 + (y&1)*2 + (x&1), y/2, x/2] = LN(x)[b, c, y, x] - the im2col of the k=2, s=2 downsample conv
 * (edgenext.py:29-34), which then runs as a pointwise GEMM over K = 4C. */
ACH_API int ach_ln_s2d(const float* x, long long x_bs, const float* w, const float* b, float* out, long long out_bs, int B,
                       int C, int H, int W, float eps, void* stream);

/* Bilinear x2 upsampling, align_corners=True (nn.Upsample, ghostdualfpn.py:34). */
ACH_API int ach_upsample2x(const float* x, long long x_bs, float* out, long long out_bs, int B, int C, int H, int W,
                   void* stream);

/* SPP max-pools (kernel 5, 9, 13, stride 1, same padding; spp.py:47,52): writes the three pooled
 * copies of x (B, C, H, W) to out5/out9/out13 views. */
ACH_API int ach_spp_maxpool(const float* x, long long x_bs, float* out5, float* out9, float* out13, long long out_bs,
                    int B, int C, int H, int W, void* stream);

/* ShuffleAttention (G groups): shuffle_attention.py:48-72.  params: cweight,cbias,sweight,sbias,
 * gn_w,gn_b each (C / 2G). */
ACH_API int ach_shuffle_attention(const float* x, long long x_bs, float* out, long long out_bs, const float* cweight,
                          const float* cbias, const float* sweight, const float* sbias, const float* gn_w,
                          const float* gn_b, int B, int C, int P, int G, float eps, void* stream);

/* Plane means: out[b, c] = mean_p (x[b,c,p] + x2[b,c,p])  (x2 may be NULL).  eca.py:12,17. */
ACH_API int ach_plane_mean(const float* x, long long x_bs, const float* x2, long long x2_bs, float* out, int B, int C,
                   int P, void* stream);

/* ECA gate + concat slot + BN + ReLU of IREncoder.forward (IREncoder.py:79-89, eca.py:16-22):
 *   out[b, c, p] = relu(scale[c] * ((x + x2)[b,c,p] * sigmoid(conv1d_k(mean[b, :])[c])) + bias[c]) */
ACH_API int ach_eca_fuse(const float* x, long long x_bs, const float* x2, long long x2_bs, const float* mean,
                 const float* w1d, int k1d, const float* scale, const float* bias, float* out, long long out_bs,
                 int B, int C, int P, void* stream);

/* AvgPool2d(3, stride 1, pad 1, count_include_pad) - RadarEncoder.py:33. */
ACH_API int ach_avgpool3(const float* x, long long x_bs, float* out, long long out_bs, int B, int C, int H, int W,
                 void* stream);
/* Same pool, output CHANNEL-LAST: out[b][pixel][ceil4(C)] (pad channels 0) - the layout ach_rc_deform gathers from
 * (one 16-byte load per 4 channels of a bilinear corner instead of 4 scattered 4-byte loads). */
ACH_API int ach_avgpool3_cl(const float* x, long long x_bs, float* out, long long out_bs, int B, int C, int H, int W,
                 void* stream);

/* RCBlock body after the pool (RadarEncoder.py:65-72, dcn.py:49-63, torchvision deform_conv2d):
 *   offset/modulator 3x3 convs on `pooled`, modulated deformable 3x3 conv of `pooled`,
 *   1x1 conv + BN + ReLU, + x.  Weights packed K-major:
 *   w_om [C*9][28] (18 offset + 9 modulator outputs, 1 pad), b_om [27], w_reg [C*9][C], w1 [C][C].
 * `pooled` is channel-major planes (pooled_cl = 0, from ach_avgpool3) or channel-last [H*W][ceil4(C)]
 * (pooled_cl = 1, from ach_avgpool3_cl; the fast path).  C must be one of {3, 8, 12, 16, 24, 30, 36}. */
typedef struct AchRcDeform {
    const float* x;
    const float* pooled;
    const float* w_om;
    const float* b_om;
    const float* w_reg;
    const float* w1;
    const float* scale;
    const float* bias;
    float* out;
    long long x_bs, pooled_bs, out_bs;
    int B, C, H, W;
    int pooled_cl;
} AchRcDeform;
ACH_API int ach_rc_deform(const AchRcDeform* p, void* stream);

/* Tensor-core version of ach_rc_deform for C in {3, 8, 12, 16, 24}: both dense contractions run as implicit GEMMs on
 * tcgen05 (3xTF32), and everything LINEAR in the block is folded into them by the host (achelous_b200/engine.py:rc_tc_fold;
 * deform_conv2d -> weight_conv1 -> eval BatchNorm has no non-linearity before the ReLU).  PK = 32 if 9C <= 32 else 16 is the
 * k per tensor-core push; with 9C % PK != 0 a spare k column exists and the constants travel as weight row 9C (K = 9C + 1)
 * against a constant-1 operand column, otherwise K = 9C.  Weights pre-packed with ach_pack_pw_tc, TAP-MAJOR rows k = tap*C + ch:
 *   wom_hi/lo  <- K-major [K][28]: columns 0..17 offset conv, 18..26 = -log2(e) * modulator conv, (row 9C: b_om below)
 *   wreg_hi/lo <- K-major [K][ceil4(C)] = 2 * diag(scale) . weight_conv1 . regular_conv, O = C, (row 9C: bias below)
 * AchRcDeform fields read: x, pooled (channel-last, pooled_cl must be 1), out, the dims, and
 *   b_om  32 floats: [2t] = offset bias + (t/3 - 1), [2t+1] = offset bias + (t%3 - 1), [18+t] = -log2(e) * modulator bias
 *   bias  C floats: folded BatchNorm(weight_conv1 + its bias) bias
 * (w_om, w_reg, w1, scale are ignored).  out = x + relu(GEMM 2).  Replaces RadarEncoder.py:65-72 + dcn.py:49-63. */
ACH_API int ach_rc_deform_tc_supported(int C);
ACH_API int ach_rc_deform_tc(const AchRcDeform* p, const float* wom_hi, const float* wom_lo, const float* wreg_hi,
                             const float* wreg_lo, void* stream);

/* XCA attention core (sdta_encoder.py:162-185): from qkv (B, 3C, N) computes per (b, head) the
 * softmax(temperature * normalize(q) normalize(k)^T) (d x d) and folds it into the output projection:
 *   wt_eff[b][h*d + j][o] = sum_i proj_wt[h*d + i][o] * attn[b,h][i][j]
 * so that proj(attn @ v) == ach_pw_conv(x0 = v, wt = wt_eff (per frame)).  proj_wt is [C][ldw]. */
ACH_API int ach_xca_fold(const float* qkv, long long qkv_bs, const float* temperature, const float* proj_wt, int ldw,
                 float* wt_eff, long long wt_eff_bs, int B, int C, int heads, int N, void* stream);

/* MobileViT token self-attention (mobilevit.py:48-73,156-158): qkv (B, 3*heads*d, H*W) channel-major
 * [q | k | v], each (heads d) ordered; tokens are pixels grouped by 2x2 patch position (y&1, x&1);
 * out (B, heads*d, H*W) = softmax(q k^T / sqrt(d)) v per (group, head).  dim_head must be 8. */
ACH_API int ach_mvit_attention(const float* qkv, long long qkv_bs, float* out, long long out_bs, int B, int heads, int dim_head,
                               int H, int W, void* stream);
/* The same contract with QK^T and PV on tcgen05 (Q and P in tensor memory).  Parity-tested; not on the default plan because the
 * CUDA-core kernel is faster at d = 8.  Shapes outside its shared-memory layout are rejected, never rerouted. */
ACH_API int ach_mvit_attention_tc(const float* qkv, long long qkv_bs, float* out, long long out_bs, int B, int heads, int dim_head,
                                  int H, int W, void* stream);

/* EdgeViT blocks (backbone/vision/edgevit_modules/edgevit.py, backbone='ev'):
 * ach_subsample: out (B, C, ceil(H/sr), ceil(W/sr)) = x[:, :, ::sr, ::sr]  (the sampler nn.AvgPool2d(1, sr), :66,76).
 * ach_mhsa: qkv (B, 3*heads*d, N) channel-major [q | k | v], each (heads d) ordered -> out (B, heads*d, N) =
 *   softmax(q k^T * scale) v per (frame, head)  (:81-87).  dim_head <= 48.
 * ach_dw_convT: LocalProp, depthwise ConvTranspose2d with kernel = stride = sr (:68,91): x (B, C, h, w), w (C, sr*sr),
 *   bias (C) or NULL -> out (B, C, h*sr, w*sr). */
 /* ach_s2d: out (B, C*p*p, H/p, W/p) with out[b, (c*p + ky)*p + kx, oy, ox] = x[b, c, oy*p + ky, ox*p + kx] - the im2col of
 * PatchEmbed.proj (kernel = stride = p, edgevit.py:184), in the order of conv.weight.flatten(1). */
ACH_API int ach_s2d(const float* x, long long x_bs, float* out, long long out_bs, int B, int C, int H, int W, int patch, void* stream);
ACH_API int ach_subsample(const float* x, long long x_bs, float* out, long long out_bs, int B, int C, int H, int W, int sr, void* stream);
ACH_API int ach_mhsa(const float* qkv, long long qkv_bs, float* out, long long out_bs, int B, int heads, int dim_head, int N, float scale,
                     void* stream);
ACH_API int ach_dw_convT(const float* x, long long x_bs, const float* w, const float* bias, float* out, long long out_bs, int B, int C,
                         int h, int w_in, int sr, void* stream);

/* EfficientFormerV2 "ImageEncoder" blocks (backbone/vision/ImageEncoder.py, backbone='ef'):
 * ach_ef_attention: Attention4D / Attention4DDownsample core (:131-160, :267-289).  q (B, heads*key_dim, Nq), k (B, heads*key_dim,
 *   Nk), v (B, heads*d, Nk) channel-major views; ab (heads, Nq, Nk) = attention_biases[:, attention_bias_idxs]; th1 / th2 =
 *   talking-head 1x1 convs as [heads*heads weights (out, in) | heads biases], or NULL; add (B, heads*d, Nq) (v_local) or NULL;
 *   out (B, heads*d, Nq) = [gelu](softmax-attention(q, k, v) + add).  heads <= 8.
 * ach_upsample2x_hp: bilinear x2 with align_corners=False (nn.Upsample(scale_factor=2, mode='bilinear'), :80), optional GELU. */
ACH_API int ach_ef_attention(const float* q, long long q_bs, const float* k, long long k_bs, const float* v, long long v_bs,
                             const float* ab, const float* th1, const float* th2, const float* add, long long add_bs, float* out,
                             long long out_bs, int B, int heads, int key_dim, int d, int Nq, int Nk, float scale, int gelu, void* stream);
ACH_API int ach_upsample2x_hp(const float* x, long long x_bs, float* out, long long out_bs, int B, int C, int H, int W, int gelu,
                              void* stream);

/* Fully connected on (B, K) rows: out[b, o] = act(scale[o] * (w[o, :] . x[b, :]) + bias[o]).
 * pointnet_utils.py:16-18,38-40 (fc + BN1d + ReLU), and the global-feature half of
 * pointnet_sem_seg.py:18 (Conv1d over a point-wise constant). */
ACH_API int ach_fc(const float* x, long long x_bs, const float* w, const float* scale, const float* bias, float* out,
           long long out_bs, int B, int K, int O, int act, void* stream);

/* (B, K, N) logits -> (B, N, K) log_softmax over K (frame b at out + b * out_bs): pointnet_sem_seg.py:34-36.  K <= 32. */
ACH_API int ach_logsoftmax_t(const float* x, long long x_bs, float* out, long long out_bs, int B, int K, int N, void* stream);

/* out[b, c, p] = x[b, c, p] + post[c, p]  (post may be NULL: plain strided copy). */
ACH_API int ach_copy_add(const float* x, long long x_bs, const float* post, float* out, long long out_bs, int B, int C,
                 int P, void* stream);

/* out[b, c, p] = a[b, c, p] + b2[b, c, p]   (fpn + map sums, ghostdualfpn.py:200) */
ACH_API int ach_add(const float* a, long long a_bs, const float* b2, long long b_bs, float* out, long long out_bs, int B,
            int C, int P, void* stream);

ACH_API int ach_fill(float* x, long long n, float value, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused segmentation-decoder stages (ghostdualfpn.py:28-39 Upsample + ghost_conv.py:25-29 GhostModule,
 * called from ghostdualfpn.py:175-197).  The Ghost primary 1x1 conv (+BN scale) is linear and commutes
 * with the bilinear x2 upsampling, so the host evaluates it at LOW resolution into `v`; these kernels do
 * the full-resolution part:
 *   ach_up_ghost:       out[:, :Ci]      = x1 = relu(up2x(v) + b1)
 *                       out[:, Ci:Ci+Cn] = relu(s2 * dw3x3(x1[:, :Cn]) + b2)          (w2 [Cn][9])
 *   ach_up_ghost_head:  the same stage with Ci = Cn = 16 followed by the head GhostModule
 *                       p = relu(w3^T [x1, x2] + b3) (w3 [32][init], BN scale folded), out[:, :init] = p,
 *                       out[:, init:K] = relu(s4 * dw3x3(p[:, :K-init]) + b4), never materialising the
 *                       32-channel full-resolution tensor.  Its weight arrays are HOST pointers: they are
 *                       copied into kernel parameters (constant bank) at launch.
 *                       Instantiated for (init, K) in {(1, 2), (5, 9)}: see ach_up_ghost_head_supported.
 */
typedef struct AchUpGhost {
    const float* v;
    const float* b1;
    const float* w2;
    const float* s2;
    const float* b2;
    float* out;
    long long v_bs, out_bs;
    int B, Ci, Cn, h, w;
} AchUpGhost;
ACH_API int ach_up_ghost(const AchUpGhost* p, void* stream);

typedef struct AchUpGhostHead {
    const float* v;       /* device (B, 16, h, w) */
    float* out;           /* device (B, K, 2h, 2w) */
    const float* b1;      /* host [16]      */
    const float* w2;      /* host [16][9]   */
    const float* s2;      /* host [16]      */
    const float* b2;      /* host [16]      */
    const float* w3;      /* host [32][init]*/
    const float* b3;      /* host [init]    */
    const float* w4;      /* host [K-init][9] */
    const float* s4;      /* host [K-init]  */
    const float* b4;      /* host [K-init]  */
    long long v_bs, out_bs;
    int B, C, init, K, h, w;
} AchUpGhostHead;
ACH_API int ach_up_ghost_head_supported(int c_in, int init, int k_out);
ACH_API int ach_up_ghost_head(const AchUpGhostHead* p, void* stream);
/* Compact variant: instead of the K logit planes (p->out is ignored) writes mask (B, 2h, 2w) uint8 = first-maximum class index
 * over exactly those logits (= torch.argmax(out, 1), which is all achelous.py:283-297 keeps of them), batch stride mask_bs
 * bytes; a class whose bit in keep_mask is clear is written as 0 (achelous.py:297 keeps classes {0, 8}: keep_mask 0x101). */
ACH_API int ach_up_ghost_head_argmax(const AchUpGhostHead* p, unsigned char* mask, long long mask_bs, unsigned keep_mask,
                                     void* stream);

/* One whole decoder stage at its output resolution: the GhostModule of stage s (as ach_up_ghost, Cn = Ci) followed by
 * the NEXT stage's Upsample 1x1 conv + BN + ReLU (w1t K-major [2*Ci][32], BN scale folded; c1 bias [32]) and the next
 * stage's Ghost primary conv (w2t K-major [32][16], BN scale folded, bias applied by the consumer):
 *   out (B, 16, 2h, 2w) = w2t^T relu(w1t^T [x1, x2] + c1).   All pointers are device arrays.
 * Replaces, per stage, ach_up_ghost + two ach_pw_conv launches and their HBM round trips (ghostdualfpn.py:175-197). */
typedef struct AchUpGhostPw2 {
    const float* v;
    float* out;
    const float* b1;
    const float* w2;
    const float* s2;
    const float* b2;
    const float* w1t;
    const float* c1;
    const float* w2t;
    long long v_bs, out_bs;
    int B, Ci, C1, N2, h, w;
} AchUpGhostPw2;
ACH_API int ach_up_ghost_pw2_supported(int ci, int c1, int n2);
ACH_API int ach_up_ghost_pw2(const AchUpGhostPw2* p, void* stream);
/* Same stage with the two 1x1 convolutions on tcgen05 (3xTF32): w1_hi/lo = ach_pack_pw_tc tiles of w1t (K = 2*Ci, O = 32,
 * ldw = 32), w2_hi/lo = tiles of w2t (K = 32, O = 16, ldw = 16); the struct's w1t / w2t fields are ignored.  The depthwise
 * stage's per-channel weights travel as KERNEL PARAMETERS: dw_host is a HOST array [Ci][12] = (9 taps of w2, s2, b2, b1) per
 * channel, read at launch time (a captured CUDA graph therefore bakes them in: re-capture after changing them); the struct's
 * device arrays b1 / w2 / s2 / b2 are ignored. */
ACH_API int ach_up_ghost_pw2_tc_supported(int ci, int c1, int n2);
ACH_API int ach_up_ghost_pw2_tc(const AchUpGhostPw2* p, const float* w1_hi, const float* w1_lo, const float* w2_hi, const float* w2_lo,
                                const float* dw_host,
                                void* stream);

/* ---------------------------------------------------------------------------------------------
 * PointNet++ (pc_seg='pn2') building blocks.  The reference advertises PN2 (README.md:63) but contains no code
 * for it (nets/Achelous.py:31-32 handles only 'pn'): these implement the builder-defined network of
 * oracle/pn2.py.  xyz tensors are (B, 3, N) views, features (B, C, N); indices are int32.
 *   ach_pn2_fps:       farthest-point sampling, start index 0, ties -> lowest index; writes idx (B, npoint) and
 *                      the sampled coordinates new_xyz (B, 3, npoint).
 *   ach_pn2_group:     ball query (first nsample indices with d2 <= radius^2, padded with the first hit) + grouping:
 *                      out[b][ch][j*nsample + s] = ch < 3 ? xyz[ch][idx] - new_xyz[ch][j] : pts[ch-3][idx];
 *                      idx_out (B, S, nsample) optional.
 *   ach_pn2_group_max: (B, C, S*nsample) -> (B, C, S), max over each run of nsample.
 *   ach_pn2_interp3:   3-NN inverse-squared-distance interpolation of pts2 (B, C2, S) at xyz1 (B, 3, N1).
 */
ACH_API int ach_pn2_fps(const float* xyz, long long xyz_bs, int B, int N, int npoint, int* idx_out, float* new_xyz,
                        long long new_bs, void* stream);
ACH_API int ach_pn2_group(const float* xyz, long long xyz_bs, const float* pts, long long pts_bs, int C, const float* new_xyz,
                          long long new_bs, int B, int N, int S, int nsample, float radius, float* out, long long out_bs,
                          int* idx_out, void* stream);
ACH_API int ach_pn2_group_max(const float* x, long long x_bs, float* out, long long out_bs, int B, int C, int S, int nsample,
                              void* stream);
ACH_API int ach_pn2_interp3(const float* xyz1, long long xyz1_bs, const float* xyz2, long long xyz2_bs, const float* pts2,
                            long long pts2_bs, int B, int C2, int N1, int S, float* out, long long out_bs, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Detection post-process (second boundary: utils/utils_bbox.py).
 * ach_decode_outputs: decode_outputs (:33-85) for up to 3 levels of (B, 5+K, H_l, W_l) raw logits
 *   -> out (B, A, 5+K), xywh normalised by input_w/input_h.
 * ach_nms: non_max_suppression (:87-130) up to (not including) the host-side letterbox un-warp:
 *   per image keeps rows [x1,y1,x2,y2,obj,cls_conf,cls_idx] in score-descending order using the
 *   torchvision batched_nms coordinate trick; kept (B, A, 7), kept_idx (B, A) anchor indices,
 *   counts (B).  workspace: ach_nms_workspace_bytes(B, A).
 */
ACH_API int ach_decode_outputs(const float* const* levels, const long long* level_bs, const int* hs, const int* ws,
                       int n_levels, float* out, int B, int K, float input_h, float input_w, void* stream);
ACH_API long long ach_nms_workspace_bytes(int B, int A);
ACH_API int ach_nms(const float* decoded, int B, int A, int K, float conf_thres, float nms_thres, float* kept,
            int* kept_idx, int* counts, void* workspace, long long workspace_bytes, void* stream);
/* ach_nms_rows: the same NMS writing only the first max_keep rows (score-descending) of image b at kept + b*kept_bs (floats;
 * rows behind the survivors are zero-filled) and the TRUE survivor count at counts + b*counts_bs (ints): the detection part
 * of the compact per-frame output record (decode + NMS inside the launch plan, achelous.py:259-275). */
ACH_API int ach_nms_rows(const float* decoded, int B, int A, int K, float conf_thres, float nms_thres, float* kept,
                 long long kept_bs, int max_keep, int* counts, long long counts_bs, void* workspace,
                 long long workspace_bytes, void* stream);

/* Segmentation post-process on device: the caller-side sequence of achelous.py:283-318
 * (softmax over classes -> letterbox crop rows [y_off, y_off+nh) x cols [x_off, x_off+nw) -> cv2.resize INTER_LINEAR to
 * (OH, OW) -> argmax).  ach_seg_softmax: (B, K, P) logits -> probabilities.  ach_seg_resize_argmax: probabilities
 * (B, K, H, W) -> uint8 class map (B, OH, OW), first max on ties, OpenCV half-pixel bilinear rule in fp32. */
ACH_API int ach_seg_softmax(const float* x, long long x_bs, float* out, long long out_bs, int B, int K, int P, void* stream);
ACH_API int ach_seg_resize_argmax(const float* prob, long long prob_bs, int B, int K, int H, int W, int y_off, int x_off, int nh,
                                  int nw, unsigned char* out, int OH, int OW, void* stream);
/* ach_seg_softmax_resize_argmax: both steps in one pass over the LOGITS (no probability map in HBM), bit-identical class map;
 * out batch stride out_bs bytes; classes whose keep_mask bit is clear are written as 0 (achelous.py:297).
 * ach_seg_argmax_u8: argmax at network resolution, (B, K, P) logits -> (B, P) uint8 (P % 4 == 0).
 * ach_logsoftmax_argmax_t: (B, K, N) point logits -> (B, N) uint8 = argmax of log_softmax over K (achelous.py:262). */
ACH_API int ach_seg_softmax_resize_argmax(const float* logits, long long bs, int B, int K, int H, int W, int y_off, int x_off,
                                          int nh, int nw, unsigned char* out, long long out_bs, int OH, int OW,
                                          unsigned keep_mask, void* stream);
ACH_API int ach_seg_argmax_u8(const float* x, long long x_bs, int B, int K, int P, unsigned keep_mask, unsigned char* out,
                              long long out_bs, void* stream);
ACH_API int ach_logsoftmax_argmax_t(const float* x, long long x_bs, unsigned char* out, long long out_bs, int B, int K, int N,
                                    void* stream);

/* Input pre-processing on device (achelous.py:200-246, utils/utils.py:20-54).
 * Image: Pillow's two-pass 8-bit BICUBIC resize (Resample.c; coefficient tables `bounds` [out][2] = (first source
 * index, tap count) and `kk` [out][ksize] 22-bit fixed point, built on the host exactly as precompute_coeffs +
 * normalize_coeffs_8bpc do) - ach_pre_resize_h: src (B, ih, iw, 3) uint8, rows [r0, r0+rows) -> tmp (B, rows, nw, 3);
 * ach_pre_resize_v_norm: vertical pass over tmp (its `bounds` are relative to r0; identity != 0: no vertical pass) fused
 * with the letterbox paste at (x_off, y_off) on a 128-grey canvas, HWC->CHW and preprocess_input -> out (B, 3, H, W) fp32.
 * ach_pre_radar: per-sample (x - min) / (max - min) + 1e-13 (utils.py:51-54), src fp32 (is_f64 = 0) or fp64 -> fp32.
 * ach_pre_points: feat (n_rows, C) fp64 row-major, idx (B, N) -> out (B, C, N) fp32 = feat[idx] / column L2 norm over the
 * N sampled rows (sklearn normalize(axis=0); zero norm -> 1), achelous.py:224-246. */
ACH_API int ach_pre_resize_h(const unsigned char* src, long long src_bs, int B, int iw, int r0, int rows, int nw, const int* bounds,
                             const int* kk, int ksize, unsigned char* tmp, long long tmp_bs, void* stream);
ACH_API int ach_pre_resize_v_norm(const unsigned char* tmp, long long tmp_bs, int B, int nw, int nh, const int* bounds, const int* kk,
                                  int ksize, int identity, float* out, long long out_bs, int H, int W, int x_off, int y_off,
                                  void* stream);
ACH_API int ach_pre_radar(const void* src, long long src_bs, int is_f64, int B, long long n, float* out, long long out_bs, void* stream);
ACH_API int ach_pre_points(const double* feat, int n_rows, int C, const int* idx, int B, int N, float* out, void* stream);

/* Output collection over NVLink peer memory (replaces the gather half of nn.DataParallel's scatter / gather,
 * /root/reference/achelous.py:176-177, for one-process-per-GPU serving).  Every rank owns one allocation from ach_peer_alloc
 * (plain cudaMalloc, zero-filled), exports it (ach_peer_export -> ach_peer_handle_bytes() opaque bytes, shipped to the other
 * processes by the host), and maps the peers' allocations with ach_peer_open.  ach_peer_copy is a stream-ordered copy-engine
 * transfer between such pointers (no SM work).  ach_peer_signal: after everything earlier on `stream`, store `value` (release,
 * system scope) to flags[i] for i < n (n <= 64; `flags` is a DEVICE array of pointers, usually into peers' allocations; NULL
 * entries are skipped).  ach_peer_wait: the stream does not advance until local words flags[0..n) have all reached `value`
 * (wrap-safe signed comparison; acquire, system scope; one spinning warp; a word that has not arrived after 60 s traps
 * the kernel - a sticky CUDA error instead of a GPU that spins forever on a dead peer). */
ACH_API int ach_peer_alloc(long long bytes, void** ptr);
ACH_API int ach_peer_free(void* ptr);
ACH_API int ach_peer_handle_bytes(void);
ACH_API int ach_peer_export(void* ptr, unsigned char* handle);
ACH_API int ach_peer_open(const unsigned char* handle, void** ptr);
ACH_API int ach_peer_close(void* ptr);
ACH_API int ach_peer_copy(void* dst, const void* src, long long bytes, void* stream);
ACH_API int ach_peer_signal(unsigned* const* flags, int n, unsigned value, void* stream);
ACH_API int ach_peer_wait(const unsigned* flags, int n, unsigned value, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACHELOUS_B200_H */
