/*
 * slak_b200.h -- C ABI of libslak_b200.so: the B200 (sm_100a) implementation of
 * SLaK's large-kernel depthwise-convolution hot path.
 *
 * Every entry point is a plain-C function over raw device pointers and sizes
 * (no torch types).  All tensors are dense NCHW, contiguous.  Every call
 * enqueues work on `stream` (a cudaStream_t passed as void*; NULL = legacy
 * default stream) and returns immediately; the return value is 0 on success or
 * a negative slak_status code, and slak_last_error() gives a message.  Nothing
 * calls exit(): the reference wrappers abort the process on a CUDA/CUTLASS
 * error (forward_fp32.cu:173-192), these return a code instead.
 *
 * Reference interfaces replaced (paths relative to the reference checkout):
 *   cutlass/examples/19_large_depthwise_conv2d_torch_extension/frontend.h:3-10
 *     forward_fp32/fp16            -> slak_dwconv2d_fwd
 *     backward_data_fp32/fp16      -> slak_dwconv2d_bwd_data
 *     backward_filter_fp32/fp16    -> slak_dwconv2d_bwd_filter
 *   models/SLaK.py:89-100 (ReparamLargeKernelConv.forward, conv + BN + sum)
 *                                  -> slak_lk_branches_* / slak_bn_* (fused)
 *   sparse_core.py:316-333 (Masking.apply_mask) -> slak_mask_apply
 *   funcs.py:107-114 (magnitude_prune)          -> slak_mask_prune_magnitude
 */
#ifndef SLAK_B200_H_
#define SLAK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SLAK_API
#else
#define SLAK_API __attribute__((visibility("default")))
#endif

/* Element types of activation / gradient tensors. */
enum slak_dtype { SLAK_F32 = 0, SLAK_F16 = 1, SLAK_BF16 = 2 };

/* Status codes (return values are 0 or negative). */
enum slak_status {
  SLAK_OK = 0,
  SLAK_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, even kernel ... */
  SLAK_ERR_UNSUPPORTED = -2,  /* shape outside what the kernels cover */
  SLAK_ERR_CUDA = -3,         /* a CUDA runtime call failed (see slak_last_error) */
  SLAK_ERR_WORKSPACE = -4     /* workspace too small */
};

/* Library version: major*10000 + minor*100 + patch. */
SLAK_API int slak_version(void);
/* Thread-local message for the last failing call on this thread. */
SLAK_API const char* slak_last_error(void);
/* 1 if a CUDA device of compute capability 10.x is visible, else 0. */
SLAK_API int slak_device_ok(void);

/* ---------------------------------------------------------------------------
 * Depthwise conv2d, stride 1, dilation 1, "same" padding (kh/2, kw/2),
 * cross-correlation:
 *   y[n,c,p,q] = sum_{r,s} x[n,c,p+r-kh/2,q+s-kw/2] * w[c,0,r,s]
 * (forward_fp32.cu:135-144,227).  kh and kw must be odd.
 *   x, y : [N,C,H,W] of `dtype`
 *   w    : [C,1,kh,kw] of `wdtype` (SLAK_F32, or the same as `dtype`).  When
 *          dtype is a 16-bit type the weights are rounded to it before use
 *          (what autocast's cast_inputs does to the reference's FP16 path,
 *          depthwise_conv2d_implicit_gemm.py:35) and products accumulate in fp32.
 * Route: bf16 activations with fp32 taps and one kernel side equal to 5 (K x 5, 5 x K, 5 x 5 -- every depthwise
 * shape of a Decom SLaK Block) on planes up to 62 x 62 run on the tcgen05 banded-Toeplitz kernels
 * (slak_dwconv2d_uses_tc() == 1); everything else -- fp32 (exact, the reference's allclose tolerance), fp16, square
 * kernels, larger planes -- runs the CUDA-core kernels. */
SLAK_API int slak_dwconv2d_uses_tc(int N, int C, int H, int W, int kh, int kw, int dtype, int wdtype);
SLAK_API int slak_dwconv2d_fwd(const void* x, const void* w, void* y,
                               int N, int C, int H, int W, int kh, int kw,
                               int dtype, int wdtype, void* stream);

/* dx[n,c,h,w] = sum_{r,s} dy[n,c,h-r+kh/2,w-s+kw/2] * w[c,0,r,s]
 * (backward_data_fp32.cu:199-263). */
SLAK_API int slak_dwconv2d_bwd_data(const void* dy, const void* w, void* dx,
                                    int N, int C, int H, int W, int kh, int kw,
                                    int dtype, int wdtype, void* stream);

/* dw[c,0,r,s] = sum_{n,p,q} dy[n,c,p,q] * x[n,c,p+r-kh/2,q+s-kw/2], fp32 output
 * whatever the input dtype (backward_filter_fp16.cu:18,187).  Deterministic:
 * per-CTA partial sums go to `workspace` and are reduced in a fixed order (the
 * reference uses fp32 atomicAdd, dwconv2d_direct_epilogue_simt.h:160-185).
 * workspace must hold slak_dwconv2d_bwd_filter_workspace(...) bytes. */
SLAK_API size_t slak_dwconv2d_bwd_filter_workspace(int N, int C, int H, int W,
                                                   int kh, int kw, int dtype);
SLAK_API int slak_dwconv2d_bwd_filter(const void* dy, const void* x, float* dw,
                                      int N, int C, int H, int W, int kh, int kw,
                                      int dtype, void* workspace,
                                      size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * The three depthwise branches of ReparamLargeKernelConv.forward (models/SLaK.py:89-100,
 * Decom=True) in one call, x read once:
 *   y1 = dwconv_{KL x KS}(x, w1)   y2 = dwconv_{KS x KL}(x, w2)   y3 = dwconv_{KS x KS}(x, w3)
 * w1 [C,1,KL,KS], w2 [C,1,KS,KL], w3 [C,1,KS,KS] are the fp32 Parameters (rounded to `dtype`
 * inside, as in slak_dwconv2d_fwd).  For bf16 tensors with KS == 5, 8 <= H,W <= 62 and
 * W % 8 == 0 this is ONE tcgen05 (tensor-core, banded-Toeplitz GEMM) kernel; every other
 * case runs the three CUDA-core kernels back to back.  w3/y3 may be NULL (no small branch,
 * models/SLaK.py:85).  slak_lk_branches_uses_tc() reports which path a shape takes.
 * ------------------------------------------------------------------------- */
SLAK_API int slak_lk_branches_uses_tc(int N, int C, int H, int W, int KL, int KS, int dtype);
SLAK_API int slak_lk_branches_fwd(const void* x, const float* w1, const float* w2, const float* w3,
                                  void* y1, void* y2, void* y3, int N, int C, int H, int W,
                                  int KL, int KS, int dtype, void* stream);

/* Backward of slak_lk_branches_fwd on the tensor cores; only for shapes where
 * slak_lk_branches_uses_tc() is 1 (otherwise SLAK_ERR_UNSUPPORTED: call the per-branch
 * slak_dwconv2d_bwd_* functions).
 *   bwd_data  : dx = dgrad(dy1,w1) + dgrad(dy2,w2) + dgrad(dy3,w3); `tmp` is a caller-provided
 *               scratch tensor of the size of dx (holds the 5x5 branch between the two launches).
 *   bwd_filter: dw1 [C,1,KL,KS], dw2 [C,1,KS,KL], dw3 [C,1,KS,KS] in fp32, deterministic;
 *               workspace of slak_lk_branches_bwd_filter_workspace() bytes. */
SLAK_API int slak_lk_branches_bwd_uses_tc(int N, int C, int H, int W, int KL, int KS, int dtype);
SLAK_API int slak_lk_branches_bwd_data(const void* dy1, const void* dy2, const void* dy3,
                                       const float* w1, const float* w2, const float* w3,
                                       void* dx, void* tmp, int N, int C, int H, int W,
                                       int KL, int KS, int dtype, void* stream);
/* the same with an fp32 result: dx = addend (fp32 [N,C,H,W], may be NULL) + the three bf16 dgrads; used by the
 * Block backward where `addend` is the shortcut gradient */
SLAK_API int slak_lk_branches_bwd_data_f32(const void* dy1, const void* dy2, const void* dy3, const float* w1,
                                           const float* w2, const float* w3, const float* addend, float* dx,
                                           void* tmp, int N, int C, int H, int W, int KL, int KS, void* stream);
/* Inference form of the Decom layer after re-parameterisation (slak_b200.slak.ReparamLargeKernelConv.merge_kernel: the
 * three BatchNorms folded as fuse_bn does, models/SLaK.py:49-58, the 5 x 5 kernel added into the centre of the 5 x KL
 * one; the reference's merge_kernel :111-122 covers only the non-Decom layout):
 *   y = dwconv_{KL x 5}(x, wv) + dwconv_{5 x KL}(x, wh) + bias[c]      x read once, y written once
 * wv [C,1,KL,5], wh [C,1,5,KL], bias [C] fp32 (may be NULL); tensor-core shapes only (slak_lk_branches_uses_tc). */
SLAK_API int slak_lk_merged_fwd(const void* x, const float* wv, const float* wh, const float* bias, void* y, int N, int C,
                                int H, int W, int KL, int dtype, void* stream);
SLAK_API size_t slak_lk_branches_bwd_filter_workspace(int N, int C, int H, int W, int KL, int KS);
SLAK_API int slak_lk_branches_bwd_filter(const void* x, const void* dy1, const void* dy2, const void* dy3,
                                         float* dw1, float* dw2, float* dw3, int N, int C, int H, int W,
                                         int KL, int KS, int dtype, void* workspace,
                                         size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * Fused glue of a SLaK Block around the branches (models/SLaK.py:89-100 BN + sum, :153-166
 * permute / LayerNorm / gamma / residual); bf16 activations, fp32 statistics and parameters.
 * Shapes: y_i, du, dy_i [N,C,H,W] bf16 (HW = H*W); xn, h2, dxn, dh2 [N,H,W,C] bf16; x, out, dout
 * [N,C,H,W] fp32; per-channel vectors fp32.
 *   block_conv_fwd    : slak_lk_branches_fwd + per-channel (sum, sumsq) of y1,y2,y3 -> sums [C][6] (double)
 *   bn3_finalize_fwd  : nn.BatchNorm2d training semantics from GLOBAL sums/count (all-reduce sums first for
 *                       SyncBN): scale[3][C] = w*istd, shift[C] = sum_i (b_i - mean_i*scale_i), mean/istd[3][C],
 *                       running stats updated in place (pass NULL to skip)
 *   bn3_eval_affine   : the same affine from running statistics (eval mode)
 *   bn3_sum_ln_fwd    : xn = LayerNorm_C(sum_i scale_i*y_i + shift), per-pixel mu/rstd saved
 *   block_residual_fwd: out = x + dp[n]*gamma[c]*h2 (dp may be NULL = 1; out_bf16 optional copy)
 *   block_residual_bwd: dh2 = dout*gamma*dp ; dgamma_part [parts][2][C] partial sums of dout*h2*dp and of dh2
 *   gelu_bwd_bias     : dh = da * gelu'(h) (exact erf GELU) over [rows][K] bf16, part [parts][K] column sums of dh
 *   bn3_sum_ln_bwd    : du = LayerNorm backward of dxn ; part [parts][6][C] = dlnw, dlnb, sum du, sum du*y_i
 *   bn3_finalize_bwd  : from GLOBAL S[4][C] -> coef[9][C] (dy_i = A_i*du + B_i*y_i + C_i), dbnw/dbnb [3][C]
 *   bn3_bwd_apply     : dy1, dy2, dy3 in one pass
 * ------------------------------------------------------------------------- */
SLAK_API size_t slak_block_conv_fwd_workspace(int N, int C, int H, int W);
SLAK_API int slak_block_conv_fwd(const void* x, const float* w1, const float* w2, const float* w3, void* y1,
                                 void* y2, void* y3, double* sums, void* workspace, size_t workspace_bytes,
                                 int N, int C, int H, int W, int KL, void* stream);
/* bnw/bnb/rmean/rvar: HOST arrays of three device pointers (branch K x 5, 5 x K, 5 x 5), each [C];
 * entries of rmean/rvar may be NULL to skip the running-statistics update */
/* count_dev (may be NULL): DEVICE pointer to the global element count per channel; when given it overrides `count`
 * (SyncBN with unequal per-rank batches: the count is all-reduced together with the sums, no host round trip) */
SLAK_API int slak_bn3_finalize_fwd(const double* sums, double count, const double* count_dev, const float* const* bnw,
                                   const float* const* bnb, float* const* rmean, float* const* rvar, float eps,
                                   float momentum, int C, float* scale, float* shift, float* mean, float* istd,
                                   void* stream);
/* SyncBatchNorm (models/SLaK.py:24-28) with the statistics exchange FUSED into the finalize kernels: a one-shot
 * all-reduce over NVLink peer memory instead of one NCCL collective per BatchNorm (torch's SyncBatchNorm issues 3
 * all_gathers per Block forward and 3 all_reduces per backward, torch/nn/modules/_functions.py:49,74,158).
 * peer_bases: HOST array of `world` device pointers, entry r = rank r's symmetric buffer mapped into this process
 * (torch.distributed._symmetric_memory supplies allocation and mapping); every rank's payload for this call site sits at
 * slot_off (fwd: double sums[C][6], count, [global count out]; bwd: float S[4][C]), flag_off addresses `world` uint32
 * flags of the site, epoch_dev is the site's device-side call counter (bumped by the call).  Ranks signal, wait for
 * all peers and add the payloads in rank order (bitwise identical results on every rank).  world <= 8. */
SLAK_API int slak_bn3_finalize_fwd_sync(const void* const* peer_bases, size_t slot_off, size_t flag_off, int rank, int world,
                                        uint32_t* epoch_dev, const float* const* bnw, const float* const* bnb,
                                        float* const* rmean, float* const* rvar, float eps, float momentum, int C,
                                        float* scale, float* shift, float* mean, float* istd, void* stream);
SLAK_API int slak_bn3_finalize_bwd_sync(const void* const* peer_bases, size_t slot_off, size_t flag_off, int rank, int world,
                                        uint32_t* epoch_dev, const double* count_dev, const float* const* bnw,
                                        const float* mean, const float* istd, int C, float* coef, float* dbnw, float* dbnb,
                                        void* stream);
SLAK_API int slak_bn3_eval_affine(const float* const* bnw, const float* const* bnb, const float* const* rmean,
                                  const float* const* rvar, float eps, int C, float* scale, float* shift,
                                  void* stream);
SLAK_API int slak_bn3_sum_ln_fwd(const void* y1, const void* y2, const void* y3, const float* scale,
                                 const float* shift, const float* lnw, const float* lnb, float eps, void* xn,
                                 float* mu, float* rstd, int N, int C, int HW, void* stream);
SLAK_API int slak_block_residual_fwd(const float* x, const void* h2, const float* gamma, const float* dp,
                                     float* out, void* out_bf16, int N, int C, int HW, void* stream);
SLAK_API int slak_block_residual_bwd_parts(int N, int C, int HW);
SLAK_API int slak_block_residual_bwd(const float* dout, const void* h2, const float* gamma, const float* dp,
                                     void* dh2, float* dgamma_part, int N, int C, int HW, void* stream);
SLAK_API int slak_gelu_bwd_bias_parts(int64_t rows, int K);
SLAK_API int slak_gelu_bwd_bias(const void* da, const void* h, void* dh, float* part, int64_t rows, int K,
                                void* stream);
SLAK_API int slak_bn3_sum_ln_bwd_parts(int N, int C, int HW);
SLAK_API int slak_bn3_sum_ln_bwd(const void* dxn, const void* y1, const void* y2, const void* y3,
                                 const float* scale, const float* shift, const float* lnw, const float* mu,
                                 const float* rstd, void* du, float* part, int N, int C, int HW, void* stream);
/* S = GLOBAL sums (dy coefficients); S_local (NULL = S) = this rank's sums, from which dbnw/dbnb are taken: like
 * torch's SyncBatchNorm the parameter gradients are per-rank and the data-parallel wrapper averages them */
SLAK_API int slak_bn3_finalize_bwd(const float* S, const float* S_local, double count, const double* count_dev,
                                   const float* const* bnw, const float* mean,
                                   const float* istd, int C, float* coef, float* dbnw, float* dbnb, void* stream);
SLAK_API int slak_bn3_bwd_apply(const void* du, const void* y1, const void* y2, const void* y3, const float* coef,
                                void* dy1, void* dy2, void* dy3, int N, int C, int HW, void* stream);

/* ---------------------------------------------------------------------------
 * Downsampling layer between two stages (models/SLaK.py:194-199: LayerNorm(channels_first) -> Conv2d(k=2, s=2)) as
 * LayerNorm + GEMM (csrc/block_glue2.cu): the LayerNorm writes its bf16 output as the A operand of the convolution-as-GEMM,
 * A[(n, h/2, w/2)][((h&1)*2 + (w&1))*C + c]; slak_mlp_gemm_nt / slak_mlp_gemm_tn_splitk do the convolution, its data and
 * weight gradients; the two layout kernels move between the token-major bf16 side and the fp32 NCHW residual stream.
 * x, dx: fp32 NCHW [N, C, H, W]; C % 8 == 0, C <= 768, H and W even; part: [parts][2][C] (dlnw, dlnb partials).
 * nchw_to_nhwc also returns the column sums of its bf16 output in part[:, 1, :] ([parts][2][C]): the bias gradient. */
SLAK_API int slak_ln2d_patch_fwd(const float* x, const float* lnw, const float* lnb, float eps, void* A, float* mean,
                                 float* rstd, int N, int C, int H, int W, void* stream);
SLAK_API int slak_ln2d_patch_bwd_parts(int N, int C, int H, int W);
SLAK_API int slak_ln2d_patch_bwd(const void* dA, const float* x, const float* lnw, const float* mean, const float* rstd,
                                 float* dx, float* part, int N, int C, int H, int W, void* stream);
SLAK_API int slak_nhwc_to_nchw(const void* src_bf16, float* dst, void* dst_bf16 /* or NULL */, int N, int C, int HW,
                               void* stream);
SLAK_API int slak_nchw_to_nhwc_parts(int N, int C, int HW);
SLAK_API int slak_nchw_to_nhwc(const float* src, void* dst_bf16, float* part, int N, int C, int HW, void* stream);

/* ---------------------------------------------------------------------------
 * Stem (models/SLaK.py:189-193: Conv2d(Cin, C, k=4, s=4) -> LayerNorm(channels_first)) as patch rows + GEMM + LayerNorm over
 * token rows (csrc/block_glue2.cu).  slak_patchify4: A[(n,ho,wo)][ci*16 + kh*4 + kw] (bf16, K = 64, zero beyond 16*Cin) from
 * the fp32 NCHW image, Cin <= 4, H % 4 == W % 4 == 0.  slak_ln_rows_fwd: Y bf16 [N*HW][C] (GEMM output) -> LayerNorm over C
 * -> fp32 NCHW (+ bf16 copy or NULL), mean / rstd per token.  slak_ln_rows_bwd: NCHW fp32 gradient -> dY bf16 [N*HW][C];
 * part [parts][3][C] = dlnw, dlnb and the column sums of dY (the convolution's bias gradient). */
SLAK_API int slak_patchify4(const float* x, void* A, int N, int Cin, int H, int W, void* stream);
SLAK_API int slak_ln_rows_fwd(const void* Y, const float* lnw, const float* lnb, float eps, float* out, void* out_bf16,
                              float* mean, float* rstd, int N, int C, int HW, void* stream);
SLAK_API int slak_ln_rows_bwd_parts(int N, int C, int HW);
SLAK_API int slak_ln_rows_bwd(const float* dout, const void* Y, const float* lnw, const float* mean, const float* rstd,
                              void* dY, float* part, int N, int C, int HW, void* stream);

/* ---------------------------------------------------------------------------
 * Pointwise MLP of a Block on the tensor cores (models/SLaK.py:157-160, pwconv1 -> GELU -> pwconv2, and its backward):
 * tcgen05 GEMMs with the elementwise passes folded into the epilogues (csrc/mlp_tc.cu).  All matrices bf16 row-major,
 * 16-byte aligned, N % 8 == 0, K % 8 == 0, N <= 3072.
 *   slak_mlp_gemm_nt(epi, ...): D[M,N] = a[M,K] b[N,K]^T, fp32 accumulate, then
 *     epi 0 FC1   : out0 (may be NULL) = H = D + bias (bf16, nn.Linear under autocast), out1 = gelu(H) (exact erf GELU
 *                   of the rounded H, nn.GELU on the bf16 tensor)
 *     epi 1 BIAS  : out0 = D + bias
 *     epi 2 DGELU : out0 = dH = D * gelu'(aux_h); colpart[slak_mlp_parts(M,N)][N] fp32 = per-CTA partial column sums of
 *                   dH (fold with slak_colsum_f32: the bias gradient of pwconv1)
 *     epi 3 PLAIN : out0 = D
 *   slak_mlp_gemm_tn_splitk: part[slak_mlp_wgrad_splits(M,Ma,Nb)][Ma][Nb] fp32 = per-split partial sums of
 *     p[M,Ma]^T q[M,Nb] (contraction over the M tokens: the weight gradients dW = dY^T X that autograd forms for
 *     nn.Linear); fold with slak_colsum_f32(part, splits, Ma*Nb, dW): fixed order, deterministic.
 *   slak_mlp_fc1_gelu_fwd / slak_mlp_fc2_dgelu_bwd: the epi 0 / epi 2 calls under their round-1 names.
 * ------------------------------------------------------------------------- */
/* fp32 weight [R][Cc] (an nn.Linear parameter) -> bf16 copy wb [R][Cc] and bf16 transpose wt [Cc][R] in one pass: the
 * autocast cast of the weight plus the K-major operand of its data-gradient GEMM */
SLAK_API int slak_cast_transpose_bf16(const float* w, void* wb, void* wt, int R, int Cc, void* stream);
SLAK_API int slak_mlp_parts(int M, int N);
SLAK_API int slak_mlp_gemm_nt(int epi, const void* a, const void* b, const float* bias, const void* aux_h, void* out0,
                              void* out1, float* colpart, int M, int N, int K, void* stream);
SLAK_API int slak_mlp_wgrad_splits(int M, int Ma, int Nb);
SLAK_API int slak_mlp_gemm_tn_splitk(const void* p, const void* q, float* part, int M, int Ma, int Nb, void* stream);
SLAK_API int slak_mlp_fc1_gelu_fwd(const void* x, const void* w, const float* bias, void* h, void* a, int M, int N,
                                   int K, void* stream);
SLAK_API int slak_mlp_fc2_dgelu_bwd(const void* g, const void* wt, const void* h, void* dh, float* colpart, int M,
                                    int N, int K, void* stream);

/* out[c] = sum over r of part[r*cols + c], rows added in a fixed order: folds the per-CTA partial rows the
 * *_parts kernels above emit (bias / gamma / LayerNorm-parameter gradients: the column sums the reference gets
 * from autograd's `grad.sum(0)` of nn.Linear / the `gamma * x` broadcast, models/SLaK.py:158-164). */
SLAK_API int slak_colsum_f32(const float* part, int rows, int cols, float* out, void* stream);

/* ---------------------------------------------------------------------------
 * LayerNorm over the channels of an NCHW tensor: the "channels_first" branch of
 * LayerNorm.forward (models/SLaK.py:256-261), used by the stem and the three
 * downsampling layers (models/SLaK.py:192-203) on either side of the Blocks.
 *     y = w[c] * (x - mean_c x) / sqrt(mean_c (x - mean_c x)^2 + eps) + b[c]
 * x, y and the incoming gradient g may each be SLAK_F32 or SLAK_BF16; mean and
 * rstd ([N*HW] fp32, may both be NULL in the forward when no backward follows)
 * are what the backward needs besides x.  The backward writes dx (dtype of x),
 * dw and db ([C] fp32, overwritten); `part` is a workspace of
 * slak_layernorm2d_bwd_parts(N, HW) * 2 * C floats.
 * ------------------------------------------------------------------------- */
SLAK_API int slak_layernorm2d_fwd(const void* x, int x_dtype, const float* w, const float* b, float eps,
                                  void* y, int y_dtype, float* mean, float* rstd, int N, int C, int HW,
                                  void* stream);
SLAK_API int slak_layernorm2d_bwd_parts(int N, int HW);
SLAK_API int slak_layernorm2d_bwd(const void* g, int g_dtype, const void* x, int x_dtype, const float* w,
                                  const float* mean, const float* rstd, void* dx, float* part, float* dw,
                                  float* db, int N, int C, int HW, void* stream);

/* ---------------------------------------------------------------------------
 * Sparse-mask engine (sparse_core.py:316-333, funcs.py:107-114).
 * ------------------------------------------------------------------------- */

/* One launch over a list of tensors: w_i[j] = w_i[j] * mask_i[j] for every i
 * (IEEE multiply, so a pruned negative weight becomes -0.0 exactly as
 * `tensor.data = tensor.data*self.masks[name]` does).  `w_ptrs`, `mask_ptrs`
 * and `numels` are DEVICE arrays of length `count`; `extra_ptrs` (may be NULL)
 * is a second list of tensors multiplied by the same masks (SGD momentum
 * buffers, sparse_core.py:325-326; entries may be NULL). */
SLAK_API int slak_mask_apply(float* const* w_ptrs, const float* const* mask_ptrs,
                             float* const* extra_ptrs, const int64_t* numels,
                             int count, int64_t max_numel, void* stream);

/* Fused multi-tensor AdamW + mask apply + mask-aware EMA, one launch (optim_factory.py:149-150 torch.optim.AdamW;
 * sparse_core.py:322-333 Masking.apply_mask; model_sema.py:67-91 ModelEma.update).  All tables are DEVICE arrays
 * indexed by tensor: p, g, m (exp_avg), v (exp_avg_sq) fp32 tensors of numel[t] elements; mask / ema tables and their
 * entries may be NULL; lr / wd per tensor (double).  Work is a flat list of nchunks chunks (chunk_tensor[c],
 * chunk_off[c]) of at most chunk_elems elements.  step_dev: device counter of completed steps (this call computes
 * step + 1 and then increments it: CUDA-graph replayable).  do_adam = 0 applies only the mask / EMA part (p is the
 * weight to average).  Per-element arithmetic: fp32 in the operation order of torch's single-tensor AdamW
 * (csrc/optim.cu). */
SLAK_API int slak_adamw_mask_ema_step(float* const* p, const float* const* g, float* const* m, float* const* v,
                                      const float* const* mask, float* const* ema, const int64_t* numel,
                                      const double* lr, const double* wd, const int32_t* chunk_tensor,
                                      const int64_t* chunk_off, int nchunks, int chunk_elems, double beta1,
                                      double beta2, double eps, double ema_decay, int64_t* step_dev, int do_adam,
                                      void* stream);

/* Magnitude prune of one layer: zero the mask at the k smallest |w| positions
 * (ties broken by lower flat index first, like a stable ascending sort).
 * mask is updated in place; `workspace` needs
 * slak_mask_prune_workspace(numel) bytes. */
SLAK_API size_t slak_mask_prune_workspace(int64_t numel);
SLAK_API int slak_mask_prune_magnitude(const float* w, float* mask, int64_t numel,
                                       int64_t k, void* workspace,
                                       size_t workspace_bytes, void* stream);

/* Growth by score (gradient_growth / momentum_growth, funcs.py:196-299): mask[i] = 1 at the k positions of largest
 * |score[i]| among the positions whose mask is 0 (the reference sorts |score * (mask == 0)| descending and takes idx[:k];
 * ties go to the lower flat index here).  Same workspace as slak_mask_prune_magnitude. */
SLAK_API int slak_mask_grow_topk(const float* score, float* mask, int64_t numel, int64_t k, void* workspace,
                                 size_t workspace_bytes, void* stream);
/* out[0] (device) = |x| of the k-th largest magnitude: SNIP's global threshold torch.topk(all_scores, keep)[-1]
 * (sparse_core.py:36-38) without sorting 30 M scores. */
SLAK_API int slak_select_kth_largest_abs(const float* x, int64_t numel, int64_t k, void* workspace, size_t workspace_bytes,
                                         float* out, void* stream);
/* 0/1 fp32 mask <-> bit mask, element i = bit (i % 32) of word i / 32; words holds (numel + 31) / 32 uint32.  For
 * packed-mask checkpoints (the reference saves no masks and rebuilds them as weight != 0, sparse_core.py:158-172) and
 * the 32x smaller mask broadcast. */
SLAK_API int slak_mask_pack_bits(const float* mask, uint32_t* words, int64_t numel, void* stream);
SLAK_API int slak_mask_unpack_bits(const uint32_t* words, float* mask, int64_t numel, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SLAK_B200_H_ */
