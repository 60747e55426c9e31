// extern "C" entry points of libslak_b200.so (see include/slak_b200.h).
#include "common.cuh"
#include "block_glue2.cuh"
#include <stdarg.h>
#include <string.h>

namespace slak {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// dwconv_simt.cu
int dwconv_simt_fwd(const void* x, const void* w, void* y, int N, int C, int H, int W, int kh, int kw,
                    int dtype, int wdtype, int flip, cudaStream_t st);
size_t dwconv_simt_wgrad_workspace(int N, int C, int H, int W, int kh, int kw);
int dwconv_simt_wgrad(const void* dy, const void* x, float* dw, int N, int C, int H, int W, int kh,
                      int kw, int dtype, void* workspace, cudaStream_t st);
// dwconv_tc_fwd.cu
namespace tc {
bool lk3_tc_supported(int N, int C, int H, int W, int KL);
bool lk3_bwd_tc_supported(int N, int C, int H, int W, int KL);
int lk3_fwd_tc(const void* x, const float* w1, const float* w2, const float* w3, void* y1, void* y2, void* y3,
               int N, int C, int H, int W, int KL, float* stats, cudaStream_t st);
int lk3_fwd_tc_splits(int N, int C, int H, int W);
int lk_conv_tc(const void* in_t, const float* wt, const void* in_n, const float* wn, const void* addend, void* out,
               const float* addend_f32, float* out_f32, int N, int C, int H, int W, int KL, int KN, int flip,
               cudaStream_t st, const float* bias = nullptr);
int mlp_parts(int M, int N);
int mlp_gemm_nt(int epi, const void* a, const void* b, const float* bias, const void* aux_h, void* out0, void* out1,
                float* colpart, int M, int N, int K, cudaStream_t st);
int mlp_wgrad_splits(int M, int Ma, int Nb);
int mlp_gemm_tn_splitk(const void* p, const void* q, float* part, int M, int Ma, int Nb, cudaStream_t st);
namespace dense {
bool supported(int N, int C, int H, int W, int KL);
int dgrad(const void* dy1, const void* dy2, const void* dy3, const float* w1, const float* w2, const float* w3, const float* addend,
          float* dx, int N, int C, int H, int W, int KL, cudaStream_t st);
}
size_t lk3_wgrad_tc_workspace(int N, int C, int H, int W, int KL);
int lk3_wgrad_tc(const void* x, const void* dy1, const void* dy2, const void* dy3, float* dw1, float* dw2,
                 float* dw3, int N, int C, int H, int W, int KL, void* workspace, cudaStream_t st);
}
// block_fused.cu
namespace blk {
int colsum(const float* part, int rows, int cols, float* out, cudaStream_t st);
int cast_transpose(const float* w, void* wb, void* wt, int R, int Cc, cudaStream_t st);
int bn3_stats_finalize(const float* part, int splits, double* sums_ws, int C, cudaStream_t st);
int bn3_finalize_fwd(const double* sums, double count, const double* count_dev, const float* const* bnw, const float* const* bnb,
                     float* const* rmean, float* const* rvar, float eps, float momentum, int C, float* scale, float* shift,
                     float* mean, float* istd, cudaStream_t st);
int bn3_eval_affine(const float* const* bnw, const float* const* bnb, const float* const* rmean, const float* const* rvar,
                    float eps, int C, float* scale, float* shift, cudaStream_t st);
int bn3_finalize_fwd_sync(const void* const* peers, size_t slot_off, size_t flag_off, int rank, int world, uint32_t* epoch_dev,
                          const float* const* bnw, const float* const* bnb, float* const* rmean, float* const* rvar, float eps,
                          float momentum, int C, float* scale, float* shift, float* mean, float* istd, cudaStream_t st);
int bn3_finalize_bwd_sync(const void* const* peers, size_t slot_off, size_t flag_off, int rank, int world, uint32_t* epoch_dev,
                          const double* count_dev, const float* const* bnw, const float* mean, const float* istd, int C,
                          float* coef, float* dbnw, float* dbnb, cudaStream_t st);
int bn3_finalize_bwd(const float* S, const float* S_local, double count, const double* count_dev, const float* const* bnw, const float* mean, const float* istd, int C,
                     float* coef, float* dbnw, float* dbnb, cudaStream_t st);
int bn3_sum_ln_fwd(const void* y1, const void* y2, const void* y3, const float* scale, const float* shift,
                   const float* lnw, const float* lnb, float eps, void* xn, float* mu, float* rstd, int N, int C, int HW,
                   cudaStream_t st);
int residual_fwd(const float* x, const void* h2, const float* gamma, const float* dp, float* out, void* out_bf16,
                 int N, int C, int HW, cudaStream_t st);
int residual_bwd_parts(int N, int C, int HW);
int residual_bwd(const float* dout, const void* h2, const float* gamma, const float* dp, void* dh2, float* dgamma_part,
                 int N, int C, int HW, cudaStream_t st);
int gelu_bwd_bias_parts(long long rows, int K);
int gelu_bwd_bias(const void* da, const void* h, void* dh, float* part, long long rows, int K, cudaStream_t st);
int bn3_sum_ln_bwd_parts(int N, int C, int HW);
int bn3_sum_ln_bwd(const void* dxn, const void* y1, const void* y2, const void* y3, const float* scale,
                   const float* shift, const float* lnw, const float* mu, const float* rstd, void* du, float* part,
                   int N, int C, int HW, cudaStream_t st);
int bn3_bwd_apply(const void* du, const void* y1, const void* y2, const void* y3, const float* coef, void* dy1,
                  void* dy2, void* dy3, int N, int C, int HW, cudaStream_t st);
}
// layernorm2d.cu
int layernorm2d_bwd_parts(int N, int HW);
int layernorm2d_fwd(const void* x, int xdt, const float* w, const float* b, float eps, void* y, int ydt, float* mean,
                    float* rstd, int N, int C, int HW, cudaStream_t st);
int layernorm2d_bwd(const void* g, int gdt, const void* x, int xdt, const float* w, const float* mean, const float* rstd,
                    void* dx, float* part, float* dw, float* db, int N, int C, int HW, cudaStream_t st);
// mask.cu
int mask_apply(float* const* w_ptrs, const float* const* m_ptrs, float* const* e_ptrs,
               const int64_t* numels, int count, int64_t max_numel, cudaStream_t st);
size_t mask_prune_workspace(int64_t n);
int mask_prune_magnitude(const float* w, float* mask, int64_t n, int64_t k, void* workspace,
                         cudaStream_t st);
int mask_grow_topk(const float* score, float* mask, int64_t n, int64_t k, void* workspace, cudaStream_t st);
int select_kth_largest_abs(const float* x, int64_t n, int64_t k, void* workspace, float* out, cudaStream_t st);
int mask_pack_bits(const float* mask, uint32_t* words, int64_t n, cudaStream_t st);
int mask_unpack_bits(const uint32_t* words, float* mask, int64_t n, cudaStream_t st);

// optim.cu
int adamw_mask_ema(float* const* p, const float* const* g, float* const* m, float* const* v, const float* const* mask,
                   float* const* ema, const int64_t* numel, const double* lr, const double* wd, const int32_t* chunk_tensor,
                   const int64_t* chunk_off, int nchunks, int chunk_elems, double beta1, double beta2, double eps,
                   double ema_decay, int64_t* step_dev, int do_adam, cudaStream_t st);

static int check_conv_args(const void* a, const void* b, const void* c, int N, int C, int H, int W,
                           int kh, int kw, int dtype, int wdtype) {
  SLAK_REQUIRE(a && b && c, SLAK_ERR_BAD_ARG, "null tensor pointer");
  SLAK_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, SLAK_ERR_BAD_ARG,
               "non-positive tensor size N=%d C=%d H=%d W=%d", N, C, H, W);
  SLAK_REQUIRE(kh > 0 && kw > 0 && (kh & 1) && (kw & 1), SLAK_ERR_BAD_ARG,
               "kernel %dx%d: both sides must be odd (same-size output needs pad=k/2)", kh, kw);
  SLAK_REQUIRE(dtype == SLAK_F32 || dtype == SLAK_F16 || dtype == SLAK_BF16, SLAK_ERR_BAD_ARG,
               "Only support fp32, fp16 and bf16, get dtype code %d", dtype);
  SLAK_REQUIRE(wdtype == SLAK_F32 || wdtype == dtype, SLAK_ERR_BAD_ARG,
               "weight dtype %d must be fp32 or equal to the activation dtype %d", wdtype, dtype);
  SLAK_REQUIRE(C <= 65535 * 32 && (long long)H * W < (1ll << 30), SLAK_ERR_UNSUPPORTED,
               "tensor too large: C=%d H=%d W=%d", C, H, W);
  return SLAK_OK;
}

}  // namespace slak

using namespace slak;

extern "C" {

SLAK_API int slak_version(void) { return 100; }

SLAK_API const char* slak_last_error(void) { return g_err; }

SLAK_API int slak_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return 0; }
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

// Tensor-core route of the single-convolution entry points: bf16 activations, fp32 taps, one kernel side equal to 5
// (K x 5, 5 x K, 5 x 5: every depthwise shape of a Decom SLaK Block) on a plane size the banded-Toeplitz kernels take.
// K x 5 runs as the transposed family alone, 5 x K / 5 x 5 as the natural family alone (csrc/dwconv_tc_dgrad.cu).
static bool single_uses_tc(int N, int C, int H, int W, int kh, int kw, int dtype, int wdtype) {
  if (dtype != SLAK_BF16 || wdtype != SLAK_F32) return false;
  if (kh != 5 && kw != 5) return false;
  const int KL = kh == 5 ? kw : kh;
  return tc::lk3_bwd_tc_supported(N, C, H, W, KL < 5 ? 5 : KL) && KL >= 5;
}
static int single_conv_tc(const void* in, const void* w, void* out, int N, int C, int H, int W, int kh, int kw, int flip,
                          cudaStream_t st) {
  if (kw == 5 && kh != 5)     // K x 5: long axis vertical -> transposed family
    return tc::lk_conv_tc(in, (const float*)w, nullptr, nullptr, nullptr, out, nullptr, nullptr, N, C, H, W, kh, 5, flip, st);
  return tc::lk_conv_tc(nullptr, nullptr, in, (const float*)w, nullptr, out, nullptr, nullptr, N, C, H, W, 5, kw, flip, st);
}

SLAK_API int slak_dwconv2d_uses_tc(int N, int C, int H, int W, int kh, int kw, int dtype, int wdtype) {
  return (N > 0 && C > 0 && single_uses_tc(N, C, H, W, kh, kw, dtype, wdtype)) ? 1 : 0;
}

SLAK_API int slak_dwconv2d_fwd(const void* x, const void* w, void* y, int N, int C, int H, int W,
                               int kh, int kw, int dtype, int wdtype, void* stream) {
  int rc = check_conv_args(x, w, y, N, C, H, W, kh, kw, dtype, wdtype);
  if (rc) return rc;
  if (single_uses_tc(N, C, H, W, kh, kw, dtype, wdtype)) return single_conv_tc(x, w, y, N, C, H, W, kh, kw, 0, (cudaStream_t)stream);
  return dwconv_simt_fwd(x, w, y, N, C, H, W, kh, kw, dtype, wdtype, /*flip=*/0, (cudaStream_t)stream);
}

SLAK_API int slak_dwconv2d_bwd_data(const void* dy, const void* w, void* dx, int N, int C, int H,
                                    int W, int kh, int kw, int dtype, int wdtype, void* stream) {
  int rc = check_conv_args(dy, w, dx, N, C, H, W, kh, kw, dtype, wdtype);
  if (rc) return rc;
  if (single_uses_tc(N, C, H, W, kh, kw, dtype, wdtype)) return single_conv_tc(dy, w, dx, N, C, H, W, kh, kw, 1, (cudaStream_t)stream);
  return dwconv_simt_fwd(dy, w, dx, N, C, H, W, kh, kw, dtype, wdtype, /*flip=*/1, (cudaStream_t)stream);
}

// the tensor-core weight gradient is the fused three-branch kernel with the one gradient tensor given for all three
// branches: the wanted branch lands in dw, the other two go to scratch behind the kernel's own workspace
static size_t single_wgrad_tc_scratch(int C, int KL) { return (((size_t)C * KL * 5 * sizeof(float)) + 255) / 256 * 256; }

SLAK_API size_t slak_dwconv2d_bwd_filter_workspace(int N, int C, int H, int W, int kh, int kw,
                                                   int dtype) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || kh <= 0 || kw <= 0) return 0;
  if (single_uses_tc(N, C, H, W, kh, kw, dtype, SLAK_F32)) {
    const int KL = kh == 5 ? kw : kh;
    return (tc::lk3_wgrad_tc_workspace(N, C, H, W, KL) + 255) / 256 * 256 + 3 * single_wgrad_tc_scratch(C, KL);
  }
  return dwconv_simt_wgrad_workspace(N, C, H, W, kh, kw);
}

SLAK_API int slak_dwconv2d_bwd_filter(const void* dy, const void* x, float* dw, int N, int C, int H,
                                      int W, int kh, int kw, int dtype, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  int rc = check_conv_args(dy, x, dw, N, C, H, W, kh, kw, dtype, dtype);
  if (rc) return rc;
  const size_t need = slak_dwconv2d_bwd_filter_workspace(N, C, H, W, kh, kw, dtype);
  SLAK_REQUIRE(workspace && workspace_bytes >= need, SLAK_ERR_WORKSPACE,
               "bwd_filter workspace too small: %zu < %zu bytes", workspace_bytes, need);
  if (single_uses_tc(N, C, H, W, kh, kw, dtype, SLAK_F32)) {
    const int KL = kh == 5 ? kw : kh;
    const size_t w0 = (tc::lk3_wgrad_tc_workspace(N, C, H, W, KL) + 255) / 256 * 256, sc = single_wgrad_tc_scratch(C, KL);
    uint8_t* wsb = (uint8_t*)workspace;
    float* t1 = (float*)(wsb + w0); float* t2 = (float*)(wsb + w0 + sc); float* t3 = (float*)(wsb + w0 + 2 * sc);
    float* o1 = t1; float* o2 = t2; float* o3 = t3;
    if (kw == 5 && kh != 5) o1 = dw;            // K x 5
    else if (kh == 5 && kw != 5) o2 = dw;       // 5 x K
    else o3 = dw;                               // 5 x 5
    return tc::lk3_wgrad_tc(x, dy, dy, dy, o1, o2, o3, N, C, H, W, KL, workspace, (cudaStream_t)stream);
  }
  return dwconv_simt_wgrad(dy, x, dw, N, C, H, W, kh, kw, dtype, workspace, (cudaStream_t)stream);
}

SLAK_API int slak_lk_branches_uses_tc(int N, int C, int H, int W, int KL, int KS, int dtype) {
  return (dtype == SLAK_BF16 && KS == 5 && N > 0 && C > 0 && tc::lk3_tc_supported(N, C, H, W, KL)) ? 1 : 0;
}

SLAK_API int slak_lk_branches_fwd(const void* x, const float* w1, const float* w2, const float* w3,
                                  void* y1, void* y2, void* y3, int N, int C, int H, int W, int KL,
                                  int KS, int dtype, void* stream) {
  int rc = check_conv_args(x, w1, y1, N, C, H, W, KL, KS, dtype, SLAK_F32);
  if (rc) return rc;
  SLAK_REQUIRE(w2 && y2, SLAK_ERR_BAD_ARG, "null tensor pointer");
  SLAK_REQUIRE((w3 == nullptr) == (y3 == nullptr), SLAK_ERR_BAD_ARG, "w3 and y3 must both be given or both be NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (w3 && slak_lk_branches_uses_tc(N, C, H, W, KL, KS, dtype))
    return tc::lk3_fwd_tc(x, w1, w2, w3, y1, y2, y3, N, C, H, W, KL, nullptr, st);
  rc = dwconv_simt_fwd(x, w1, y1, N, C, H, W, KL, KS, dtype, SLAK_F32, 0, st);
  if (rc) return rc;
  rc = dwconv_simt_fwd(x, w2, y2, N, C, H, W, KS, KL, dtype, SLAK_F32, 0, st);
  if (rc || !w3) return rc;
  return dwconv_simt_fwd(x, w3, y3, N, C, H, W, KS, KS, dtype, SLAK_F32, 0, st);
}

SLAK_API int slak_lk_branches_bwd_uses_tc(int N, int C, int H, int W, int KL, int KS, int dtype) {
  return (dtype == SLAK_BF16 && KS == 5 && N > 0 && C > 0 && tc::lk3_bwd_tc_supported(N, C, H, W, KL)) ? 1 : 0;
}

SLAK_API int slak_lk_branches_bwd_data(const void* dy1, const void* dy2, const void* dy3, const float* w1,
                                       const float* w2, const float* w3, void* dx, void* tmp, int N, int C,
                                       int H, int W, int KL, int KS, int dtype, void* stream) {
  int rc = check_conv_args(dy1, w1, dx, N, C, H, W, KL, KS, dtype, SLAK_F32);
  if (rc) return rc;
  SLAK_REQUIRE(dy2 && dy3 && w2 && w3 && tmp, SLAK_ERR_BAD_ARG, "null tensor pointer");
  SLAK_REQUIRE(slak_lk_branches_bwd_uses_tc(N, C, H, W, KL, KS, dtype), SLAK_ERR_UNSUPPORTED,
               "fused bwd_data covers only the tensor-core shapes (see slak_lk_branches_uses_tc)");
  cudaStream_t st = (cudaStream_t)stream;
  rc = tc::lk_conv_tc(nullptr, nullptr, dy3, w3, nullptr, tmp, nullptr, nullptr, N, C, H, W, KL, KS, /*flip=*/1, st);
  if (rc) return rc;
  return tc::lk_conv_tc(dy1, w1, dy2, w2, tmp, dx, nullptr, nullptr, N, C, H, W, KL, KL, /*flip=*/1, st);
}

SLAK_API int slak_lk_branches_bwd_data_f32(const void* dy1, const void* dy2, const void* dy3, const float* w1,
                                           const float* w2, const float* w3, const float* addend, float* dx,
                                           void* tmp, int N, int C, int H, int W, int KL, int KS, void* stream) {
  int rc = check_conv_args(dy1, w1, dx, N, C, H, W, KL, KS, SLAK_BF16, SLAK_F32);
  if (rc) return rc;
  SLAK_REQUIRE(dy2 && dy3 && w2 && w3 && tmp, SLAK_ERR_BAD_ARG, "null tensor pointer");
  SLAK_REQUIRE(slak_lk_branches_bwd_uses_tc(N, C, H, W, KL, KS, SLAK_BF16), SLAK_ERR_UNSUPPORTED,
               "fused bwd_data covers only the tensor-core shapes (see slak_lk_branches_uses_tc)");
  cudaStream_t st = (cudaStream_t)stream;
  if (KS == 5 && addend && tc::dense::supported(N, C, H, W, KL))      // small planes: one dense-GEMM launch for the three branches
    return tc::dense::dgrad(dy1, dy2, dy3, w1, w2, w3, addend, dx, N, C, H, W, KL, st);
  rc = tc::lk_conv_tc(nullptr, nullptr, dy3, w3, nullptr, tmp, nullptr, nullptr, N, C, H, W, KL, KS, /*flip=*/1, st);
  if (rc) return rc;
  return tc::lk_conv_tc(dy1, w1, dy2, w2, tmp, nullptr, addend, dx, N, C, H, W, KL, KL, /*flip=*/1, st);
}

// Inference form of the Decom large-kernel layer after re-parameterisation (three BatchNorms folded, 5 x 5 merged into
// the 5 x K kernel): y = dwconv_{KL x 5}(x, wv) + dwconv_{5 x KL}(x, wh) + bias, x read once, y written once.
SLAK_API int slak_lk_merged_fwd(const void* x, const float* wv, const float* wh, const float* bias, void* y, int N, int C,
                                int H, int W, int KL, int dtype, void* stream) {
  int rc = check_conv_args(x, wv, y, N, C, H, W, KL, 5, dtype, SLAK_F32);
  if (rc) return rc;
  SLAK_REQUIRE(wh, SLAK_ERR_BAD_ARG, "null tensor pointer");
  SLAK_REQUIRE(slak_lk_branches_bwd_uses_tc(N, C, H, W, KL, 5, dtype), SLAK_ERR_UNSUPPORTED,
               "slak_lk_merged_fwd covers only the tensor-core shapes (see slak_lk_branches_uses_tc)");
  return tc::lk_conv_tc(x, wv, x, wh, nullptr, y, nullptr, nullptr, N, C, H, W, KL, KL, /*flip=*/0, (cudaStream_t)stream, bias);
}

SLAK_API size_t slak_lk_branches_bwd_filter_workspace(int N, int C, int H, int W, int KL, int KS) {
  (void)KS;
  if (N <= 0 || C <= 0 || KL <= 0) return 0;
  return tc::lk3_wgrad_tc_workspace(N, C, H, W, KL);
}

SLAK_API int slak_lk_branches_bwd_filter(const void* x, const void* dy1, const void* dy2, const void* dy3,
                                         float* dw1, float* dw2, float* dw3, int N, int C, int H, int W, int KL,
                                         int KS, int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_conv_args(x, dy1, dw1, N, C, H, W, KL, KS, dtype, dtype);
  if (rc) return rc;
  SLAK_REQUIRE(dy2 && dy3 && dw2 && dw3, SLAK_ERR_BAD_ARG, "null tensor pointer");
  SLAK_REQUIRE(slak_lk_branches_bwd_uses_tc(N, C, H, W, KL, KS, dtype), SLAK_ERR_UNSUPPORTED,
               "fused bwd_filter covers only the tensor-core shapes (see slak_lk_branches_uses_tc)");
  const size_t need = tc::lk3_wgrad_tc_workspace(N, C, H, W, KL);
  SLAK_REQUIRE(workspace && workspace_bytes >= need, SLAK_ERR_WORKSPACE, "bwd_filter workspace too small: %zu < %zu",
               workspace_bytes, need);
  return tc::lk3_wgrad_tc(x, dy1, dy2, dy3, dw1, dw2, dw3, N, C, H, W, KL, workspace, (cudaStream_t)stream);
}

// ---- fused Block glue (block_fused.cu) --------------------------------------------------------
SLAK_API size_t slak_block_conv_fwd_workspace(int N, int C, int H, int W) {
  const int s = tc::lk3_fwd_tc_splits(N, C, H, W);
  return s <= 0 ? 0 : (size_t)C * s * 6 * sizeof(float);
}

SLAK_API int slak_block_conv_fwd(const void* x, const float* w1, const float* w2, const float* w3, void* y1, void* y2,
                                 void* y3, double* sums, void* workspace, size_t workspace_bytes, int N, int C, int H,
                                 int W, int KL, void* stream) {
  int rc = check_conv_args(x, w1, y1, N, C, H, W, KL, 5, SLAK_BF16, SLAK_F32);
  if (rc) return rc;
  SLAK_REQUIRE(w2 && w3 && y2 && y3 && sums, SLAK_ERR_BAD_ARG, "null tensor pointer");
  SLAK_REQUIRE(slak_lk_branches_uses_tc(N, C, H, W, KL, 5, SLAK_BF16), SLAK_ERR_UNSUPPORTED,
               "slak_block_conv_fwd covers only the tensor-core shapes");
  const size_t need = slak_block_conv_fwd_workspace(N, C, H, W);
  SLAK_REQUIRE(workspace && workspace_bytes >= need, SLAK_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  rc = tc::lk3_fwd_tc(x, w1, w2, w3, y1, y2, y3, N, C, H, W, KL, (float*)workspace, st);
  if (rc) return rc;
  return blk::bn3_stats_finalize((const float*)workspace, tc::lk3_fwd_tc_splits(N, C, H, W), sums, C, st);
}

SLAK_API int slak_bn3_finalize_fwd(const double* sums, double count, const double* count_dev, const float* const* bnw, const float* const* bnb,
                                   float* const* rmean, float* const* rvar, float eps, float momentum, int C,
                                   float* scale, float* shift, float* mean, float* istd, void* stream) {
  SLAK_REQUIRE(sums && bnw && bnb && rmean && rvar && scale && shift && mean && istd && C > 0 && (count > 0 || count_dev), SLAK_ERR_BAD_ARG, "bad argument");
  for (int i = 0; i < 3; ++i) SLAK_REQUIRE(bnw[i] && bnb[i], SLAK_ERR_BAD_ARG, "null BN parameter");
  return blk::bn3_finalize_fwd(sums, count, count_dev, bnw, bnb, rmean, rvar, eps, momentum, C, scale, shift, mean, istd, (cudaStream_t)stream);
}

SLAK_API int slak_bn3_finalize_fwd_sync(const void* const* peer_bases, size_t slot_off, size_t flag_off, int rank, int world,
                                        uint32_t* epoch_dev, const float* const* bnw, const float* const* bnb,
                                        float* const* rmean, float* const* rvar, float eps, float momentum, int C, float* scale,
                                        float* shift, float* mean, float* istd, void* stream) {
  SLAK_REQUIRE(peer_bases && epoch_dev && bnw && bnb && rmean && rvar && scale && shift && mean && istd && C > 0, SLAK_ERR_BAD_ARG, "bad argument");
  SLAK_REQUIRE(rank >= 0 && rank < world, SLAK_ERR_BAD_ARG, "rank %d outside world %d", rank, world);
  for (int i = 0; i < 3; ++i) SLAK_REQUIRE(bnw[i] && bnb[i], SLAK_ERR_BAD_ARG, "null BN parameter");
  return blk::bn3_finalize_fwd_sync(peer_bases, slot_off, flag_off, rank, world, epoch_dev, bnw, bnb, rmean, rvar, eps, momentum, C,
                                    scale, shift, mean, istd, (cudaStream_t)stream);
}

SLAK_API int slak_bn3_finalize_bwd_sync(const void* const* peer_bases, size_t slot_off, size_t flag_off, int rank, int world,
                                        uint32_t* epoch_dev, const double* count_dev, const float* const* bnw, const float* mean,
                                        const float* istd, int C, float* coef, float* dbnw, float* dbnb, void* stream) {
  SLAK_REQUIRE(peer_bases && epoch_dev && count_dev && bnw && bnw[0] && bnw[1] && bnw[2] && mean && istd && coef && dbnw && dbnb && C > 0,
               SLAK_ERR_BAD_ARG, "bad argument");
  SLAK_REQUIRE(rank >= 0 && rank < world, SLAK_ERR_BAD_ARG, "rank %d outside world %d", rank, world);
  return blk::bn3_finalize_bwd_sync(peer_bases, slot_off, flag_off, rank, world, epoch_dev, count_dev, bnw, mean, istd, C, coef,
                                    dbnw, dbnb, (cudaStream_t)stream);
}

SLAK_API int slak_bn3_eval_affine(const float* const* bnw, const float* const* bnb, const float* const* rmean,
                                  const float* const* rvar, float eps, int C, float* scale, float* shift, void* stream) {
  SLAK_REQUIRE(bnw && bnb && rmean && rvar && scale && shift && C > 0, SLAK_ERR_BAD_ARG, "bad argument");
  for (int i = 0; i < 3; ++i) SLAK_REQUIRE(bnw[i] && bnb[i] && rmean[i] && rvar[i], SLAK_ERR_BAD_ARG, "null BN tensor");
  return blk::bn3_eval_affine(bnw, bnb, rmean, rvar, eps, C, scale, shift, (cudaStream_t)stream);
}

SLAK_API int slak_bn3_sum_ln_fwd(const void* y1, const void* y2, const void* y3, const float* scale, const float* shift,
                                 const float* lnw, const float* lnb, float eps, void* xn, float* mu, float* rstd, int N,
                                 int C, int HW, void* stream) {
  SLAK_REQUIRE(y1 && y2 && y3 && scale && shift && lnw && lnb && xn && mu && rstd, SLAK_ERR_BAD_ARG, "null tensor pointer");
  SLAK_REQUIRE(N > 0 && C > 0 && HW > 0, SLAK_ERR_BAD_ARG, "non-positive size");
  return blk::bn3_sum_ln_fwd(y1, y2, y3, scale, shift, lnw, lnb, eps, xn, mu, rstd, N, C, HW, (cudaStream_t)stream);
}

SLAK_API int slak_block_residual_fwd(const float* x, const void* h2, const float* gamma, const float* dp, float* out,
                                     void* out_bf16, int N, int C, int HW, void* stream) {
  SLAK_REQUIRE(x && h2 && gamma && out && N > 0 && C > 0 && HW > 0, SLAK_ERR_BAD_ARG, "bad argument");
  return blk::residual_fwd(x, h2, gamma, dp, out, out_bf16, N, C, HW, (cudaStream_t)stream);
}

SLAK_API int slak_ln2d_patch_fwd(const float* x, const float* lnw, const float* lnb, float eps, void* A, float* mean,
                                 float* rstd, int N, int C, int H, int W, void* stream) {
  SLAK_REQUIRE(x && lnw && lnb && A && mean && rstd && N > 0 && C > 0 && H > 0 && W > 0, SLAK_ERR_BAD_ARG, "bad argument");
  return blk::g2::ln2d_patch_fwd(x, lnw, lnb, eps, A, mean, rstd, N, C, H, W, (cudaStream_t)stream);
}
SLAK_API int slak_ln2d_patch_bwd_parts(int N, int C, int H, int W) { return blk::g2::ln2d_patch_bwd_parts(N, C, H, W); }
SLAK_API int slak_ln2d_patch_bwd(const void* dA, const float* x, const float* lnw, const float* mean, const float* rstd,
                                 float* dx, float* part, int N, int C, int H, int W, void* stream) {
  SLAK_REQUIRE(dA && x && lnw && mean && rstd && dx && part && N > 0 && C > 0 && H > 0 && W > 0, SLAK_ERR_BAD_ARG, "bad argument");
  return blk::g2::ln2d_patch_bwd(dA, x, lnw, mean, rstd, dx, part, N, C, H, W, (cudaStream_t)stream);
}
SLAK_API int slak_nhwc_to_nchw(const void* src_bf16, float* dst, void* dst_bf16, int N, int C, int HW, void* stream) {
  SLAK_REQUIRE(src_bf16 && dst && N > 0 && C > 0 && HW > 0, SLAK_ERR_BAD_ARG, "bad argument");
  return blk::g2::nhwc_to_nchw(src_bf16, dst, dst_bf16, N, C, HW, (cudaStream_t)stream);
}
SLAK_API int slak_nchw_to_nhwc_parts(int N, int C, int HW) { return blk::g2::nchw_to_nhwc_parts(N, C, HW); }
SLAK_API int slak_nchw_to_nhwc(const float* src, void* dst_bf16, float* part, int N, int C, int HW, void* stream) {
  SLAK_REQUIRE(src && dst_bf16 && part && N > 0 && C > 0 && HW > 0, SLAK_ERR_BAD_ARG, "bad argument");
  return blk::g2::nchw_to_nhwc(src, dst_bf16, part, N, C, HW, (cudaStream_t)stream);
}

SLAK_API int slak_patchify4(const float* x, void* A, int N, int Cin, int H, int W, void* stream) {
  SLAK_REQUIRE(x && A && N > 0 && Cin > 0 && H > 0 && W > 0, SLAK_ERR_BAD_ARG, "bad argument");
  return blk::g2::patchify4(x, A, N, Cin, H, W, (cudaStream_t)stream);
}
SLAK_API int slak_ln_rows_fwd(const void* Y, const float* lnw, const float* lnb, float eps, float* out, void* out_bf16,
                              float* mean, float* rstd, int N, int C, int HW, void* stream) {
  SLAK_REQUIRE(Y && lnw && lnb && out && mean && rstd && N > 0 && C > 0 && HW > 0, SLAK_ERR_BAD_ARG, "bad argument");
  return blk::g2::ln_rows_fwd(Y, lnw, lnb, eps, out, out_bf16, mean, rstd, N, C, HW, (cudaStream_t)stream);
}
SLAK_API int slak_ln_rows_bwd_parts(int N, int C, int HW) { return blk::g2::ln_rows_bwd_parts(N, C, HW); }
SLAK_API int slak_ln_rows_bwd(const float* dout, const void* Y, const float* lnw, const float* mean, const float* rstd,
                              void* dY, float* part, int N, int C, int HW, void* stream) {
  SLAK_REQUIRE(dout && Y && lnw && mean && rstd && dY && part && N > 0 && C > 0 && HW > 0, SLAK_ERR_BAD_ARG, "bad argument");
  return blk::g2::ln_rows_bwd(dout, Y, lnw, mean, rstd, dY, part, N, C, HW, (cudaStream_t)stream);
}

SLAK_API int slak_block_residual_bwd_parts(int N, int C, int HW) { return blk::residual_bwd_parts(N, C, HW); }

SLAK_API int slak_block_residual_bwd(const float* dout, const void* h2, const float* gamma, const float* dp, void* dh2,
                                     float* dgamma_part, int N, int C, int HW, void* stream) {
  SLAK_REQUIRE(dout && h2 && gamma && dh2 && dgamma_part && N > 0 && C > 0 && HW > 0, SLAK_ERR_BAD_ARG, "bad argument");
  return blk::residual_bwd(dout, h2, gamma, dp, dh2, dgamma_part, N, C, HW, (cudaStream_t)stream);
}

SLAK_API int slak_gelu_bwd_bias_parts(int64_t rows, int K) { return blk::gelu_bwd_bias_parts(rows, K); }

SLAK_API int slak_gelu_bwd_bias(const void* da, const void* h, void* dh, float* part, int64_t rows, int K, void* stream) {
  SLAK_REQUIRE(da && h && dh && part && rows > 0 && K > 0, SLAK_ERR_BAD_ARG, "bad argument");
  return blk::gelu_bwd_bias(da, h, dh, part, rows, K, (cudaStream_t)stream);
}

SLAK_API int slak_bn3_sum_ln_bwd_parts(int N, int C, int HW) { return blk::bn3_sum_ln_bwd_parts(N, C, HW); }

SLAK_API int slak_bn3_sum_ln_bwd(const void* dxn, const void* y1, const void* y2, const void* y3, const float* scale,
                                 const float* shift, const float* lnw, const float* mu, const float* rstd, void* du,
                                 float* part, int N, int C, int HW, void* stream) {
  SLAK_REQUIRE(dxn && y1 && y2 && y3 && scale && shift && lnw && mu && rstd && du && part, SLAK_ERR_BAD_ARG, "null tensor pointer");
  SLAK_REQUIRE(N > 0 && C > 0 && HW > 0, SLAK_ERR_BAD_ARG, "non-positive size");
  return blk::bn3_sum_ln_bwd(dxn, y1, y2, y3, scale, shift, lnw, mu, rstd, du, part, N, C, HW, (cudaStream_t)stream);
}

SLAK_API int slak_bn3_finalize_bwd(const float* S, const float* S_local, double count, const double* count_dev, const float* const* bnw, const float* mean, const float* istd,
                                   int C, float* coef, float* dbnw, float* dbnb, void* stream) {
  SLAK_REQUIRE(S && bnw && bnw[0] && bnw[1] && bnw[2] && mean && istd && coef && dbnw && dbnb && C > 0 && (count > 0 || count_dev), SLAK_ERR_BAD_ARG, "bad argument");
  return blk::bn3_finalize_bwd(S, S_local, count, count_dev, bnw, mean, istd, C, coef, dbnw, dbnb, (cudaStream_t)stream);
}

SLAK_API int slak_bn3_bwd_apply(const void* du, const void* y1, const void* y2, const void* y3, const float* coef,
                                void* dy1, void* dy2, void* dy3, int N, int C, int HW, void* stream) {
  SLAK_REQUIRE(du && y1 && y2 && y3 && coef && dy1 && dy2 && dy3 && N > 0 && C > 0 && HW > 0, SLAK_ERR_BAD_ARG, "bad argument");
  return blk::bn3_bwd_apply(du, y1, y2, y3, coef, dy1, dy2, dy3, N, C, HW, (cudaStream_t)stream);
}

SLAK_API int slak_mask_apply(float* const* w_ptrs, const float* const* mask_ptrs,
                             float* const* extra_ptrs, const int64_t* numels, int count,
                             int64_t max_numel, void* stream) {
  SLAK_REQUIRE(count >= 0, SLAK_ERR_BAD_ARG, "negative tensor count");
  if (count == 0) return SLAK_OK;
  SLAK_REQUIRE(w_ptrs && mask_ptrs && numels, SLAK_ERR_BAD_ARG, "null pointer table");
  SLAK_REQUIRE(count <= 65535, SLAK_ERR_UNSUPPORTED, "too many tensors in one launch: %d", count);
  return mask_apply(w_ptrs, mask_ptrs, extra_ptrs, numels, count, max_numel, (cudaStream_t)stream);
}

SLAK_API int slak_adamw_mask_ema_step(float* const* p, const float* const* g, float* const* m, float* const* v,
                                      const float* const* mask, float* const* ema, const int64_t* numel, const double* lr,
                                      const double* wd, const int32_t* chunk_tensor, const int64_t* chunk_off, int nchunks,
                                      int chunk_elems, double beta1, double beta2, double eps, double ema_decay,
                                      int64_t* step_dev, int do_adam, void* stream) {
  SLAK_REQUIRE(nchunks >= 0 && chunk_elems > 0, SLAK_ERR_BAD_ARG, "bad chunk list");
  if (nchunks == 0) return SLAK_OK;
  SLAK_REQUIRE(p && numel && chunk_tensor && chunk_off && step_dev, SLAK_ERR_BAD_ARG, "null table");
  SLAK_REQUIRE(!do_adam || (g && m && v && lr && wd), SLAK_ERR_BAD_ARG, "AdamW needs gradient / moment / lr / wd tables");
  SLAK_REQUIRE(do_adam || mask || ema, SLAK_ERR_BAD_ARG, "nothing to do");
  return adamw_mask_ema(p, g, m, v, mask, ema, numel, lr, wd, chunk_tensor, chunk_off, nchunks, chunk_elems, beta1, beta2,
                        eps, ema_decay, step_dev, do_adam, (cudaStream_t)stream);
}

SLAK_API size_t slak_mask_prune_workspace(int64_t numel) { return mask_prune_workspace(numel); }

SLAK_API int slak_mask_prune_magnitude(const float* w, float* mask, int64_t numel, int64_t k,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  SLAK_REQUIRE(w && mask, SLAK_ERR_BAD_ARG, "null tensor pointer");
  SLAK_REQUIRE(numel >= 0 && numel < (1ll << 32), SLAK_ERR_UNSUPPORTED, "numel %lld out of range",
               (long long)numel);
  SLAK_REQUIRE(workspace && workspace_bytes >= mask_prune_workspace(numel), SLAK_ERR_WORKSPACE,
               "prune workspace too small");
  return mask_prune_magnitude(w, mask, numel, k, workspace, (cudaStream_t)stream);
}

SLAK_API int slak_mask_grow_topk(const float* score, float* mask, int64_t numel, int64_t k, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  SLAK_REQUIRE(score && mask, SLAK_ERR_BAD_ARG, "null tensor pointer");
  SLAK_REQUIRE(numel >= 0 && numel < (1ll << 32), SLAK_ERR_UNSUPPORTED, "numel %lld out of range", (long long)numel);
  SLAK_REQUIRE(workspace && workspace_bytes >= mask_prune_workspace(numel), SLAK_ERR_WORKSPACE, "select workspace too small");
  return mask_grow_topk(score, mask, numel, k, workspace, (cudaStream_t)stream);
}

SLAK_API int slak_select_kth_largest_abs(const float* x, int64_t numel, int64_t k, void* workspace, size_t workspace_bytes,
                                         float* out, void* stream) {
  SLAK_REQUIRE(x && out, SLAK_ERR_BAD_ARG, "null tensor pointer");
  SLAK_REQUIRE(numel > 0 && numel < (1ll << 32) && k >= 1 && k <= numel, SLAK_ERR_BAD_ARG, "need 1 <= k <= numel < 2^32");
  SLAK_REQUIRE(workspace && workspace_bytes >= mask_prune_workspace(numel), SLAK_ERR_WORKSPACE, "select workspace too small");
  return select_kth_largest_abs(x, numel, k, workspace, out, (cudaStream_t)stream);
}

SLAK_API int slak_mask_pack_bits(const float* mask, uint32_t* words, int64_t numel, void* stream) {
  SLAK_REQUIRE(mask && words && numel >= 0, SLAK_ERR_BAD_ARG, "bad argument");
  return mask_pack_bits(mask, words, numel, (cudaStream_t)stream);
}

SLAK_API int slak_mask_unpack_bits(const uint32_t* words, float* mask, int64_t numel, void* stream) {
  SLAK_REQUIRE(mask && words && numel >= 0, SLAK_ERR_BAD_ARG, "bad argument");
  return mask_unpack_bits(words, mask, numel, (cudaStream_t)stream);
}

}  // extern "C"

// ---- LayerNorm over channels of NCHW (models/SLaK.py:256-261) -------------------------------------
SLAK_API int slak_layernorm2d_fwd(const void* x, int x_dtype, const float* w, const float* b, float eps, void* y,
                                  int y_dtype, float* mean, float* rstd, int N, int C, int HW, void* stream) {
  SLAK_REQUIRE(x && w && b && y, SLAK_ERR_BAD_ARG, "null tensor pointer");
  SLAK_REQUIRE((mean == nullptr) == (rstd == nullptr), SLAK_ERR_BAD_ARG, "mean and rstd must be given together");
  SLAK_REQUIRE(N >= 0 && C > 0 && HW >= 0, SLAK_ERR_BAD_ARG, "bad size N=%d C=%d HW=%d", N, C, HW);
  return layernorm2d_fwd(x, x_dtype, w, b, eps, y, y_dtype, mean, rstd, N, C, HW, (cudaStream_t)stream);
}

SLAK_API int slak_layernorm2d_bwd_parts(int N, int HW) { return layernorm2d_bwd_parts(N, HW); }

SLAK_API int slak_layernorm2d_bwd(const void* g, int g_dtype, const void* x, int x_dtype, const float* w, const float* mean,
                                  const float* rstd, void* dx, float* part, float* dw, float* db, int N, int C, int HW,
                                  void* stream) {
  SLAK_REQUIRE(g && x && w && mean && rstd && dx && part && dw && db, SLAK_ERR_BAD_ARG, "null tensor pointer");
  SLAK_REQUIRE(N >= 0 && C > 0 && HW >= 0, SLAK_ERR_BAD_ARG, "bad size N=%d C=%d HW=%d", N, C, HW);
  return layernorm2d_bwd(g, g_dtype, x, x_dtype, w, mean, rstd, dx, part, dw, db, N, C, HW, (cudaStream_t)stream);
}

// ---- fixed-order column sum of per-CTA partial rows --------------------------------------------
SLAK_API int slak_colsum_f32(const float* part, int rows, int cols, float* out, void* stream) {
  SLAK_REQUIRE(part && out && rows > 0 && cols > 0, SLAK_ERR_BAD_ARG, "bad argument");
  return blk::colsum(part, rows, cols, out, (cudaStream_t)stream);
}

SLAK_API int slak_cast_transpose_bf16(const float* w, void* wb, void* wt, int R, int Cc, void* stream) {
  SLAK_REQUIRE(w && wb && wt && R > 0 && Cc > 0, SLAK_ERR_BAD_ARG, "bad argument");
  return blk::cast_transpose(w, wb, wt, R, Cc, (cudaStream_t)stream);
}

// ---- pointwise MLP GEMMs with fused epilogues (csrc/mlp_tc.cu) -----------------------------------------
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

SLAK_API int slak_mlp_parts(int M, int N) { return tc::mlp_parts(M, N); }

SLAK_API int slak_mlp_gemm_nt(int epi, const void* a, const void* b, const float* bias, const void* aux_h, void* out0,
                              void* out1, float* colpart, int M, int N, int K, void* stream) {
  SLAK_REQUIRE(epi >= 0 && epi <= 3, SLAK_ERR_BAD_ARG, "unknown epilogue %d", epi);
  SLAK_REQUIRE(a && b, SLAK_ERR_BAD_ARG, "null operand");
  SLAK_REQUIRE(aligned16(a) && aligned16(b) && aligned16(bias) && aligned16(aux_h) && aligned16(out0) && aligned16(out1),
               SLAK_ERR_BAD_ARG, "operands must be 16-byte aligned");
  if (epi == 0) SLAK_REQUIRE(bias && out1, SLAK_ERR_BAD_ARG, "FC1 needs bias and the activation output");
  if (epi == 1) SLAK_REQUIRE(bias && out0, SLAK_ERR_BAD_ARG, "BIAS needs bias and an output");
  if (epi == 2) SLAK_REQUIRE(aux_h && out0 && colpart, SLAK_ERR_BAD_ARG, "DGELU needs H, an output and the column partials");
  if (epi == 3) SLAK_REQUIRE(out0, SLAK_ERR_BAD_ARG, "PLAIN needs an output");
  return tc::mlp_gemm_nt(epi, a, b, bias, aux_h, out0, out1, colpart, M, N, K, (cudaStream_t)stream);
}

SLAK_API int slak_mlp_fc1_gelu_fwd(const void* x, const void* w, const float* bias, void* h, void* a, int M, int N, int K,
                                   void* stream) {
  return slak_mlp_gemm_nt(0, x, w, bias, nullptr, h, a, nullptr, M, N, K, stream);
}

SLAK_API int slak_mlp_fc2_dgelu_bwd(const void* g, const void* wt, const void* h, void* dh, float* colpart, int M, int N,
                                    int K, void* stream) {
  return slak_mlp_gemm_nt(2, g, wt, nullptr, h, dh, nullptr, colpart, M, N, K, stream);
}

SLAK_API int slak_mlp_wgrad_splits(int M, int Ma, int Nb) { return tc::mlp_wgrad_splits(M, Ma, Nb); }

SLAK_API int slak_mlp_gemm_tn_splitk(const void* p, const void* q, float* part, int M, int Ma, int Nb, void* stream) {
  SLAK_REQUIRE(p && q && part, SLAK_ERR_BAD_ARG, "null tensor pointer");
  SLAK_REQUIRE(aligned16(p) && aligned16(q) && aligned16(part), SLAK_ERR_BAD_ARG, "operands must be 16-byte aligned");
  return tc::mlp_gemm_tn_splitk(p, q, part, M, Ma, Nb, (cudaStream_t)stream);
}
