// Launchers of the register-resident glue kernels (block_glue2.cu); block_fused.cu tries them first and falls back to its
// shared-memory-tile kernels when they return SLAK_G2_UNSUPPORTED (C % 8 != 0, C > 768, misaligned NHWC tensors, or
// SLAK_GLUE_V1=1 in the environment).
#pragma once
#include <cuda_runtime.h>

#define SLAK_G2_UNSUPPORTED (-1000)

namespace slak {
namespace blk {
namespace g2 {
bool supported(int N, int C, int HW);
int ln_fwd(const void* y1, const void* y2, const void* y3, const float* scale, const float* shift, const float* lnw,
           const float* lnb, float eps, void* xn, float* mu, float* rstd, int N, int C, int HW, cudaStream_t st);
int res_fwd(const float* x, const void* h2, const float* gamma, const float* dp, float* out, void* out_bf16, int N, int C,
            int HW, cudaStream_t st);
int res_bwd_parts(int N, int C, int HW);      // 0 when unsupported
int res_bwd(const float* dout, const void* h2, const float* gamma, const float* dp, void* dh2, float* part, int N, int C,
            int HW, cudaStream_t st);
int ln_bwd_parts(int N, int C, int HW);       // 0 when unsupported
int ln_bwd(const void* dxn, const void* y1, const void* y2, const void* y3, const float* scale, const float* shift,
           const float* lnw, const float* mu, const float* rstd, void* du, float* part, int N, int C, int HW, cudaStream_t st);
// downsampling layer (LayerNorm2d + 2 x 2 stride-2 convolution as a GEMM over patch rows)
int ln2d_patch_fwd(const float* x, const float* lnw, const float* lnb, float eps, void* A, float* mu, float* rstd, int N, int C,
                   int H, int W, cudaStream_t st);
int ln2d_patch_bwd_parts(int N, int C, int H, int W);
int ln2d_patch_bwd(const void* dA, const float* x, const float* lnw, const float* mu, const float* rstd, float* dx, float* part,
                   int N, int C, int H, int W, cudaStream_t st);
int nhwc_to_nchw(const void* h, float* out, void* out_bf16, int N, int C, int HW, cudaStream_t st);
int nchw_to_nhwc_parts(int N, int C, int HW);
int nchw_to_nhwc(const float* src, void* dst_bf16, float* part, int N, int C, int HW, cudaStream_t st);
// stem (4 x 4 stride-4 convolution as a GEMM over patch rows, then LayerNorm over the channels of each token row)
int patchify4(const float* x, void* A, int N, int Cin, int H, int W, cudaStream_t st);
int ln_rows_fwd(const void* Y, const float* lnw, const float* lnb, float eps, float* out, void* out_bf16, float* mu, float* rstd,
                int N, int C, int HW, cudaStream_t st);
int ln_rows_bwd_parts(int N, int C, int HW);
int ln_rows_bwd(const float* dout, const void* Y, const float* lnw, const float* mu, const float* rstd, void* dY, float* part,
                int N, int C, int HW, cudaStream_t st);
}  // namespace g2
}  // namespace blk
}  // namespace slak
