// oracle/ref_driver.cu -- TEST INFRASTRUCTURE.  A 3-function C shim that CALLS the
// reference's own host implementation where it lies under /root/reference:
//   cutlass/tools/util/include/cutlass/util/reference/host/convolution.h
//     Depsep_Fprop :237, Depsep_Dgrad :332, Depsep_Wgrad :431
// with the exact problem description the torch extension builds
// (forward_fp32.cu:135-144,221-241: NCHW, pad=k/2, stride 1, cross-correlation, alpha=1).
// Built by oracle/Makefile into oracle/_ref/libslak_ref.so (git-ignored, host-only code
// compiled with nvcc because the fork's headers need CUTLASS_DEVICE macros); used to pin
// the C restatement in oracle/dwconv_oracle.c.  No reference source is copied here.
#include "cutlass/cutlass.h"
#include "cutlass/conv/conv2d_problem_size.h"
#include "cutlass/layout/tensor.h"
#include "cutlass/tensor_ref.h"
#include "cutlass/util/reference/host/convolution.h"

using L = cutlass::layout::TensorNCHW;
using Ref = cutlass::TensorRef<float, L>;

static cutlass::conv::Conv2dProblemSize problem(int N, int G, int H, int W, int R, int S) {
  return cutlass::conv::Conv2dProblemSize(N, H, W, G, G, R, S, H, W, R / 2, S / 2, 1, 1, 1, 1,
                                          cutlass::conv::Mode::kCrossCorrelation, 1, G);
}

extern "C" {

void ref_dwconv_fwd(const float* x, const float* f, float* y, int N, int G, int H, int W, int R, int S) {
  auto ps = problem(N, G, H, W, R, S);
  Ref tx(const_cast<float*>(x), L::packed({N, H, W, G}));
  Ref tf(const_cast<float*>(f), L::packed({G, R, S, 1}));
  Ref ty(y, L::packed({N, H, W, G}));
  Ref none(nullptr, L::packed({1, 1, 1, G}));
  cutlass::reference::host::Depsep_Fprop<float, L, float, L, float, L, float, L, float, float>(
      ps, tx, tf, none, none, ty, 1.f, 0.f, 0.f);
}

void ref_dwconv_bwd_data(const float* dy, const float* f, float* dx, int N, int G, int H, int W, int R, int S) {
  auto ps = problem(N, G, H, W, R, S);
  Ref tdy(const_cast<float*>(dy), L::packed({N, H, W, G}));
  Ref tf(const_cast<float*>(f), L::packed({G, R, S, 1}));
  Ref tdx(dx, L::packed({N, H, W, G}));
  Ref none(nullptr, L::packed({1, 1, 1, G}));
  cutlass::reference::host::Depsep_Dgrad<float, L, float, L, float, L, float, L, float, float>(
      ps, tdy, tf, none, none, tdx, 1.f, 0.f, 0.f);
}

void ref_dwconv_bwd_filter(const float* dy, const float* x, float* df, int N, int G, int H, int W, int R, int S) {
  auto ps = problem(N, G, H, W, R, S);
  Ref tx(const_cast<float*>(x), L::packed({N, H, W, G}));
  Ref tdy(const_cast<float*>(dy), L::packed({N, H, W, G}));
  Ref tdf(df, L::packed({G, R, S, 1}));
  cutlass::reference::host::Depsep_Wgrad<float, L, float, L, float, L, float, float>(ps, tx, tdy, tdf, 1.f);
}

}  // extern "C"
