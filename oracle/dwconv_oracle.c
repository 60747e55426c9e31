/*
 * oracle/dwconv_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the reference's depthwise convolution semantics, used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker.
 *
 * Follows the loop nests of the reference's own host reference
 *   cutlass/tools/util/include/cutlass/util/reference/host/convolution.h
 *     Depsep_Fprop :237-322, Depsep_Dgrad :332-421, Depsep_Wgrad :431-490
 * specialised to what the torch extension hard-wires (forward_fp32.cu:135-144,227,241):
 * stride 1, dilation 1, pad = (R/2, S/2), cross-correlation, alpha=1, beta=gamma=0,
 * and to the extension's NCHW tensors (frontend.h:3-10).  Accumulation is in double so
 * the oracle is at least as accurate as any fp32-accumulating implementation.
 * Pinned against the reference code itself by oracle/_ref (see oracle/Makefile) and
 * against torch.nn.functional.conv2d (test_correctness.py:8-9) in tests/test_oracle.py.
 */
#include <stddef.h>

#define IDX4(n, c, h, w, C, H, W) ((((size_t)(n) * (C) + (c)) * (H) + (h)) * (W) + (w))

/* y[n,g,p,q] = sum_{r,s} x[n,g,p-PH+r,q-PW+s] * f[g,r,s]   (Depsep_Fprop :268-309) */
void oracle_dwconv_fwd(const float* x, const float* f, float* y, int N, int G, int H, int W, int R,
                       int S) {
  const int PH = R / 2, PW = S / 2;
  for (int n = 0; n < N; ++n)
    for (int g = 0; g < G; ++g)
      for (int p = 0; p < H; ++p)
        for (int q = 0; q < W; ++q) {
          double acc = 0.0;
          for (int r = 0; r < R; ++r)
            for (int s = 0; s < S; ++s) {
              const int h = p - PH + r, w = q - PW + s;
              if (h >= 0 && h < H && w >= 0 && w < W)
                acc += (double)x[IDX4(n, g, h, w, G, H, W)] * (double)f[((size_t)g * R + r) * S + s];
            }
          y[IDX4(n, g, p, q, G, H, W)] = (float)acc;
        }
}

/* dx[n,g,h,w] = sum_{r,s} dy[n,g,h+PH-r,w+PW-s] * f[g,r,s]   (Depsep_Dgrad :363-398) */
void oracle_dwconv_bwd_data(const float* dy, const float* f, float* dx, int N, int G, int H, int W,
                            int R, int S) {
  const int PH = R / 2, PW = S / 2;
  for (int n = 0; n < N; ++n)
    for (int g = 0; g < G; ++g)
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
          double acc = 0.0;
          for (int r = 0; r < R; ++r)
            for (int s = 0; s < S; ++s) {
              const int p = h + PH - r, q = w + PW - s;
              if (p >= 0 && p < H && q >= 0 && q < W)
                acc += (double)dy[IDX4(n, g, p, q, G, H, W)] * (double)f[((size_t)g * R + r) * S + s];
            }
          dx[IDX4(n, g, h, w, G, H, W)] = (float)acc;
        }
}

/* df[g,r,s] = sum_{n,p,q} x[n,g,p-PH+r,q-PW+s] * dy[n,g,p,q]   (Depsep_Wgrad :459-487) */
void oracle_dwconv_bwd_filter(const float* dy, const float* x, float* df, int N, int G, int H, int W,
                              int R, int S) {
  const int PH = R / 2, PW = S / 2;
  for (int g = 0; g < G; ++g)
    for (int r = 0; r < R; ++r)
      for (int s = 0; s < S; ++s) {
        double acc = 0.0;
        for (int n = 0; n < N; ++n)
          for (int p = 0; p < H; ++p) {
            const int h = p - PH + r;
            if (h < 0 || h >= H) continue;
            for (int q = 0; q < W; ++q) {
              const int w = q - PW + s;
              if (w >= 0 && w < W)
                acc += (double)x[IDX4(n, g, h, w, G, H, W)] * (double)dy[IDX4(n, g, p, q, G, H, W)];
            }
          }
        df[((size_t)g * R + r) * S + s] = (float)acc;
      }
}
