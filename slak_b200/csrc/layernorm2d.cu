// LayerNorm over the channels of an NCHW tensor ("channels_first" LayerNorm of models/SLaK.py:256-261: the stem
// and the three downsampling layers apply it between the stages), forward and backward, without leaving NCHW:
//     u = mean_c x,  s = mean_c (x-u)^2,  y = w[c] * (x-u)/sqrt(s+eps) + b[c]
// HBM-bound streaming kernels: one thread per pixel walks the C planes (consecutive threads = consecutive pixels,
// so every plane access of a warp is one contiguous segment); the second walk over the same planes hits L2.
// x / y / g may be fp32 or bf16 independently (bf16 stem-conv output -> fp32 residual stream; fp32 residual
// stream -> bf16 input of the stride-2 conv under autocast).  Statistics and the reductions are fp32.
// The parameter gradients are reduced deterministically: warp shuffle -> per-warp shared slots -> per-CTA partial
// rows -> fixed-order reduce kernel.
#include "common.cuh"

namespace slak {
namespace ln2d {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

template <typename TX, typename TY>
__global__ void __launch_bounds__(kThreads)
ln2d_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, TY* __restrict__ y,
                float* __restrict__ mean, float* __restrict__ rstd, long long total, int C, int HW, float eps) {
  extern __shared__ float wb[];                    // [2][C]
  for (int i = threadIdx.x; i < C; i += kThreads) { wb[i] = w[i]; wb[C + i] = b[i]; }
  __syncthreads();
  const float inv_c = 1.f / (float)C;
  for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads) {
    const long long n = idx / HW;
    const int p = (int)(idx - n * HW);
    const TX* xp = x + (size_t)n * C * HW + p;
    // one pass for both moments, shifted by the first channel's value (keeps the fp32 sums well conditioned)
    const float piv = to_f32<TX>(xp[0]);
    float s = 0.f, q = 0.f;
#pragma unroll 8
    for (int c = 0; c < C; ++c) {
      const float d = to_f32<TX>(xp[(size_t)c * HW]) - piv;
      s += d;
      q = fmaf(d, d, q);
    }
    const float m = s * inv_c;
    const float var = fmaxf(q * inv_c - m * m, 0.f);
    const float mu = piv + m;
    const float r = 1.f / sqrtf(var + eps);
    if (mean) { mean[idx] = mu; rstd[idx] = r; }
    TY* yp = y + (size_t)n * C * HW + p;
#pragma unroll 8
    for (int c = 0; c < C; ++c)
      yp[(size_t)c * HW] = from_f32<TY>(fmaf((to_f32<TX>(xp[(size_t)c * HW]) - mu) * r, wb[c], wb[C + c]));
  }
}

// dx = rstd * (g*w - mean_c(g*w) - xhat * mean_c(g*w*xhat));  dw[c] = sum g*xhat;  db[c] = sum g
template <typename TX, typename TG>
__global__ void __launch_bounds__(kThreads)
ln2d_bwd_kernel(const TG* __restrict__ g, const TX* __restrict__ x, const float* __restrict__ w,
                const float* __restrict__ mean, const float* __restrict__ rstd, TX* __restrict__ dx,
                float* __restrict__ part, long long total, int C, int HW) {
  extern __shared__ float sm[];                    // w [C] | acc [kWarps][2][C]
  float* ws = sm;
  float* acc = sm + C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < C; i += kThreads) ws[i] = w[i];
  for (int i = threadIdx.x; i < kWarps * 2 * C; i += kThreads) acc[i] = 0.f;
  __syncthreads();
  float* my = acc + warp * 2 * C;
  const float inv_c = 1.f / (float)C;
  for (long long base = (long long)blockIdx.x * kThreads + warp * 32; base < total; base += (long long)gridDim.x * kThreads) {
    const long long idx = base + lane;
    const bool ok = idx < total;
    const long long n = ok ? idx / HW : 0;
    const int p = ok ? (int)(idx - n * HW) : 0;
    const size_t off = (size_t)n * C * HW + p;
    const float mu = ok ? mean[idx] : 0.f, r = ok ? rstd[idx] : 0.f;
    float s1 = 0.f, s2 = 0.f;
    if (ok) {
#pragma unroll 8
      for (int c = 0; c < C; ++c) {
        const float gw = to_f32<TG>(g[off + (size_t)c * HW]) * ws[c];
        const float xh = (to_f32<TX>(x[off + (size_t)c * HW]) - mu) * r;
        s1 += gw;
        s2 = fmaf(gw, xh, s2);
      }
    }
    s1 *= inv_c; s2 *= inv_c;
    // second walk in chunks of 32 channels: 64 loads in flight per thread, then the per-channel sums over the
    // warp's 32 pixels by a transposing butterfly (31 shuffles per 32 channels instead of 5 per channel)
    for (int c0 = 0; c0 < C; c0 += 32) {
      float a[32], bs[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        float gv = 0.f, xv = 0.f;
        if (ok && c0 + k < C) {
          gv = to_f32<TG>(g[off + (size_t)(c0 + k) * HW]);
          xv = to_f32<TX>(x[off + (size_t)(c0 + k) * HW]);
        }
        a[k] = xv; bs[k] = gv;
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float xh = (a[k] - mu) * r;
        if (ok && c0 + k < C) dx[off + (size_t)(c0 + k) * HW] = from_f32<TX>(r * (bs[k] * ws[c0 + k] - s1 - xh * s2));
        a[k] = (ok && c0 + k < C) ? bs[k] * xh : 0.f;
      }
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < o; ++k) {
          const float sa = up ? a[k] : a[k + o], ka = up ? a[k + o] : a[k];
          const float sb = up ? bs[k] : bs[k + o], kb = up ? bs[k + o] : bs[k];
          a[k] = ka + __shfl_xor_sync(0xffffffffu, sa, o);
          bs[k] = kb + __shfl_xor_sync(0xffffffffu, sb, o);
        }
      }
      if (c0 + lane < C) { my[c0 + lane] += a[0]; my[C + c0 + lane] += bs[0]; }   // lane l ends up with channel c0 + l
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += kThreads) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kWarps; ++k) t += acc[k * 2 * C + i];
    part[(size_t)blockIdx.x * 2 * C + i] = t;
  }
}

__global__ void ln2d_reduce_kernel(const float* __restrict__ part, int parts, int C, float* __restrict__ dw, float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * C) return;
  float t = 0.f;
  for (int k = 0; k < parts; ++k) t += part[(size_t)k * 2 * C + i];
  if (i < C) dw[i] = t; else db[i - C] = t;
}

static int grid_for(long long total) {
  long long want = (total + kThreads - 1) / kThreads;
  const long long cap = 8LL * sm_count();          // 8 x 256 threads fill an SM
  if (want > cap) want = cap;
  return (int)(want < 1 ? 1 : want);
}

}  // namespace ln2d

int layernorm2d_bwd_parts(int N, int HW) { return ln2d::grid_for((long long)N * HW); }

int layernorm2d_fwd(const void* x, int xdt, const float* w, const float* b, float eps, void* y, int ydt, float* mean,
                    float* rstd, int N, int C, int HW, cudaStream_t st) {
  using namespace ln2d;
  const long long total = (long long)N * HW;
  if (total == 0) return SLAK_OK;
  const int grid = grid_for(total);
  const size_t smem = 2 * (size_t)C * sizeof(float);
  SLAK_REQUIRE(smem <= 48 * 1024, SLAK_ERR_UNSUPPORTED, "layernorm2d: C=%d too large", C);
#define LAUNCH(TX, TY)                                                                                         \
  ln2d_fwd_kernel<TX, TY><<<grid, kThreads, smem, st>>>((const TX*)x, w, b, (TY*)y, mean, rstd, total, C, HW, eps)
  if (xdt == SLAK_F32 && ydt == SLAK_F32) LAUNCH(float, float);
  else if (xdt == SLAK_F32 && ydt == SLAK_BF16) LAUNCH(float, __nv_bfloat16);
  else if (xdt == SLAK_BF16 && ydt == SLAK_F32) LAUNCH(__nv_bfloat16, float);
  else if (xdt == SLAK_BF16 && ydt == SLAK_BF16) LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else SLAK_REQUIRE(false, SLAK_ERR_UNSUPPORTED, "layernorm2d: dtypes must be fp32 or bf16");
#undef LAUNCH
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

int layernorm2d_bwd(const void* g, int gdt, const void* x, int xdt, const float* w, const float* mean, const float* rstd,
                    void* dx, float* part, float* dw, float* db, int N, int C, int HW, cudaStream_t st) {
  using namespace ln2d;
  const long long total = (long long)N * HW;
  if (total == 0) return SLAK_OK;
  const int grid = grid_for(total);
  const size_t smem = (size_t)(1 + 2 * kWarps) * C * sizeof(float);
  SLAK_REQUIRE(smem <= 200 * 1024, SLAK_ERR_UNSUPPORTED, "layernorm2d: C=%d too large", C);
#define LAUNCH(TX, TG)                                                                                         \
  do {                                                                                                         \
    SLAK_SET_MAX_SMEM((ln2d_bwd_kernel<TX, TG>), smem);                                                        \
    ln2d_bwd_kernel<TX, TG><<<grid, kThreads, smem, st>>>((const TG*)g, (const TX*)x, w, mean, rstd, (TX*)dx,  \
                                                          part, total, C, HW);                                 \
  } while (0)
  if (xdt == SLAK_F32 && gdt == SLAK_F32) LAUNCH(float, float);
  else if (xdt == SLAK_F32 && gdt == SLAK_BF16) LAUNCH(float, __nv_bfloat16);
  else if (xdt == SLAK_BF16 && gdt == SLAK_F32) LAUNCH(__nv_bfloat16, float);
  else if (xdt == SLAK_BF16 && gdt == SLAK_BF16) LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else SLAK_REQUIRE(false, SLAK_ERR_UNSUPPORTED, "layernorm2d: dtypes must be fp32 or bf16");
#undef LAUNCH
  SLAK_CUDA_TRY(cudaGetLastError());
  ln2d_reduce_kernel<<<(2 * C + 127) / 128, 128, 0, st>>>(part, grid, C, dw, db);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

}  // namespace slak
