// Memory-bound glue of a SLaK Block around the depthwise branches, fused into a few passes
// (models/SLaK.py:89-100 BN+sum, :153-166 permute / LayerNorm / gamma / residual):
//
//   bn3_finalize_fwd   per-channel batch statistics of the three branch outputs -> BN scale/shift,
//                      running-stat update                                  (nn.BatchNorm2d semantics)
//   bn3_sum_ln_fwd     xn[n,h,w,:] = LayerNorm_C( sum_i scale_i*y_i + shift )   NCHW bf16 x3 -> NHWC bf16
//   residual_fwd       out = x + dp[n]*gamma[c]*h2[n,h,w,c]                       NHWC bf16 -> NCHW fp32
//   residual_bwd       d_h2 = dOut*gamma*dp (NHWC bf16), per-CTA partial of dgamma
//   bn3_sum_ln_bwd     LayerNorm backward per pixel -> du (NCHW bf16) + per-CTA partials of dln_w,
//                      dln_b and of the BatchNorm reductions sum(du), sum(du*y_i)
//   bn3_finalize_bwd   BatchNorm backward coefficients: dy_i = A_i*du + B_i*y_i + C_i, dgamma_i, dbeta_i
//   bn3_bwd_apply      dy_i (NCHW bf16) for the three branches in one pass
//
// All kernels work on a (C x PIX) tile of one image staged in shared memory as fp32, so that the
// NCHW side is accessed along pixels and the NHWC side along channels (both coalesced); LayerNorm
// runs one warp per pixel with lanes over channels.  Reductions are per-CTA partials combined in a
// fixed order (deterministic).
#include "common.cuh"
#include "block_glue2.cuh"
#include <stdlib.h>

namespace slak {
namespace blk {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float bf(const __nv_bfloat16 v) { return __bfloat162float(v); }

// load VP consecutive bf16 -> float
template <int VP>
__device__ __forceinline__ void ld_bf16(const __nv_bfloat16* p, float* f) {
  if constexpr (VP == 8) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 t = __bfloat1622float2(h[k]); f[2 * k] = t.x; f[2 * k + 1] = t.y; }
  } else if constexpr (VP == 4) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int k = 0; k < 2; ++k) { const float2 t = __bfloat1622float2(h[k]); f[2 * k] = t.x; f[2 * k + 1] = t.y; }
  } else {
    f[0] = __bfloat162float(*p);
  }
}
template <int VP>
__device__ __forceinline__ void st_bf16(__nv_bfloat16* p, const float* f) {
  if constexpr (VP == 8) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  } else if constexpr (VP == 4) {
    uint2 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int k = 0; k < 2; ++k) h[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
    *reinterpret_cast<uint2*>(p) = r;
  } else {
    *p = __float2bfloat16_rn(f[0]);
  }
}

struct Geo {
  int N, C, HW, PIX, tiles_per_img, total_tiles, pitch;
};

// ------------------------------------------------------------------------------------------
// forward: BN(3) + sum + LayerNorm, NCHW -> NHWC
// ------------------------------------------------------------------------------------------
template <int VP>
__global__ void __launch_bounds__(kThreads)
bn3_sum_ln_fwd_kernel(const __nv_bfloat16* __restrict__ y1, const __nv_bfloat16* __restrict__ y2,
                      const __nv_bfloat16* __restrict__ y3, const float* __restrict__ scale /*[3][C]*/,
                      const float* __restrict__ shift /*[C]*/, const float* __restrict__ lnw,
                      const float* __restrict__ lnb, float eps, __nv_bfloat16* __restrict__ xn,
                      float* __restrict__ mu, float* __restrict__ rstd, Geo g) {
  extern __shared__ float tile[];   // [C][pitch] | lnw [C] | lnb [C] | ps, pq [kThreads] | mean, rstd [PIX]
  const int tid = threadIdx.x;
  const int C = g.C, HW = g.HW, PIX = g.PIX, pitch = g.pitch;
  const int vec_per_row = PIX / VP;
  float* lnw_s = tile + (size_t)C * pitch;
  float* lnb_s = lnw_s + C;
  float* ps = lnb_s + C;
  float* pq = ps + kThreads;
  float* mean_s = pq + kThreads;
  float* rstd_s = mean_s + PIX;
  for (int i = tid; i < C; i += kThreads) { lnw_s[i] = lnw[i]; lnb_s[i] = lnb[i]; }
  const int nparts = kThreads / PIX;          // threads per pixel in the statistics pass (PIX <= 128)
  const int pj = tid % PIX, part = tid / PIX;
  for (int t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
    const int n = t / g.tiles_per_img, p0 = (t - n * g.tiles_per_img) * PIX;
    const int npix = min(PIX, HW - p0);
    __syncthreads();
    // phase 1: u = sum_i scale_i*y_i + shift, NCHW rows -> tile
    for (int idx = tid; idx < C * vec_per_row; idx += kThreads) {
      const int c = idx / vec_per_row, jv = idx - c * vec_per_row;
      const int j = jv * VP;
      if (j < npix) {
        const size_t off = ((size_t)n * C + c) * HW + p0 + j;
        float a[VP], b[VP], d[VP];
        ld_bf16<VP>(y1 + off, a); ld_bf16<VP>(y2 + off, b); ld_bf16<VP>(y3 + off, d);
        const float s1 = scale[c], s2 = scale[C + c], s3 = scale[2 * C + c], sh = shift[c];
#pragma unroll
        for (int k = 0; k < VP; ++k) tile[c * pitch + j + k] = fmaf(s1, a[k], fmaf(s2, b[k], fmaf(s3, d[k], sh)));
      }
    }
    __syncthreads();
    // phase 2a: moments per pixel, nparts threads per pixel, consecutive threads = consecutive pixels (conflict-free
    // column walks); both moments in one pass, shifted by the pixel's first channel
    if (part < nparts && pj < npix) {
      const float piv = tile[pj];
      float s = 0.f, q = 0.f;
      for (int c = part; c < C; c += nparts) { const float d = tile[c * pitch + pj] - piv; s += d; q = fmaf(d, d, q); }
      ps[part * PIX + pj] = s; pq[part * PIX + pj] = q;
    }
    __syncthreads();
    if (tid < npix) {
      float s = 0.f, q = 0.f;
      for (int k = 0; k < nparts; ++k) { s += ps[k * PIX + tid]; q += pq[k * PIX + tid]; }
      const float m = s / C;
      const float var = fmaxf(q / C - m * m, 0.f);
      const float mean = tile[tid] + m, r = rsqrtf(var + eps);
      mean_s[tid] = mean; rstd_s[tid] = r;
      const size_t pix = (size_t)n * HW + p0 + tid;
      mu[pix] = mean; rstd[pix] = r;
    }
    __syncthreads();
    // phase 2b: xn rows (NHWC bf16), consecutive threads = consecutive channel pairs of a pixel
    __nv_bfloat16* orow = xn + ((size_t)n * HW + p0) * C;
    if ((C & 1) == 0) {
      const int half = C / 2;
      for (int idx = tid; idx < npix * half; idx += kThreads) {
        const int j = idx / half, c = 2 * (idx - j * half);
        const float mean = mean_s[j], r = rstd_s[j];
        const float v0 = (tile[c * pitch + j] - mean) * r * lnw_s[c] + lnb_s[c];
        const float v1 = (tile[(c + 1) * pitch + j] - mean) * r * lnw_s[c + 1] + lnb_s[c + 1];
        *reinterpret_cast<__nv_bfloat162*>(orow + (size_t)j * C + c) = __floats2bfloat162_rn(v0, v1);
      }
    } else {
      for (int idx = tid; idx < npix * C; idx += kThreads) {
        const int j = idx / C, c = idx - j * C;
        orow[idx] = __float2bfloat16_rn((tile[c * pitch + j] - mean_s[j]) * rstd_s[j] * lnw_s[c] + lnb_s[c]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// forward: out = x + dp*gamma*h2^T
// ------------------------------------------------------------------------------------------
template <int VP>
__global__ void __launch_bounds__(kThreads)
residual_fwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ h2,
                    const float* __restrict__ gamma, const float* __restrict__ dp /*[N] or null*/,
                    float* __restrict__ out, __nv_bfloat16* __restrict__ out_bf16 /*or null*/, Geo g) {
  extern __shared__ float tile[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = g.C, HW = g.HW, PIX = g.PIX, pitch = g.pitch;
  const int vec_per_row = PIX / VP;
  for (int t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
    const int n = t / g.tiles_per_img, p0 = (t - n * g.tiles_per_img) * PIX;
    const int npix = min(PIX, HW - p0);
    const float dps = dp ? dp[n] : 1.f;
    __syncthreads();
    for (int j = warp; j < npix; j += kWarps) {
      const __nv_bfloat16* hp = h2 + ((size_t)n * HW + p0 + j) * C;
      if ((C & 1) == 0) {
        for (int c = 2 * lane; c < C; c += 64) {
          const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(hp + c));
          tile[c * pitch + j] = v.x * gamma[c] * dps;
          tile[(c + 1) * pitch + j] = v.y * gamma[c + 1] * dps;
        }
      } else {
        for (int c = lane; c < C; c += 32) tile[c * pitch + j] = bf(hp[c]) * gamma[c] * dps;
      }
    }
    __syncthreads();
    for (int idx = tid; idx < C * vec_per_row; idx += kThreads) {
      const int c = idx / vec_per_row, jv = idx - c * vec_per_row;
      const int j = jv * VP;
      if (j < npix) {
        const size_t off = ((size_t)n * C + c) * HW + p0 + j;
        float o[VP];
        if constexpr (VP >= 4) {
#pragma unroll
          for (int k = 0; k < VP; k += 4) {
            const float4 xv = *reinterpret_cast<const float4*>(x + off + k);
            o[k] = xv.x + tile[c * pitch + j + k]; o[k + 1] = xv.y + tile[c * pitch + j + k + 1];
            o[k + 2] = xv.z + tile[c * pitch + j + k + 2]; o[k + 3] = xv.w + tile[c * pitch + j + k + 3];
          }
        } else {
          o[0] = x[off] + tile[c * pitch + j];
        }
        if constexpr (VP >= 4) {
#pragma unroll
          for (int k = 0; k < VP; k += 4) *reinterpret_cast<float4*>(out + off + k) = make_float4(o[k], o[k + 1], o[k + 2], o[k + 3]);
        } else {
          out[off] = o[0];
        }
        if (out_bf16) st_bf16<VP>(out_bf16 + off, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward: d_h2 = dOut*gamma*dp (NHWC bf16); partial dgamma[cta][C] = sum dOut*h2*dp
// ------------------------------------------------------------------------------------------
template <int VP>
__global__ void __launch_bounds__(kThreads)
residual_bwd_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ h2,
                    const float* __restrict__ gamma, const float* __restrict__ dp,
                    __nv_bfloat16* __restrict__ dh2, float* __restrict__ dgamma_part /*[grid][2][C]: dgamma, db2*/, Geo g) {
  extern __shared__ float tile[];
  float* acc = tile + (size_t)g.C * g.pitch;   // [kWarps][C] dgamma
  float* acc2 = acc + kWarps * g.C;            // [kWarps][C] column sums of dh2
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = g.C, HW = g.HW, PIX = g.PIX, pitch = g.pitch;
  const int vec_per_row = PIX / VP;
  for (int i = tid; i < 2 * kWarps * C; i += kThreads) acc[i] = 0.f;
  for (int t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
    const int n = t / g.tiles_per_img, p0 = (t - n * g.tiles_per_img) * PIX;
    const int npix = min(PIX, HW - p0);
    const float dps = dp ? dp[n] : 1.f;
    __syncthreads();
    for (int idx = tid; idx < C * vec_per_row; idx += kThreads) {
      const int c = idx / vec_per_row, jv = idx - c * vec_per_row;
      const int j = jv * VP;
      if (j < npix) {
        const size_t off = ((size_t)n * C + c) * HW + p0 + j;
        if constexpr (VP >= 4) {
#pragma unroll
          for (int k = 0; k < VP; k += 4) {
            const float4 dv = *reinterpret_cast<const float4*>(dout + off + k);
            tile[c * pitch + j + k] = dv.x; tile[c * pitch + j + k + 1] = dv.y;
            tile[c * pitch + j + k + 2] = dv.z; tile[c * pitch + j + k + 3] = dv.w;
          }
        } else {
          tile[c * pitch + j] = dout[off];
        }
      }
    }
    __syncthreads();
    for (int j = warp; j < npix; j += kWarps) {
      const size_t pix = (size_t)n * HW + p0 + j;
      const __nv_bfloat16* hp = h2 + pix * C;
      __nv_bfloat16* dp_out = dh2 + pix * C;
      if ((C & 1) == 0) {
        for (int c = 2 * lane; c < C; c += 64) {
          const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(hp + c));
          const float g0 = tile[c * pitch + j] * dps, g1 = tile[(c + 1) * pitch + j] * dps;
          const __nv_bfloat162 d2 = __floats2bfloat162_rn(g0 * gamma[c], g1 * gamma[c + 1]);
          *reinterpret_cast<__nv_bfloat162*>(dp_out + c) = d2;
          const float2 d2f = __bfloat1622float2(d2);
          acc[warp * C + c] += g0 * h.x;
          acc[warp * C + c + 1] += g1 * h.y;
          acc2[warp * C + c] += d2f.x;
          acc2[warp * C + c + 1] += d2f.y;
        }
      } else {
        for (int c = lane; c < C; c += 32) {
          const float g0 = tile[c * pitch + j] * dps;
          const __nv_bfloat16 d1 = __float2bfloat16_rn(g0 * gamma[c]);
          dp_out[c] = d1;
          acc[warp * C + c] += g0 * bf(hp[c]);
          acc2[warp * C + c] += bf(d1);
        }
      }
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += kThreads) {
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { s += acc[w * C + c]; s2 += acc2[w * C + c]; }
    dgamma_part[(size_t)blockIdx.x * 2 * C + c] = s;
    dgamma_part[(size_t)blockIdx.x * 2 * C + C + c] = s2;
  }
}

// ------------------------------------------------------------------------------------------
// MLP backward glue: dh = da * gelu'(h) (exact erf GELU), plus per-CTA column sums of dh
// (bias gradient of pwconv1).  da, h, dh: [rows][K] bf16 row-major, K % 8 == 0.
// ------------------------------------------------------------------------------------------
// d/dx [x * Phi(x)] = Phi(x) + x * phi(x).  Phi through erf's rational approximation (Abramowitz & Stegun 7.1.26,
// |error| <= 1.5e-7, far below the bf16 rounding of the result): it shares the one exponential with phi, which
// keeps this kernel on the memory side of its roofline (erff() alone costs more than the two loads and the store)
__device__ __forceinline__ float gelu_grad(float x) {
  const float e = __expf(-0.5f * x * x);                       // exp(-z^2), z = |x| / sqrt(2)
  const float t = __fdividef(1.f, fmaf(0.3275911f * 0.70710678118654752f, fabsf(x), 1.f));
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  const float erfz = fmaf(-poly, e, 1.f);                      // erf(|x| / sqrt(2))
  const float cdf = 0.5f + copysignf(0.5f * erfz, x);
  return fmaf(x, 0.39894228040143268f * e, cdf);
}
__global__ void __launch_bounds__(kThreads)
gelu_bwd_bias_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ h,
                     __nv_bfloat16* __restrict__ dh, float* __restrict__ part /*[grid][K]*/, long long rows, int K,
                     int rows_per_cta) {
  extern __shared__ float red[];               // [rsub][K]
  const int nv = K / 8;
  const int tid = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  const int rsubs = nv >= kThreads ? 1 : kThreads / nv;
  for (int v0 = 0; v0 < nv; v0 += kThreads) {
    const int vec = v0 + (nv >= kThreads ? tid : tid % nv);
    const int rsub = nv >= kThreads ? 0 : tid / nv;
    float s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = 0.f;
    if (vec < nv && rsub < rsubs) {
      constexpr int U = 4;                       // rows in flight per thread: 8 x 16-byte loads before the first use
      for (long long r = r0 + rsub; r < r1; r += (long long)U * rsubs) {
        uint4 ra[U], rx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long rr = r + (long long)u * rsubs;
          if (rr < r1) {
            const size_t off = (size_t)rr * K + (size_t)vec * 8;
            ra[u] = *reinterpret_cast<const uint4*>(da + off);
            rx[u] = *reinterpret_cast<const uint4*>(h + off);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long rr = r + (long long)u * rsubs;
          if (rr < r1) {
            const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&ra[u]);
            const __nv_bfloat162* px = reinterpret_cast<const __nv_bfloat162*>(&rx[u]);
            float o[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 a2 = __bfloat1622float2(pa[k]), x2 = __bfloat1622float2(px[k]);
              o[2 * k] = a2.x * gelu_grad(x2.x);
              o[2 * k + 1] = a2.y * gelu_grad(x2.y);
            }
            st_bf16<8>(dh + (size_t)rr * K + (size_t)vec * 8, o);
#pragma unroll
            for (int k = 0; k < 8; ++k) s[k] += __bfloat162float(__float2bfloat16_rn(o[k]));
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) red[(size_t)rsub * K + vec * 8 + k] = s[k];
    }
  }
  __syncthreads();
  for (int c = tid; c < K; c += kThreads) {
    float t = 0.f;
    for (int rs = 0; rs < rsubs; ++rs) t += red[(size_t)rs * K + c];
    part[(size_t)blockIdx.x * K + c] = t;
  }
}

// ------------------------------------------------------------------------------------------
// backward: LayerNorm backward + BatchNorm reductions.  part layout per CTA: [6][C] =
//   dlnw, dlnb, S0 = sum du, S1..S3 = sum du*y_i
// Per tile: (1) stage u = sum_i scale_i*y_i + shift (fp32 [C][pitch]), the incoming gradient rows (bf16
// [PIX][ge], ge/2 odd so that a column walk is bank-conflict free), mu and rstd -- every global load of the tile is
// in flight at once; (2) one warp per pixel reduces mean_c(g*w) and mean_c(g*w*xhat) from shared memory; (3) one
// thread per (channel, VP pixels) forms du, stores it (NCHW) and accumulates all six per-channel sums, which a
// segmented shuffle folds into shared accumulators owned by exactly one lane each (deterministic).
// ------------------------------------------------------------------------------------------
__host__ __device__ inline int ln_bwd_gpitch(int C) {   // bf16 elements per staged gradient row
  int ge = (C + 1) & ~1;
  if (((ge / 2) & 1) == 0) ge += 2;
  return ge;
}
template <int VP>
__global__ void __launch_bounds__(kThreads)
bn3_sum_ln_bwd_kernel(const __nv_bfloat16* __restrict__ dxn, const __nv_bfloat16* __restrict__ y1,
                      const __nv_bfloat16* __restrict__ y2, const __nv_bfloat16* __restrict__ y3,
                      const float* __restrict__ scale, const float* __restrict__ shift,
                      const float* __restrict__ lnw, const float* __restrict__ mu, const float* __restrict__ rstd,
                      __nv_bfloat16* __restrict__ du, float* __restrict__ part /*[grid][6][C]*/, Geo g) {
  extern __shared__ float tile[];
  const int C = g.C, HW = g.HW, PIX = g.PIX, pitch = g.pitch;
  const int ge = ln_bwd_gpitch(C);
  float* accs = tile + (size_t)C * pitch;      // [6][C]
  float* lnw_s = accs + 6 * C;                 // [C]
  float* mus = lnw_s + C;                      // [PIX] mu, rstd, m1, m2
  float* rs = mus + PIX;
  float* m1s = rs + PIX;
  float* m2s = m1s + PIX;
  float* ps1 = m2s + PIX;                      // [kThreads] partial sums of phase 2
  float* ps2 = ps1 + kThreads;
  __nv_bfloat16* gs = reinterpret_cast<__nv_bfloat16*>(ps2 + kThreads);   // [PIX][ge]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int vec_per_row = PIX / VP;
  for (int i = tid; i < 6 * C; i += kThreads) accs[i] = 0.f;
  for (int i = tid; i < C; i += kThreads) lnw_s[i] = lnw[i];
  for (int t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
    const int n = t / g.tiles_per_img, p0 = (t - n * g.tiles_per_img) * PIX;
    const int npix = min(PIX, HW - p0);
    __syncthreads();
    // phase 1a: gradient rows of the tile: one contiguous run of npix*C bf16 in NHWC
    {
      const __nv_bfloat16* gsrc = dxn + ((size_t)n * HW + p0) * C;
      if ((C & 7) == 0 && (reinterpret_cast<uintptr_t>(dxn) & 15) == 0) {
        const int nvec = npix * C / 8;
        for (int v = tid; v < nvec; v += kThreads) {
          const uint4 r = *reinterpret_cast<const uint4*>(gsrc + (size_t)v * 8);
          const int e = v * 8, j = e / C, c = e - j * C;
          uint32_t* d32 = reinterpret_cast<uint32_t*>(gs + (size_t)j * ge + c);
          d32[0] = r.x; d32[1] = r.y; d32[2] = r.z; d32[3] = r.w;
        }
      } else {
        for (int e = tid; e < npix * C; e += kThreads) {
          const int j = e / C, c = e - j * C;
          gs[(size_t)j * ge + c] = gsrc[e];
        }
      }
      for (int j = tid; j < npix; j += kThreads) { mus[j] = mu[(size_t)n * HW + p0 + j]; rs[j] = rstd[(size_t)n * HW + p0 + j]; }
    }
    // phase 1b: recompute u = sum_i scale_i*y_i + shift
    for (int idx = tid; idx < C * vec_per_row; idx += kThreads) {
      const int c = idx / vec_per_row, jv = idx - c * vec_per_row;
      const int j = jv * VP;
      if (j < npix) {
        const size_t off = ((size_t)n * C + c) * HW + p0 + j;
        float a[VP], b[VP], d[VP];
        ld_bf16<VP>(y1 + off, a); ld_bf16<VP>(y2 + off, b); ld_bf16<VP>(y3 + off, d);
        const float s1 = scale[c], s2 = scale[C + c], s3 = scale[2 * C + c], sh = shift[c];
#pragma unroll
        for (int k = 0; k < VP; ++k) tile[c * pitch + j + k] = fmaf(s1, a[k], fmaf(s2, b[k], fmaf(s3, d[k], sh)));
      }
    }
    __syncthreads();
    // phase 2: per pixel m1 = mean_c(g*w), m2 = mean_c(g*w*xhat); nparts threads per pixel, consecutive threads =
    // consecutive pixels (conflict-free walks of the u tile columns and of the odd-pitch gradient rows)
    {
      const int nparts = kThreads / PIX;
      const int pj = tid % PIX, prt = tid / PIX;
      if (prt < nparts && pj < npix) {
        const float m = mus[pj], r = rs[pj];
        float s1 = 0.f, s2 = 0.f;
        for (int c = prt; c < C; c += nparts) {
          const float gg = bf(gs[(size_t)pj * ge + c]) * lnw_s[c];
          s1 += gg;
          s2 = fmaf(gg, (tile[c * pitch + pj] - m) * r, s2);
        }
        ps1[prt * PIX + pj] = s1; ps2[prt * PIX + pj] = s2;
      }
      __syncthreads();
      if (tid < npix) {
        float s1 = 0.f, s2 = 0.f;
        for (int k = 0; k < nparts; ++k) { s1 += ps1[k * PIX + tid]; s2 += ps2[k * PIX + tid]; }
        m1s[tid] = s1 / C; m2s[tid] = s2 / C;
      }
    }
    __syncthreads();
    // phase 3: du (NCHW bf16) and the six per-channel sums
    auto body = [&](int c, int j, float* a) {
      const size_t off = ((size_t)n * C + c) * HW + p0 + j;
      float d[VP], q1[VP], q2[VP], q3[VP];
      ld_bf16<VP>(y1 + off, q1); ld_bf16<VP>(y2 + off, q2); ld_bf16<VP>(y3 + off, q3);
      const float w = lnw_s[c];
#pragma unroll
      for (int k = 0; k < VP; ++k) {
        const float r = rs[j + k];
        const float xh = (tile[c * pitch + j + k] - mus[j + k]) * r;
        const float gv = bf(gs[(size_t)(j + k) * ge + c]);
        d[k] = r * (gv * w - m1s[j + k] - xh * m2s[j + k]);
        a[0] = fmaf(gv, xh, a[0]); a[1] += gv;
      }
      st_bf16<VP>(du + off, d);
#pragma unroll
      for (int k = 0; k < VP; ++k) {
        const float dr = __bfloat162float(__float2bfloat16_rn(d[k]));   // the value BN backward will see
        a[2] += dr; a[3] = fmaf(dr, q1[k], a[3]); a[4] = fmaf(dr, q2[k], a[4]); a[5] = fmaf(dr, q3[k], a[5]);
      }
    };
    if (vec_per_row <= 32) {
      // vec_per_row (a power of two) consecutive lanes share a channel: segmented shuffle reduction
      const int total = C * vec_per_row;
      for (int idx0 = 0; idx0 < total; idx0 += kThreads) {
        const int idx = idx0 + tid;
        const int c = idx / vec_per_row, jv = idx - c * vec_per_row;
        const int j = jv * VP;
        float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (idx < total && j < npix) body(c, j, a);
        for (int o = vec_per_row >> 1; o > 0; o >>= 1) {
#pragma unroll
          for (int q = 0; q < 6; ++q) a[q] += __shfl_xor_sync(0xffffffffu, a[q], o);
        }
        if (idx < total && jv == 0) {
#pragma unroll
          for (int q = 0; q < 6; ++q) accs[q * C + c] += a[q];
        }
      }
    } else {
      for (int c = warp; c < C; c += kWarps) {
        float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int jv = lane; jv < vec_per_row; jv += 32) {
          const int j = jv * VP;
          if (j < npix) body(c, j, a);
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) a[q] = warp_sum(a[q]);
        if (lane == 0) {
#pragma unroll
          for (int q = 0; q < 6; ++q) accs[q * C + c] += a[q];
        }
      }
    }
  }
  __syncthreads();
  float* o = part + (size_t)blockIdx.x * 6 * C;
  for (int i = tid; i < 6 * C; i += kThreads) o[i] = accs[i];
}

// dy_i = A_i*du + B_i*y_i + C_i   (coef: [9][C] = A1..A3, B1..B3, C1..C3)
template <int VP>
__global__ void __launch_bounds__(kThreads)
bn3_bwd_apply_kernel(const __nv_bfloat16* __restrict__ du, const __nv_bfloat16* __restrict__ y1,
                     const __nv_bfloat16* __restrict__ y2, const __nv_bfloat16* __restrict__ y3,
                     const float* __restrict__ coef, __nv_bfloat16* __restrict__ dy1,
                     __nv_bfloat16* __restrict__ dy2, __nv_bfloat16* __restrict__ dy3, int C, int HW, size_t planes) {
  const int vec_per_plane = HW / VP;
  const size_t total = planes * vec_per_plane;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
    const size_t plane = i / vec_per_plane;
    const int c = (int)(plane % C);
    const size_t off = i * VP;
    float d[VP], a[VP], b[VP], e[VP], o[VP];
    ld_bf16<VP>(du + off, d); ld_bf16<VP>(y1 + off, a); ld_bf16<VP>(y2 + off, b); ld_bf16<VP>(y3 + off, e);
    const float A1 = coef[c], A2 = coef[C + c], A3 = coef[2 * C + c];
    const float B1 = coef[3 * C + c], B2 = coef[4 * C + c], B3 = coef[5 * C + c];
    const float C1 = coef[6 * C + c], C2 = coef[7 * C + c], C3 = coef[8 * C + c];
#pragma unroll
    for (int k = 0; k < VP; ++k) o[k] = fmaf(A1, d[k], fmaf(B1, a[k], C1));
    st_bf16<VP>(dy1 + off, o);
#pragma unroll
    for (int k = 0; k < VP; ++k) o[k] = fmaf(A2, d[k], fmaf(B2, b[k], C2));
    st_bf16<VP>(dy2 + off, o);
#pragma unroll
    for (int k = 0; k < VP; ++k) o[k] = fmaf(A3, d[k], fmaf(B3, e[k], C3));
    st_bf16<VP>(dy3 + off, o);
  }
}

// The same for planes whose size is not a multiple of 4 (7 x 7): the whole tensor is contiguous, so it is walked with 16-byte
// vectors of 8 elements regardless of the plane boundaries; a vector touches at most two channels (HW >= 8).
__global__ void __launch_bounds__(kThreads)
bn3_bwd_apply_flat_kernel(const __nv_bfloat16* __restrict__ du, const __nv_bfloat16* __restrict__ y1,
                          const __nv_bfloat16* __restrict__ y2, const __nv_bfloat16* __restrict__ y3,
                          const float* __restrict__ coef, __nv_bfloat16* __restrict__ dy1, __nv_bfloat16* __restrict__ dy2,
                          __nv_bfloat16* __restrict__ dy3, int C, int HW, size_t total /* elements, multiple of 8 */) {
  const size_t nvec = total / 8;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += (size_t)gridDim.x * kThreads) {
    const size_t e0 = i * 8;
    const size_t plane = e0 / HW;
    const int left = (int)((plane + 1) * HW - e0);            // elements of this vector inside the first plane (>= 1)
    const int c0 = (int)(plane % C), c1 = c0 + 1 == C ? 0 : c0 + 1;
    float d[8], a[8], b[8], e[8], o[8];
    ld_bf16<8>(du + e0, d); ld_bf16<8>(y1 + e0, a); ld_bf16<8>(y2 + e0, b); ld_bf16<8>(y3 + e0, e);
    float k0[9], k1[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) { k0[q] = coef[q * C + c0]; k1[q] = coef[q * C + c1]; }
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = k < left ? fmaf(k0[0], d[k], fmaf(k0[3], a[k], k0[6])) : fmaf(k1[0], d[k], fmaf(k1[3], a[k], k1[6]));
    st_bf16<8>(dy1 + e0, o);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = k < left ? fmaf(k0[1], d[k], fmaf(k0[4], b[k], k0[7])) : fmaf(k1[1], d[k], fmaf(k1[4], b[k], k1[7]));
    st_bf16<8>(dy2 + e0, o);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = k < left ? fmaf(k0[2], d[k], fmaf(k0[5], e[k], k0[8])) : fmaf(k1[2], d[k], fmaf(k1[5], e[k], k1[8]));
    st_bf16<8>(dy3 + e0, o);
  }
}

// per-channel batch statistics from the conv kernel's per-CTA partials -> BN scale/shift (training)
//   part: [C][splits][6] = (sum, sumsq) x 3 branches ; stats_out: [C][6] (sum, sumsq totals, for SyncBN)
__global__ void bn3_reduce_partials_kernel(const float* __restrict__ part, int splits, int C, double* __restrict__ sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over C*6
  if (i >= C * 6) return;
  const int c = i / 6, k = i - c * 6;
  double s = 0.0;
  for (int sp = 0; sp < splits; ++sp) s += (double)part[((size_t)c * splits + sp) * 6 + k];
  sums[i] = s;
}
// sums: [C][6] doubles (global over all ranks), count = N*H*W (global)
// bnw/bnb: [3][C]; running_mean/var: [3][C] (updated in place when momentum >= 0)
// out: scale[3][C], shift[C], mean[3][C], istd[3][C]
struct P3 { const float* p[3]; };
struct M3 { float* p[3]; };
__device__ __forceinline__ void bn3_finalize_fwd_channel(const double* s6, double count, int c, const P3& bnw, const P3& bnb,
                                                         const M3& rmean, const M3& rvar, float eps, float momentum, int C,
                                                         float* __restrict__ scale, float* __restrict__ shift,
                                                         float* __restrict__ mean, float* __restrict__ istd) {
  float sh = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double m = s6[2 * i] / count;
    double var = s6[2 * i + 1] / count - m * m;
    if (var < 0.0) var = 0.0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = bnw.p[i][c] * is;
    scale[i * C + c] = sc;
    mean[i * C + c] = (float)m;
    istd[i * C + c] = is;
    sh += bnb.p[i][c] - (float)m * sc;
    if (rmean.p[i]) {
      const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
      rmean.p[i][c] = (1.f - momentum) * rmean.p[i][c] + momentum * (float)m;
      rvar.p[i][c] = (1.f - momentum) * rvar.p[i][c] + momentum * (float)unb;
    }
  }
  shift[c] = sh;
}
__global__ void bn3_finalize_fwd_kernel(const double* __restrict__ sums, double count, const double* __restrict__ count_dev,
                                        P3 bnw, P3 bnb, M3 rmean, M3 rvar,
                                        float eps, float momentum, int C, float* __restrict__ scale,
                                        float* __restrict__ shift, float* __restrict__ mean, float* __restrict__ istd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (count_dev) count = *count_dev;            // SyncBN: the all-reduced number of elements per channel
  double s6[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) s6[k] = sums[c * 6 + k];
  bn3_finalize_fwd_channel(s6, count, c, bnw, bnb, rmean, rvar, eps, momentum, C, scale, shift, mean, istd);
}

// ---- SyncBatchNorm statistics exchange fused into the finalize kernels: one-shot all-reduce over NVLink peer memory --------
// Every rank holds a SYMMETRIC buffer (same layout on all ranks, mapped into every peer's address space: torch's
// symmetric-memory allocator supplies the memory and the peer pointers, the exchange is this kernel).  A rank's
// payload for this call site sits at `slot_off`; `flag_off` addresses a row of `world` 32-bit flags for the site.
//   1. block 0 stores the site's epoch into flag[my rank] of every peer (release, system scope): "my payload is written"
//      (it was written by earlier kernels on this stream);
//   2. every block waits until its own flag row shows the epoch from every rank (acquire, system scope);
//   3. every thread reads its channel's values from all ranks through the peer pointers (fixed rank order: the sums are
//      bitwise identical on all ranks) and runs the BatchNorm finalize on them.
// Two ranks can be at most one call site apart (a rank passes site s + 1 only after every peer has signalled it, which a
// peer does after its site-s kernel finished), so per-site slots are never overwritten while a peer still reads them.
// The epoch lives on the device (one counter per site, bumped by a 1-thread kernel right after) -> CUDA-graph replayable.
constexpr int kMaxPeers = 8;
struct PeerTable { const uint8_t* base[kMaxPeers]; };
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ double ld_sys_f64(const double* p) { double v; asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float ld_sys_f32(const float* p) { float v; asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory"); return v; }

__device__ __forceinline__ void peer_signal_and_wait(const PeerTable& pt, size_t flag_off, int rank, int world, uint32_t epoch) {
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(pt.base[threadIdx.x]) + flag_off) + rank, epoch);
  }
  if (threadIdx.x < world) {
    const uint32_t* f = reinterpret_cast<const uint32_t*>(pt.base[rank] + flag_off) + threadIdx.x;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
      if (clock64() - t0 > 20000000000ll) __trap();      // ~10 s: a peer died; fail loudly instead of hanging the GPU
    }
  }
  __syncthreads();
}

__global__ void bn3_finalize_fwd_sync_kernel(PeerTable pt, size_t slot_off, size_t flag_off, int rank, int world,
                                             const uint32_t* __restrict__ epoch_dev, P3 bnw, P3 bnb, M3 rmean, M3 rvar,
                                             float eps, float momentum, int C, float* __restrict__ scale,
                                             float* __restrict__ shift, float* __restrict__ mean, float* __restrict__ istd) {
  peer_signal_and_wait(pt, flag_off, rank, world, *epoch_dev + 1u);
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, count = 0.0;
  for (int r = 0; r < world; ++r) {
    const double* ps = reinterpret_cast<const double*>(pt.base[r] + slot_off);
#pragma unroll
    for (int k = 0; k < 6; ++k) s6[k] += ld_sys_f64(ps + c * 6 + k);
    count += ld_sys_f64(ps + C * 6);
  }
  bn3_finalize_fwd_channel(s6, count, c, bnw, bnb, rmean, rvar, eps, momentum, C, scale, shift, mean, istd);
  if (c == 0) const_cast<double*>(reinterpret_cast<const double*>(pt.base[rank] + slot_off))[C * 6 + 1] = count;   // for backward
}
__global__ void epoch_bump_kernel(uint32_t* e) { *e += 1u; }

// eval mode: scale/shift from the running statistics
__global__ void bn3_eval_affine_kernel(P3 bnw, P3 bnb, P3 rmean, P3 rvar, float eps, int C,
                                       float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float sh = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float sc = bnw.p[i][c] * rsqrtf(rvar.p[i][c] + eps);
    scale[i * C + c] = sc;
    sh += bnb.p[i][c] - rmean.p[i][c] * sc;
  }
  shift[c] = sh;
}
// S: [4][C] = sum du, sum du*y_i (global) ; out coef [9][C], dbnw[3][C], dbnb[3][C]
// S_local (may be NULL = S): this rank's own sums.  The parameter gradients come from them, as torch's SyncBatchNorm
// takes grad_weight / grad_bias before its all-reduce (torch/nn/modules/_functions.py:140-160) and leaves their
// averaging to the data-parallel wrapper; the dy coefficients need the global sums.
__device__ __forceinline__ void bn3_finalize_bwd_channel(const double* Sg, const double* Sl, double count, int c, const P3& bnw,
                                                         const float* __restrict__ mean, const float* __restrict__ istd, int C,
                                                         float* __restrict__ coef, float* __restrict__ dbnw,
                                                         float* __restrict__ dbnb) {
  const double S0 = Sg[0], L0 = Sl[0];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double m = mean[i * C + c], is = istd[i * C + c], w = bnw.p[i][c];
    const double Si = Sg[i + 1];
    const double D = is * (Si - m * S0);            // sum du * yhat_i
    const double a = w * is;
    coef[i * C + c] = (float)a;
    coef[(3 + i) * C + c] = (float)(-a * is * D / count);
    coef[(6 + i) * C + c] = (float)(-a * S0 / count + a * is * m * D / count);
    dbnw[i * C + c] = (float)(is * (Sl[i + 1] - m * L0));
    dbnb[i * C + c] = (float)L0;
  }
}
__global__ void bn3_finalize_bwd_kernel(const float* __restrict__ S, const float* __restrict__ S_local, double count,
                                        const double* __restrict__ count_dev, P3 bnw,
                                        const float* __restrict__ mean, const float* __restrict__ istd, int C,
                                        float* __restrict__ coef, float* __restrict__ dbnw, float* __restrict__ dbnb) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (count_dev) count = *count_dev;
  if (!S_local) S_local = S;
  double Sg[4], Sl[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { Sg[k] = S[k * C + c]; Sl[k] = S_local[k * C + c]; }
  bn3_finalize_bwd_channel(Sg, Sl, count, c, bnw, mean, istd, C, coef, dbnw, dbnb);
}
// the same with the one-shot NVLink exchange of S (float [4][C] at slot_off of every rank's symmetric buffer)
__global__ void bn3_finalize_bwd_sync_kernel(PeerTable pt, size_t slot_off, size_t flag_off, int rank, int world,
                                             const uint32_t* __restrict__ epoch_dev, const double* __restrict__ count_dev,
                                             P3 bnw, const float* __restrict__ mean, const float* __restrict__ istd, int C,
                                             float* __restrict__ coef, float* __restrict__ dbnw, float* __restrict__ dbnb) {
  peer_signal_and_wait(pt, flag_off, rank, world, *epoch_dev + 1u);
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double Sg[4] = {0.0, 0.0, 0.0, 0.0}, Sl[4];
  for (int r = 0; r < world; ++r) {
    const float* ps = reinterpret_cast<const float*>(pt.base[r] + slot_off);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float v = ld_sys_f32(ps + k * C + c);
      Sg[k] += (double)v;
      if (r == rank) Sl[k] = (double)v;
    }
  }
  bn3_finalize_bwd_channel(Sg, Sl, *count_dev, c, bnw, mean, istd, C, coef, dbnw, dbnb);
}

// out[c] = sum_r part[r][c], rows added in a fixed order (the per-CTA partial rows of the kernels above)
constexpr int kCsCols = 32, kCsGroups = 16;
__global__ void __launch_bounds__(kCsCols * kCsGroups)
colsum_kernel(const float* __restrict__ part, int rows, int cols, float* __restrict__ out) {
  __shared__ float red[kCsGroups][kCsCols + 1];
  const int tx = threadIdx.x % kCsCols, ty = threadIdx.x / kCsCols;
  const int c = blockIdx.x * kCsCols + tx;
  float a0 = 0.f, a1 = 0.f;
  if (c < cols) {
    int r = ty;
    for (; r + kCsGroups < rows; r += 2 * kCsGroups) {
      a0 += part[(size_t)r * cols + c];
      a1 += part[(size_t)(r + kCsGroups) * cols + c];
    }
    if (r < rows) a0 += part[(size_t)r * cols + c];
  }
  red[ty][tx] = a0 + a1;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kCsGroups; ++k) t += red[k][tx];
    out[c] = t;
  }
}

// ---- host launchers -------------------------------------------------------------------------
// few rows x very many columns (split-K partials of a weight gradient): one thread per 4 columns, rows added in order
__global__ void __launch_bounds__(256) colsum_wide_kernel(const float4* __restrict__ part, int rows, int cols4,
                                                          float4* __restrict__ out) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols4; c += gridDim.x * blockDim.x) {
    float4 a = part[c];
    for (int r = 1; r < rows; ++r) {
      const float4 b = part[(size_t)r * cols4 + c];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    out[c] = a;
  }
}
// fp32 [R][Cc] -> bf16 [R][Cc] and bf16 [Cc][R] in one pass (the two operand layouts of an nn.Linear weight: forward
// and data-gradient GEMMs both want their contraction axis contiguous); 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) cast_transpose_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wb,
                                                             __nv_bfloat16* __restrict__ wt, int R, int Cc) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < Cc) {
      v = w[(size_t)r * Cc + c];
      wb[(size_t)r * Cc + c] = __float2bfloat16_rn(v);
    }
    tile[ty + 8 * k][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, r = r0 + tx;
    if (r < R && c < Cc) wt[(size_t)c * R + r] = __float2bfloat16_rn(tile[tx][ty + 8 * k]);
  }
}
int cast_transpose(const float* w, void* wb, void* wt, int R, int Cc, cudaStream_t st) {
  dim3 grid((Cc + 31) / 32, (R + 31) / 32);
  cast_transpose_kernel<<<grid, 256, 0, st>>>(w, (__nv_bfloat16*)wb, (__nv_bfloat16*)wt, R, Cc);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

int colsum(const float* part, int rows, int cols, float* out, cudaStream_t st) {
  if (rows <= 16 && cols >= 16384 && cols % 4 == 0 && ((reinterpret_cast<uintptr_t>(part) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const int cols4 = cols / 4;
    int grid = (cols4 + 255) / 256;
    if (grid > 8 * sm_count()) grid = 8 * sm_count();
    colsum_wide_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(part), rows, cols4, reinterpret_cast<float4*>(out));
    SLAK_CUDA_TRY(cudaGetLastError());
    return SLAK_OK;
  }
  colsum_kernel<<<(cols + kCsCols - 1) / kCsCols, kCsCols * kCsGroups, 0, st>>>(part, rows, cols, out);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

static Geo make_geo(int N, int C, int HW, int extra_floats_per_c, size_t* smem_bytes) {
  Geo g{};
  g.N = N; g.C = C; g.HW = HW;
  // tile of PIX pixels x C channels in fp32; PIX a multiple of 8, about 24 KB (several CTAs per SM: the
  // phases of these kernels are latency-bound, occupancy is what hides it)
  static int tile_kb = 0;                 // SLAK_GLUE_TILE_KB: tuning knob for the tile size (default 24 KB)
  if (tile_kb == 0) { const char* e = getenv("SLAK_GLUE_TILE_KB"); tile_kb = e ? atoi(e) : 24; if (tile_kb < 8 || tile_kb > 128) tile_kb = 24; }
  int pix = (tile_kb * 1024 / 4) / C;
  pix = pix >= 128 ? 128 : (pix >= 64 ? 64 : (pix >= 32 ? 32 : (pix >= 16 ? 16 : 8)));
  while (pix > 8 && pix / 2 >= HW) pix /= 2;
  g.PIX = pix;
  g.pitch = pix + 1;
  g.tiles_per_img = (HW + pix - 1) / pix;
  g.total_tiles = g.tiles_per_img * N;
  *smem_bytes = ((size_t)C * g.pitch + (size_t)extra_floats_per_c * C) * sizeof(float);
  return g;
}
static int pick_vp(int HW, const void* a, const void* b, const void* c, const void* d) {
  const uintptr_t m = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
                      reinterpret_cast<uintptr_t>(d);
  if (HW % 8 == 0 && (m & 15) == 0) return 8;
  if (HW % 4 == 0 && (m & 7) == 0) return 4;
  return 1;
}
static int grid_for(const Geo& g, size_t smem) {
  int per_sm = smem > 0 ? (int)(200 * 1024 / smem) : 4;
  if (per_sm > 6) per_sm = 6;
  if (per_sm < 1) per_sm = 1;
  int grid = sm_count() * per_sm;
  if (grid > g.total_tiles) grid = g.total_tiles;
  return grid < 1 ? 1 : grid;
}

#define SLAK_VP_DISPATCH(vp, CALL)               \
  do {                                           \
    if ((vp) == 8) { CALL(8); }                  \
    else if ((vp) == 4) { CALL(4); }             \
    else { CALL(1); }                            \
  } while (0)

int bn3_stats_finalize(const float* part, int splits, double* sums_ws, int C, cudaStream_t st) {
  bn3_reduce_partials_kernel<<<(C * 6 + 127) / 128, 128, 0, st>>>(part, splits, C, sums_ws);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}
int bn3_finalize_fwd(const double* sums, double count, const double* count_dev, const float* const* bnw,
                     const float* const* bnb, float* const* rmean, float* const* rvar, float eps, float momentum, int C,
                     float* scale, float* shift, float* mean, float* istd, cudaStream_t st) {
  P3 w{{bnw[0], bnw[1], bnw[2]}}, b{{bnb[0], bnb[1], bnb[2]}};
  M3 rm{{rmean[0], rmean[1], rmean[2]}}, rv{{rvar[0], rvar[1], rvar[2]}};
  bn3_finalize_fwd_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, count, count_dev, w, b, rm, rv, eps, momentum, C, scale, shift, mean, istd);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}
static int fill_peers(PeerTable* pt, const void* const* peers, int world) {
  SLAK_REQUIRE(world >= 2 && world <= kMaxPeers, SLAK_ERR_UNSUPPORTED, "world size %d outside 2..%d", world, kMaxPeers);
  for (int r = 0; r < kMaxPeers; ++r) pt->base[r] = r < world ? (const uint8_t*)peers[r] : nullptr;
  return SLAK_OK;
}
int bn3_finalize_fwd_sync(const void* const* peers, size_t slot_off, size_t flag_off, int rank, int world, uint32_t* epoch_dev,
                          const float* const* bnw, const float* const* bnb, float* const* rmean, float* const* rvar, float eps,
                          float momentum, int C, float* scale, float* shift, float* mean, float* istd, cudaStream_t st) {
  PeerTable pt;
  int rc = fill_peers(&pt, peers, world);
  if (rc) return rc;
  P3 w{{bnw[0], bnw[1], bnw[2]}}, b{{bnb[0], bnb[1], bnb[2]}};
  M3 rm{{rmean[0], rmean[1], rmean[2]}}, rv{{rvar[0], rvar[1], rvar[2]}};
  bn3_finalize_fwd_sync_kernel<<<(C + 127) / 128, 128, 0, st>>>(pt, slot_off, flag_off, rank, world, epoch_dev, w, b, rm, rv, eps,
                                                               momentum, C, scale, shift, mean, istd);
  epoch_bump_kernel<<<1, 1, 0, st>>>(epoch_dev);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}
int bn3_finalize_bwd_sync(const void* const* peers, size_t slot_off, size_t flag_off, int rank, int world, uint32_t* epoch_dev,
                          const double* count_dev, const float* const* bnw, const float* mean, const float* istd, int C,
                          float* coef, float* dbnw, float* dbnb, cudaStream_t st) {
  PeerTable pt;
  int rc = fill_peers(&pt, peers, world);
  if (rc) return rc;
  P3 w{{bnw[0], bnw[1], bnw[2]}};
  bn3_finalize_bwd_sync_kernel<<<(C + 127) / 128, 128, 0, st>>>(pt, slot_off, flag_off, rank, world, epoch_dev, count_dev, w, mean,
                                                               istd, C, coef, dbnw, dbnb);
  epoch_bump_kernel<<<1, 1, 0, st>>>(epoch_dev);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}
int bn3_eval_affine(const float* const* bnw, const float* const* bnb, const float* const* rmean, const float* const* rvar,
                    float eps, int C, float* scale, float* shift, cudaStream_t st) {
  P3 w{{bnw[0], bnw[1], bnw[2]}}, b{{bnb[0], bnb[1], bnb[2]}}, rm{{rmean[0], rmean[1], rmean[2]}}, rv{{rvar[0], rvar[1], rvar[2]}};
  bn3_eval_affine_kernel<<<(C + 127) / 128, 128, 0, st>>>(w, b, rm, rv, eps, C, scale, shift);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}
int bn3_finalize_bwd(const float* S, const float* S_local, double count, const double* count_dev, const float* const* bnw,
                     const float* mean, const float* istd, int C, float* coef, float* dbnw, float* dbnb, cudaStream_t st) {
  P3 w{{bnw[0], bnw[1], bnw[2]}};
  bn3_finalize_bwd_kernel<<<(C + 127) / 128, 128, 0, st>>>(S, S_local, count, count_dev, w, mean, istd, C, coef, dbnw, dbnb);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

int bn3_sum_ln_fwd(const void* y1, const void* y2, const void* y3, const float* scale, const float* shift,
                   const float* lnw, const float* lnb, float eps, void* xn, float* mu, float* rstd, int N, int C, int HW,
                   cudaStream_t st) {
  {
    const int rc = g2::ln_fwd(y1, y2, y3, scale, shift, lnw, lnb, eps, xn, mu, rstd, N, C, HW, st);
    if (rc != SLAK_G2_UNSUPPORTED) return rc;
  }
  size_t smem;
  Geo g = make_geo(N, C, HW, 2, &smem);              // u tile + lnw + lnb
  smem += (2 * (size_t)kThreads + 2 * (size_t)g.PIX) * sizeof(float);
  SLAK_REQUIRE(smem <= 200 * 1024, SLAK_ERR_UNSUPPORTED, "C=%d too large for the fused Block kernels", C);
  const int vp = pick_vp(HW, y1, y2, y3, y1);
  const int grid = grid_for(g, smem);
#define CALL(V)                                                                                                   \
  SLAK_SET_MAX_SMEM(bn3_sum_ln_fwd_kernel<V>, smem); \
  bn3_sum_ln_fwd_kernel<V><<<grid, kThreads, smem, st>>>((const __nv_bfloat16*)y1, (const __nv_bfloat16*)y2,        \
      (const __nv_bfloat16*)y3, scale, shift, lnw, lnb, eps, (__nv_bfloat16*)xn, mu, rstd, g)
  SLAK_VP_DISPATCH(vp, CALL);
#undef CALL
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

int residual_fwd(const float* x, const void* h2, const float* gamma, const float* dp, float* out, void* out_bf16,
                 int N, int C, int HW, cudaStream_t st) {
  {
    const int rc = g2::res_fwd(x, h2, gamma, dp, out, out_bf16, N, C, HW, st);
    if (rc != SLAK_G2_UNSUPPORTED) return rc;
  }
  size_t smem;
  Geo g = make_geo(N, C, HW, 0, &smem);
  SLAK_REQUIRE(smem <= 200 * 1024, SLAK_ERR_UNSUPPORTED, "C=%d too large for the fused Block kernels", C);
  int vp = pick_vp(HW, x, out, out_bf16, x);
  if (vp == 8 && HW % 8 != 0) vp = 1;
  const int grid = grid_for(g, smem);
#define CALL(V)                                                                                                   \
  SLAK_SET_MAX_SMEM(residual_fwd_kernel<V>, smem); \
  residual_fwd_kernel<V><<<grid, kThreads, smem, st>>>(x, (const __nv_bfloat16*)h2, gamma, dp, out, (__nv_bfloat16*)out_bf16, g)
  SLAK_VP_DISPATCH(vp, CALL);
#undef CALL
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

int residual_bwd_parts(int N, int C, int HW) {
  if (const int p2 = g2::res_bwd_parts(N, C, HW)) return p2;
  size_t smem;
  Geo g = make_geo(N, C, HW, 2 * kWarps, &smem);
  return grid_for(g, smem);
}
int gelu_bwd_bias_parts(long long rows, int K) {
  (void)K;
  long long ctas = 4LL * sm_count();
  if (ctas > rows) ctas = rows;
  return (int)(ctas < 1 ? 1 : ctas);
}
int gelu_bwd_bias(const void* da, const void* h, void* dh, float* part, long long rows, int K, cudaStream_t st) {
  SLAK_REQUIRE(K % 8 == 0 && K > 0, SLAK_ERR_UNSUPPORTED, "hidden width %d must be a multiple of 8", K);
  const int grid = gelu_bwd_bias_parts(rows, K);
  const int rows_per_cta = (int)((rows + grid - 1) / grid);
  const int nv = K / 8;
  const int rsubs = nv >= kThreads ? 1 : kThreads / nv;
  const size_t smem = (size_t)rsubs * K * sizeof(float);
  SLAK_REQUIRE(smem <= 200 * 1024, SLAK_ERR_UNSUPPORTED, "hidden width %d too large", K);
  SLAK_SET_MAX_SMEM(gelu_bwd_bias_kernel, smem);
  gelu_bwd_bias_kernel<<<grid, kThreads, smem, st>>>((const __nv_bfloat16*)da, (const __nv_bfloat16*)h, (__nv_bfloat16*)dh,
                                                    part, rows, K, rows_per_cta);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}
int residual_bwd(const float* dout, const void* h2, const float* gamma, const float* dp, void* dh2, float* dgamma_part,
                 int N, int C, int HW, cudaStream_t st) {
  {
    const int rc = g2::res_bwd(dout, h2, gamma, dp, dh2, dgamma_part, N, C, HW, st);
    if (rc != SLAK_G2_UNSUPPORTED) return rc;
  }
  size_t smem;
  Geo g = make_geo(N, C, HW, 2 * kWarps, &smem);
  SLAK_REQUIRE(smem <= 200 * 1024, SLAK_ERR_UNSUPPORTED, "C=%d too large for the fused Block kernels", C);
  const int vp = pick_vp(HW, dout, dout, dout, dout);
  const int grid = residual_bwd_parts(N, C, HW);      // one partial row per CTA: the caller sized the buffer with this
#define CALL(V)                                                                                                   \
  SLAK_SET_MAX_SMEM(residual_bwd_kernel<V>, smem); \
  residual_bwd_kernel<V><<<grid, kThreads, smem, st>>>(dout, (const __nv_bfloat16*)h2, gamma, dp, (__nv_bfloat16*)dh2, dgamma_part, g)
  SLAK_VP_DISPATCH(vp, CALL);
#undef CALL
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

static Geo ln_bwd_geo(int N, int C, int HW, size_t* smem) {
  Geo g = make_geo(N, C, HW, 7, smem);              // u tile + accs [6][C] + lnw [C]
  *smem += (4 * (size_t)g.PIX + 2 * (size_t)kThreads) * sizeof(float) + (size_t)g.PIX * ln_bwd_gpitch(C) * sizeof(__nv_bfloat16);
  return g;
}
int bn3_sum_ln_bwd_parts(int N, int C, int HW) {
  if (const int p2 = g2::ln_bwd_parts(N, C, HW)) return p2;
  size_t smem;
  Geo g = ln_bwd_geo(N, C, HW, &smem);
  return grid_for(g, smem);
}
int bn3_sum_ln_bwd(const void* dxn, const void* y1, const void* y2, const void* y3, const float* scale,
                   const float* shift, const float* lnw, const float* mu, const float* rstd, void* du, float* part,
                   int N, int C, int HW, cudaStream_t st) {
  {
    const int rc = g2::ln_bwd(dxn, y1, y2, y3, scale, shift, lnw, mu, rstd, du, part, N, C, HW, st);
    if (rc != SLAK_G2_UNSUPPORTED) return rc;
  }
  size_t smem;
  Geo g = ln_bwd_geo(N, C, HW, &smem);
  SLAK_REQUIRE(smem <= 200 * 1024, SLAK_ERR_UNSUPPORTED, "C=%d too large for the fused Block kernels", C);
  const int vp = pick_vp(HW, y1, y2, y3, du);
  const int grid = bn3_sum_ln_bwd_parts(N, C, HW);    // one partial row per CTA: the caller sized the buffer with this
#define CALL(V)                                                                                                   \
  SLAK_SET_MAX_SMEM(bn3_sum_ln_bwd_kernel<V>, smem); \
  bn3_sum_ln_bwd_kernel<V><<<grid, kThreads, smem, st>>>((const __nv_bfloat16*)dxn, (const __nv_bfloat16*)y1,       \
      (const __nv_bfloat16*)y2, (const __nv_bfloat16*)y3, scale, shift, lnw, mu, rstd, (__nv_bfloat16*)du, part, g)
  SLAK_VP_DISPATCH(vp, CALL);
#undef CALL
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

int bn3_bwd_apply(const void* du, const void* y1, const void* y2, const void* y3, const float* coef, void* dy1,
                  void* dy2, void* dy3, int N, int C, int HW, cudaStream_t st) {
  int vp = pick_vp(HW, du, y1, y2, y3);
  const int vp2 = pick_vp(HW, dy1, dy2, dy3, dy1);
  if (vp2 < vp) vp = vp2;
  const size_t planes = (size_t)N * C;
  if (vp == 1 && HW >= 8 && (planes * HW) % 8 == 0 && pick_vp(8, du, y1, y2, y3) == 8 && pick_vp(8, dy1, dy2, dy3, dy1) == 8) {
    const size_t total = planes * HW;
    int grid = (int)((total / 8 + kThreads - 1) / kThreads);
    if (grid > 8 * sm_count()) grid = 8 * sm_count();
    bn3_bwd_apply_flat_kernel<<<grid, kThreads, 0, st>>>((const __nv_bfloat16*)du, (const __nv_bfloat16*)y1, (const __nv_bfloat16*)y2,
        (const __nv_bfloat16*)y3, coef, (__nv_bfloat16*)dy1, (__nv_bfloat16*)dy2, (__nv_bfloat16*)dy3, C, HW, total);
    SLAK_CUDA_TRY(cudaGetLastError());
    return SLAK_OK;
  }
  const size_t total = planes * (HW / vp);
  int grid = (int)((total + kThreads - 1) / kThreads);
  if (grid > 8 * sm_count()) grid = 8 * sm_count();
#define CALL(V)                                                                                                   \
  bn3_bwd_apply_kernel<V><<<grid, kThreads, 0, st>>>((const __nv_bfloat16*)du, (const __nv_bfloat16*)y1,            \
      (const __nv_bfloat16*)y2, (const __nv_bfloat16*)y3, coef, (__nv_bfloat16*)dy1, (__nv_bfloat16*)dy2,           \
      (__nv_bfloat16*)dy3, C, HW, planes)
  SLAK_VP_DISPATCH(vp, CALL);
#undef CALL
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

}  // namespace blk
}  // namespace slak
