// Second generation of the memory-bound glue of a SLaK Block (models/SLaK.py:89-100 BN+sum, :153-166 permute /
// LayerNorm / gamma / residual): the same four passes as block_fused.cu (bn3_sum_ln_fwd, residual_fwd, residual_bwd,
// bn3_sum_ln_bwd), restructured around the instruction budget instead of a shared-memory tile.
//
//  * A tile is C channels x PIX pixels of one image (C * PIX <= 6144).  Every thread owns three (channel, 8-pixel span)
//    items of it and keeps them in REGISTERS from the global load to the global store: the NCHW side is touched with
//    one 16-byte (8-byte for planes that are only 8-byte aligned) access per tensor and item, all issued before the
//    first use.  The thread's channels are the same for every tile, so per-channel reductions live in registers for
//    the whole kernel and are folded once at the end (fixed order: deterministic).
//  * Only the NHWC side goes through shared memory: rows of C channels are staged as bf16 (odd word pitch) and read /
//    written transposed with 2-byte accesses.
//  * Per-pixel reductions over channels: each thread adds its three channels in registers, the per-thread partials
//    meet in a [group][pixel] table that is summed in two conflict-free steps.
//  * LayerNorm backward is rewritten with per-pixel constants: du = (r*w_c)*g + beta_j*d + alpha_j with d = u - mu_j,
//    beta_j = -r^2 m2, alpha_j = -r m1 (three packed FMAs per two elements).  Arithmetic is packed fp32
//    (fma.rn.f32x2), sums of bf16 products (sum du*y_i) use the mixed-precision FMA (fma.rn.f32.bf16 = FHFMA.BF16 on
//    sm_100a), which needs no unpacking.
//
// Planes whose size is not a multiple of 4 pixels (7 x 7) use 2-byte accesses on the NCHW side (LW = 1).
#include "common.cuh"
#include "tc_common.cuh"
#include "block_glue2.cuh"
#include <stdlib.h>

namespace slak {
namespace blk {
namespace g2 {

using tc::f2;
using tc::mk2;
using tc::mk2u;
using tc::un2;
using tc::fma2;
using tc::mul2;
using tc::add2;
using tc::splat;
using tc::pack2;
using tc::unpack2;

constexpr int kThreads = 256;
constexpr int kItems = 3;                                   // (channel, span) items per thread and tile
constexpr int kSpan = 8;                                    // pixels per item
constexpr int kTileElems = kThreads * kItems * kSpan;       // 6144
constexpr int kMaxVec = kTileElems / 8 / kThreads;          // 16-byte vectors of NHWC rows per thread and tile (3)

struct G2 {
  int N, C, HW, PIX, vshift /* log2(PIX / 8) */, tiles_per_img, total_tiles, ge /* bf16 per staged NHWC row */;
  float invC;
  int W, Wo, Ho;          // downsampling (2 x 2, stride 2) patch layout of the NHWC side, 0 when unused
  unsigned magicW;        // floor(2^32 / W) + 1: p / W == umulhi(p, magicW) for p < 2^20, W < 2^10
};

__host__ __device__ inline int row_pitch(int C) {            // even, half of it odd: transposed 2-byte walks spread over banks
  int ge = (C + 1) & ~1;
  if (((ge / 2) & 1) == 0) ge += 2;
  return ge;
}

// ---- mixed-precision accumulation: c + a*b with a, b bf16 halves of packed words -------------------------------
__device__ __forceinline__ void fh_acc(float& acc, uint32_t a, uint32_t b) {       // acc += a.lo*b.lo + a.hi*b.hi
  unsigned short al, ah, bl, bh;
  asm("mov.b32 {%0,%1}, %2;" : "=h"(al), "=h"(ah) : "r"(a));
  asm("mov.b32 {%0,%1}, %2;" : "=h"(bl), "=h"(bh) : "r"(b));
  asm("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc) : "h"(al), "h"(bl));
  asm("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc) : "h"(ah), "h"(bh));
}
__device__ __forceinline__ void fh_add(float& acc, uint32_t a) {                   // acc += a.lo + a.hi
  unsigned short al, ah;
  asm("mov.b32 {%0,%1}, %2;" : "=h"(al), "=h"(ah) : "r"(a));
  asm("add.rn.f32.bf16 %0, %1, %0;" : "+f"(acc) : "h"(al));
  asm("add.rn.f32.bf16 %0, %1, %0;" : "+f"(acc) : "h"(ah));
}

// ---- NCHW side: one 8-pixel span of a plane ---------------------------------------------------------------------
// LW = widest aligned access in pixels (8: 16-byte bf16 vectors, 4: 8-byte, 1: element-wise); nv = valid pixels of the
// span (LW = 8: 8 or <= 0; LW = 4: 8, 4 or <= 0; LW = 1: anything).  Invalid pixels read as zero.
template <int LW>
__device__ __forceinline__ void ldg_span(const __nv_bfloat16* __restrict__ p, int nv, uint32_t (&r)[4]) {
  if constexpr (LW == 8) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (nv > 0) v = *reinterpret_cast<const uint4*>(p);
    r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
  } else if constexpr (LW == 4) {
    uint2 a = make_uint2(0u, 0u), b = make_uint2(0u, 0u);
    if (nv > 0) a = *reinterpret_cast<const uint2*>(p);
    if (nv > 4) b = *reinterpret_cast<const uint2*>(p + 4);
    r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y;
  } else {
    // 2-byte aligned span: up to three aligned 8-byte words cover it (instead of eight 2-byte loads with every lane of the
    // warp in another plane); only words that hold a valid element are touched, so nothing outside the tensor is read
    // (its size in bytes is a multiple of 8: C % 8 == 0)
    r[0] = r[1] = r[2] = r[3] = 0u;
    if (nv > 0) {
      const uintptr_t a = reinterpret_cast<uintptr_t>(p);
      const int sft = (int)(a & 7);
      const uint2* w = reinterpret_cast<const uint2*>(a - sft);
      const int need = sft + 2 * (nv < 8 ? nv : 8);                 // bytes from the start of the first word
      uint32_t W[6] = {0u, 0u, 0u, 0u, 0u, 0u};
      { const uint2 t = w[0]; W[0] = t.x; W[1] = t.y; }
      if (need > 8) { const uint2 t = w[1]; W[2] = t.x; W[3] = t.y; }
      if (need > 16) { const uint2 t = w[2]; W[4] = t.x; W[5] = t.y; }
      const bool up = (sft & 4) != 0;
      const uint32_t sh = (uint32_t)(sft & 2) * 8u;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t lo = up ? W[k + 1] : W[k], hi = up ? W[k + 2] : W[k + 1];
        uint32_t v = __funnelshift_r(lo, hi, sh);
        if (2 * k >= nv) v = 0u;
        else if (2 * k + 1 >= nv) v &= 0xffffu;
        r[k] = v;
      }
    }
  }
}
template <int LW>
__device__ __forceinline__ void stg_span(__nv_bfloat16* __restrict__ p, int nv, const uint32_t (&r)[4]) {
  if constexpr (LW == 8) {
    if (nv > 0) *reinterpret_cast<uint4*>(p) = make_uint4(r[0], r[1], r[2], r[3]);
  } else if constexpr (LW == 4) {
    if (nv > 0) *reinterpret_cast<uint2*>(p) = make_uint2(r[0], r[1]);
    if (nv > 4) *reinterpret_cast<uint2*>(p + 4) = make_uint2(r[2], r[3]);
  } else {
    unsigned short* q = reinterpret_cast<unsigned short*>(p);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (2 * k < nv) q[2 * k] = (unsigned short)(r[k] & 0xffffu);
      if (2 * k + 1 < nv) q[2 * k + 1] = (unsigned short)(r[k] >> 16);
    }
  }
}
template <int LW>
__device__ __forceinline__ void ldg_span_f32(const float* __restrict__ p, int nv, f2 (&r)[4]) {
  if constexpr (LW >= 4) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (nv > 0) a = *reinterpret_cast<const float4*>(p);
    if (nv > 4) b = *reinterpret_cast<const float4*>(p + 4);
    r[0] = mk2(a.x, a.y); r[1] = mk2(a.z, a.w); r[2] = mk2(b.x, b.y); r[3] = mk2(b.z, b.w);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = mk2(2 * k < nv ? p[2 * k] : 0.f, 2 * k + 1 < nv ? p[2 * k + 1] : 0.f);
  }
}
template <int LW>
__device__ __forceinline__ void stg_span_f32(float* __restrict__ p, int nv, const f2 (&r)[4]) {
  float v[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) un2(r[k], v[2 * k], v[2 * k + 1]);
  if constexpr (LW >= 4) {
    if (nv > 0) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    if (nv > 4) *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < nv) p[k] = v[k];
  }
}

// ---- NHWC side: rows of C bf16 (C % 8 == 0), one contiguous run per tile, staged as [PIX][ge] ----------------------
struct RowWalk { int j, c, dj, dc; };                          // first vector of this thread and the step of 256 vectors
__device__ __forceinline__ RowWalk row_walk(int C) {
  RowWalk w;
  const int e = 8 * (int)threadIdx.x;
  w.j = e / C; w.c = e - w.j * C;
  const int s = 8 * kThreads;
  w.dj = s / C; w.dc = s - w.dj * C;
  return w;
}
__device__ __forceinline__ void rows_ldg(const __nv_bfloat16* __restrict__ src, int nvec, uint4 (&r)[kMaxVec]) {
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int v = (int)threadIdx.x + i * kThreads;
    r[i] = make_uint4(0u, 0u, 0u, 0u);
    if (v < nvec) r[i] = *reinterpret_cast<const uint4*>(src + (size_t)v * 8);
  }
}
// rows at and beyond the valid ones are written as zeros (vectors up to tvec = PIX*C/8)
__device__ __forceinline__ void rows_sts(__nv_bfloat16* gs, const G2& g, const RowWalk& w, int tvec, const uint4 (&r)[kMaxVec]) {
  int j = w.j, c = w.c;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int v = (int)threadIdx.x + i * kThreads;
    if (v < tvec) {
      uint32_t* d = reinterpret_cast<uint32_t*>(gs + (size_t)j * g.ge + c);
      d[0] = r[i].x; d[1] = r[i].y; d[2] = r[i].z; d[3] = r[i].w;
    }
    c += w.dc; j += w.dj;
    if (c >= g.C) { c -= g.C; ++j; }
  }
}
__device__ __forceinline__ void rows_out(__nv_bfloat16* __restrict__ dst, const __nv_bfloat16* gs, const G2& g, const RowWalk& w,
                                         int nvec) {
  int j = w.j, c = w.c;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int v = (int)threadIdx.x + i * kThreads;
    if (v < nvec) {
      const uint32_t* s = reinterpret_cast<const uint32_t*>(gs + (size_t)j * g.ge + c);
      *reinterpret_cast<uint4*>(dst + (size_t)v * 8) = make_uint4(s[0], s[1], s[2], s[3]);
    }
    c += w.dc; j += w.dj;
    if (c >= g.C) { c -= g.C; ++j; }
  }
}
__device__ __forceinline__ uint32_t lds16(const __nv_bfloat16* p) { return *reinterpret_cast<const unsigned short*>(p); }
__device__ __forceinline__ void sts16(__nv_bfloat16* p, uint32_t v) { *reinterpret_cast<unsigned short*>(p) = (unsigned short)v; }
// transposed read: channel c of the 8 rows j0 .. j0+7 as fp32 pairs
__device__ __forceinline__ void lds_span_t(const __nv_bfloat16* gs, int ge, int j0, int c, f2 (&r)[4]) {
  const __nv_bfloat16* p = gs + (size_t)j0 * ge + c;
#pragma unroll
  for (int k = 0; k < 4; ++k) r[k] = mk2u(lds16(p + (2 * k) * ge) << 16, lds16(p + (2 * k + 1) * ge) << 16);
}
__device__ __forceinline__ void sts_span_t(__nv_bfloat16* gs, int ge, int j0, int c, const uint32_t (&r)[4]) {
  __nv_bfloat16* p = gs + (size_t)j0 * ge + c;
#pragma unroll
  for (int k = 0; k < 4; ++k) { sts16(p + (2 * k) * ge, r[k]); sts16(p + (2 * k + 1) * ge, r[k] >> 16); }
}

// ---- per-pixel sums over channels ----------------------------------------------------------------------------------
// Every thread holds partial sums (over its three channels) of NQ quantities for its 8 pixels.  They are written to
// red[group][q][PIX] (group = tid >> vshift); then NQ*PIX columns are summed: tpo = 256 / (NQ*PIX) threads per column
// (conflict-free, consecutive threads = consecutive columns) -> red2[sub][NQ*PIX]; the caller folds the tpo rows.
template <int NQ>
__device__ __forceinline__ void pix_partials_store(float* red, const G2& g, int sp, const f2 (&s)[NQ][4]) {
  float* base = red + (size_t)((int)threadIdx.x >> g.vshift) * (NQ * g.PIX) + sp * kSpan;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) un2(s[q][k], v[2 * k], v[2 * k + 1]);
    *reinterpret_cast<float4*>(base + q * g.PIX) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(base + q * g.PIX + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}
template <int NQ>   // NQ = 2 (the column count must be a power of two)
__device__ __forceinline__ void pix_partials_fold(const float* red, float* red2, const G2& g) {
  const int O = NQ * g.PIX;                                   // columns; a power of two times NQ
  const int ngroups = kThreads >> g.vshift;
  if (O >= kThreads) {                                        // one or more columns per thread
    for (int o = threadIdx.x; o < O; o += kThreads) {
      float s = 0.f;
      for (int gq = 0; gq < ngroups; ++gq) s += red[(size_t)gq * O + o];
      red2[o] = s;
    }
  } else {
    const int tpo = kThreads / O;                             // NQ = 2: O = 2^(vshift + 4)
    const int sub = (int)threadIdx.x >> (g.vshift + 4), o = (int)threadIdx.x & (O - 1);
    float s = 0.f;
    for (int gq = sub; gq < ngroups; gq += tpo) s += red[(size_t)gq * O + o];
    red2[sub * O + o] = s;
  }
}
template <int NQ>
__device__ __forceinline__ float pix_total(const float* red2, const G2& g, int q, int px) {
  const int O = NQ * g.PIX;
  const int tpo = O >= kThreads ? 1 : kThreads / O;
  float s = 0.f;
  for (int k = 0; k < tpo; ++k) s += red2[k * O + q * g.PIX + px];
  return s;
}

struct TileIdx { int n, p0, npix; };
__device__ __forceinline__ TileIdx tile_of(const G2& g, int t) {
  TileIdx ti;
  ti.n = t / g.tiles_per_img;
  ti.p0 = (t - ti.n * g.tiles_per_img) * g.PIX;
  ti.npix = min(g.PIX, g.HW - ti.p0);
  return ti;
}

// u = s1*y1 + s2*y2 + s3*y3 + sh - sub   for one span (packed fp32)
__device__ __forceinline__ void bn3_sum_span(const uint32_t (&q1)[4], const uint32_t (&q2)[4], const uint32_t (&q3)[4],
                                             const float4 ch, const f2 (&sub)[4], f2 (&u)[4]) {
  const f2 s1 = splat(ch.x), s2 = splat(ch.y), s3 = splat(ch.z), sh = splat(ch.w);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    u[k] = fma2(s1, unpack2(q1[k]), fma2(s2, unpack2(q2[k]), fma2(s3, unpack2(q3[k]), add2(sh, sub[k]))));
}

// ====================================================================================================================
// forward: xn[n,h,w,:] = LayerNorm_C(sum_i scale_i*y_i + shift)     NCHW bf16 x3 -> NHWC bf16, mu / rstd per pixel
// ====================================================================================================================
template <int LW>
__global__ void __launch_bounds__(kThreads, 2)
ln_fwd2_kernel(const __nv_bfloat16* __restrict__ y1, const __nv_bfloat16* __restrict__ y2, const __nv_bfloat16* __restrict__ y3,
               const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ lnw,
               const float* __restrict__ lnb, float eps, __nv_bfloat16* __restrict__ xn, float* __restrict__ mu,
               float* __restrict__ rstd, G2 g) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* red = reinterpret_cast<float*>(smem);                // [256 >> vshift][2][PIX] = 4096 floats
  float* red2 = red + 4096;                                   // [256]
  float* cst = red2 + kThreads;                               // [2][PIX]  r, -md*r
  float4* chs = reinterpret_cast<float4*>(cst + 2 * g.PIX);   // [C] s1 s2 s3 shift
  float* lnw_s = reinterpret_cast<float*>(chs + g.C);         // [C]
  float* lnb_s = lnw_s + g.C;                                 // [C]
  __nv_bfloat16* gs = reinterpret_cast<__nv_bfloat16*>(lnb_s + g.C);   // [PIX][ge]
  const int tid = threadIdx.x, C = g.C, HW = g.HW;
  const int sp = tid & ((1 << g.vshift) - 1), j0 = sp * kSpan;
  int ci[kItems];
#pragma unroll
  for (int i = 0; i < kItems; ++i) ci[i] = (tid + i * kThreads) >> g.vshift;
  // per-channel constants; pivot of the variance = mean over channels of the BN shift (the batch mean of the LN mean)
  float psum = 0.f;
  for (int c = tid; c < C; c += kThreads) {
    const float sh = shift[c];
    chs[c] = make_float4(scale[c], scale[C + c], scale[2 * C + c], sh);
    lnw_s[c] = lnw[c]; lnb_s[c] = lnb[c];
    psum += sh;
  }
  red2[tid] = psum;
  __syncthreads();
  float pivot = 0.f;
  for (int k = 0; k < kThreads; ++k) pivot += red2[k];
  pivot *= g.invC;
  const RowWalk rw = row_walk(C);
  uint32_t q[kItems][3][4];
  auto load_tile = [&](int t) {
    const TileIdx ti = tile_of(g, t);
    const int nv = ti.npix - j0;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      const int nvi = ci[i] < C ? nv : 0;
      const size_t off = ((size_t)ti.n * C + ci[i]) * HW + ti.p0 + j0;
      ldg_span<LW>(y1 + off, nvi, q[i][0]); ldg_span<LW>(y2 + off, nvi, q[i][1]); ldg_span<LW>(y3 + off, nvi, q[i][2]);
    }
  };
  int t = blockIdx.x;
  if (t < g.total_tiles) load_tile(t);
  const f2 npiv = splat(-pivot);
  const f2 npiv4[4] = {npiv, npiv, npiv, npiv};
  for (; t < g.total_tiles; t += gridDim.x) {
    const TileIdx ti = tile_of(g, t);
    f2 d[kItems][4];
    f2 s[2][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { s[0][k] = 0ull; s[1][k] = 0ull; }
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      if (ci[i] < C) {
        bn3_sum_span(q[i][0], q[i][1], q[i][2], chs[ci[i]], npiv4, d[i]);
#pragma unroll
        for (int k = 0; k < 4; ++k) { s[0][k] = add2(s[0][k], d[i][k]); s[1][k] = fma2(d[i][k], d[i][k], s[1][k]); }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) d[i][k] = 0ull;
      }
    }
    if (t + gridDim.x < g.total_tiles) load_tile(t + gridDim.x);      // next tile's loads fly during the rest of this one
    pix_partials_store<2>(red, g, sp, s);
    __syncthreads();
    pix_partials_fold<2>(red, red2, g);
    __syncthreads();
    if (tid < g.PIX) {
      const float md = pix_total<2>(red2, g, 0, tid) * g.invC;
      const float var = fmaxf(pix_total<2>(red2, g, 1, tid) * g.invC - md * md, 0.f);
      const float r = rsqrtf(var + eps);
      cst[tid] = r; cst[g.PIX + tid] = -md * r;
      if (tid < ti.npix) {
        const size_t pix = (size_t)ti.n * HW + ti.p0 + tid;
        mu[pix] = pivot + md; rstd[pix] = r;
      }
    }
    __syncthreads();
    {
      f2 r2[4], nm2[4];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(cst + j0 + 4 * k);
        const float4 b = *reinterpret_cast<const float4*>(cst + g.PIX + j0 + 4 * k);
        r2[2 * k] = mk2(a.x, a.y); r2[2 * k + 1] = mk2(a.z, a.w);
        nm2[2 * k] = mk2(b.x, b.y); nm2[2 * k + 1] = mk2(b.z, b.w);
      }
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        if (ci[i] < C) {
          const f2 w2 = splat(lnw_s[ci[i]]), b2 = splat(lnb_s[ci[i]]);
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k] = pack2(fma2(fma2(d[i][k], r2[k], nm2[k]), w2, b2));
          sts_span_t(gs, g.ge, j0, ci[i], o);
        }
      }
    }
    __syncthreads();
    rows_out(xn + ((size_t)ti.n * HW + ti.p0) * C, gs, g, rw, ti.npix * C / 8);
  }
}

// ====================================================================================================================
// forward: out = x + dp[n]*gamma[c]*h2[n,h,w,c]        NHWC bf16 -> NCHW fp32 (+ bf16 copy)
// ====================================================================================================================
template <int LW>
__global__ void __launch_bounds__(kThreads, 3)
res_fwd2_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ h2, const float* __restrict__ gamma,
                const float* __restrict__ dp, float* __restrict__ out, __nv_bfloat16* __restrict__ out_bf16, G2 g) {
  extern __shared__ __align__(16) unsigned char smem[];
  __nv_bfloat16* gs = reinterpret_cast<__nv_bfloat16*>(smem);
  const int tid = threadIdx.x, C = g.C, HW = g.HW;
  const int sp = tid & ((1 << g.vshift) - 1), j0 = sp * kSpan;
  int ci[kItems];
  float gam[kItems];
#pragma unroll
  for (int i = 0; i < kItems; ++i) { ci[i] = (tid + i * kThreads) >> g.vshift; gam[i] = ci[i] < C ? (gamma ? gamma[ci[i]] : 1.f) : 0.f; }
  const RowWalk rw = row_walk(C);
  const int tvec = g.PIX * C / 8;
  uint4 hr[kMaxVec];
  f2 xr[kItems][4];
  auto load_tile = [&](int t) {
    const TileIdx ti = tile_of(g, t);
    rows_ldg(h2 + ((size_t)ti.n * HW + ti.p0) * C, ti.npix * C / 8, hr);
    const int nv = ti.npix - j0;
#pragma unroll
    for (int i = 0; i < kItems; ++i)
      ldg_span_f32<LW>(x + ((size_t)ti.n * C + ci[i]) * HW + ti.p0 + j0, (x && ci[i] < C) ? nv : 0, xr[i]);
  };
  int t = blockIdx.x;
  if (t < g.total_tiles) load_tile(t);
  for (; t < g.total_tiles; t += gridDim.x) {
    const TileIdx ti = tile_of(g, t);
    const float dps = dp ? dp[ti.n] : 1.f;
    __syncthreads();                                         // the previous tile's transposed reads are done
    rows_sts(gs, g, rw, tvec, hr);
    f2 xc[kItems][4];
#pragma unroll
    for (int i = 0; i < kItems; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) xc[i][k] = xr[i][k];
    __syncthreads();
    if (t + gridDim.x < g.total_tiles) load_tile(t + gridDim.x);
    const int nv = ti.npix - j0;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      if (ci[i] < C && nv > 0) {
        f2 h[4], o[4];
        lds_span_t(gs, g.ge, j0, ci[i], h);
        const f2 gd = splat(gam[i] * dps);
        uint32_t ob[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { o[k] = fma2(h[k], gd, xc[i][k]); ob[k] = pack2(o[k]); }
        const size_t off = ((size_t)ti.n * C + ci[i]) * HW + ti.p0 + j0;
        stg_span_f32<LW>(out + off, nv, o);
        if (out_bf16) stg_span<LW>(out_bf16 + off, nv, ob);
      }
    }
  }
}

// ====================================================================================================================
// backward: d_h2 = dOut*gamma*dp (NHWC bf16); per-CTA partials [2][C]: dgamma = sum dOut*dp*h2, column sums of d_h2
// ====================================================================================================================
template <int LW>
__global__ void __launch_bounds__(kThreads, 3)
res_bwd2_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ h2, const float* __restrict__ gamma,
                const float* __restrict__ dp, __nv_bfloat16* __restrict__ dh2, float* __restrict__ part /*[grid][2][C]*/, G2 g) {
  extern __shared__ __align__(16) unsigned char smem[];
  __nv_bfloat16* gs = reinterpret_cast<__nv_bfloat16*>(smem);
  const int tid = threadIdx.x, C = g.C, HW = g.HW;
  const int spr = 1 << g.vshift, sp = tid & (spr - 1), j0 = sp * kSpan;
  int ci[kItems];
  float gam[kItems];
  f2 acc[kItems];
  float acc2[kItems];
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    ci[i] = (tid + i * kThreads) >> g.vshift; gam[i] = ci[i] < C ? (gamma ? gamma[ci[i]] : 1.f) : 0.f;
    acc[i] = 0ull; acc2[i] = 0.f;
  }
  const RowWalk rw = row_walk(C);
  const int tvec = g.PIX * C / 8;
  uint4 hr[kMaxVec];
  f2 dr[kItems][4];
  auto load_tile = [&](int t) {
    const TileIdx ti = tile_of(g, t);
    rows_ldg(h2 + ((size_t)ti.n * HW + ti.p0) * C, h2 ? ti.npix * C / 8 : 0, hr);
    const int nv = ti.npix - j0;
#pragma unroll
    for (int i = 0; i < kItems; ++i)
      ldg_span_f32<LW>(dout + ((size_t)ti.n * C + ci[i]) * HW + ti.p0 + j0, ci[i] < C ? nv : 0, dr[i]);
  };
  int t = blockIdx.x;
  if (t < g.total_tiles) load_tile(t);
  for (; t < g.total_tiles; t += gridDim.x) {
    const TileIdx ti = tile_of(g, t);
    const float dps = dp ? dp[ti.n] : 1.f;
    __syncthreads();                                         // the previous tile's rows have been copied out
    rows_sts(gs, g, rw, tvec, hr);                           // rows beyond npix are zeros
    f2 dc[kItems][4];
#pragma unroll
    for (int i = 0; i < kItems; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) dc[i][k] = dr[i][k];
    __syncthreads();
    if (t + gridDim.x < g.total_tiles) load_tile(t + gridDim.x);
    const f2 dps2 = splat(dps);
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      if (ci[i] < C) {
        f2 h[4];
        lds_span_t(gs, g.ge, j0, ci[i], h);
        const f2 gm = splat(gam[i]);
        uint32_t ob[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const f2 g0 = mul2(dc[i][k], dps2);
          ob[k] = pack2(mul2(g0, gm));
          acc[i] = fma2(g0, h[k], acc[i]);
          fh_add(acc2[i], ob[k]);
        }
        sts_span_t(gs, g.ge, j0, ci[i], ob);                 // in place: this thread owns these 8 elements
      }
    }
    __syncthreads();
    rows_out(dh2 + ((size_t)ti.n * HW + ti.p0) * C, gs, g, rw, ti.npix * C / 8);
  }
  // fold the spans of a channel (adjacent lanes), fixed order
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    float lo, hi;
    un2(acc[i], lo, hi);
    float a = lo + hi, b = acc2[i];
    for (int o = 1; o < spr; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (sp == 0 && ci[i] < C) {
      part[(size_t)blockIdx.x * 2 * C + ci[i]] = a;
      part[(size_t)blockIdx.x * 2 * C + C + ci[i]] = b;
    }
  }
}

// ====================================================================================================================
// backward: LayerNorm backward + BatchNorm reductions.  part per CTA: [6][C] = dlnw, dlnb, S0 = sum du, S1..S3 = sum du*y_i
// ====================================================================================================================
template <int LW>
__global__ void __launch_bounds__(kThreads, 2)
ln_bwd2_kernel(const __nv_bfloat16* __restrict__ dxn, const __nv_bfloat16* __restrict__ y1, const __nv_bfloat16* __restrict__ y2,
               const __nv_bfloat16* __restrict__ y3, const float* __restrict__ scale, const float* __restrict__ shift,
               const float* __restrict__ lnw, const float* __restrict__ mu, const float* __restrict__ rstd,
               __nv_bfloat16* __restrict__ du, float* __restrict__ part /*[grid][6][C]*/, G2 g) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint4* ys = reinterpret_cast<uint4*>(smem);                 // [3][kItems][256]: this thread's raw y spans (thread-private)
  float* red = reinterpret_cast<float*>(ys + 3 * kItems * kThreads);   // 4096 floats
  float* red2 = red + 4096;                                   // [256]
  float* cst = red2 + kThreads;                               // [3][PIX] r, alpha, beta
  float* mus = cst + 3 * g.PIX;                               // [PIX]
  float* rs = mus + g.PIX;                                    // [PIX]
  float4* chs = reinterpret_cast<float4*>(rs + g.PIX);        // [C]
  __nv_bfloat16* gs = reinterpret_cast<__nv_bfloat16*>(chs + g.C);     // [PIX][ge]
  const int tid = threadIdx.x, C = g.C, HW = g.HW;
  const int spr = 1 << g.vshift, sp = tid & (spr - 1), j0 = sp * kSpan;
  int ci[kItems];
  float wi[kItems];
  f2 A0[kItems], A1[kItems];
  float a2[kItems], a3[kItems], a4[kItems], a5[kItems];
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    ci[i] = (tid + i * kThreads) >> g.vshift; wi[i] = ci[i] < C ? lnw[ci[i]] : 0.f;
    A0[i] = 0ull; A1[i] = 0ull; a2[i] = a3[i] = a4[i] = a5[i] = 0.f;
  }
  for (int c = tid; c < C; c += kThreads) chs[c] = make_float4(scale[c], scale[C + c], scale[2 * C + c], shift[c]);
  const RowWalk rw = row_walk(C);
  const int tvec = g.PIX * C / 8;
  for (int t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
    const TileIdx ti = tile_of(g, t);
    const int nv = ti.npix - j0;
    // every global load of the tile is issued here
    uint4 gr[kMaxVec];
    uint32_t q[kItems][3][4];
    rows_ldg(dxn + ((size_t)ti.n * HW + ti.p0) * C, ti.npix * C / 8, gr);
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      const int nvi = ci[i] < C ? nv : 0;
      const size_t off = ((size_t)ti.n * C + ci[i]) * HW + ti.p0 + j0;
      ldg_span<LW>(y1 + off, nvi, q[i][0]); ldg_span<LW>(y2 + off, nvi, q[i][1]); ldg_span<LW>(y3 + off, nvi, q[i][2]);
    }
    float mval = 0.f, rval = 0.f;
    if (tid < ti.npix) { mval = mu[(size_t)ti.n * HW + ti.p0 + tid]; rval = rstd[(size_t)ti.n * HW + ti.p0 + tid]; }
    rows_sts(gs, g, rw, tvec, gr);                           // (the previous tile's readers of gs / mus / rs are past their last barrier)
    if (tid < g.PIX) { mus[tid] = mval; rs[tid] = rval; }
    __syncthreads();
    // phase B: d = u - mu, g (transposed), per-pixel partials S1 = sum g*w, P = sum g*w*d
    f2 d[kItems][4], gg[kItems][4];
    {
      f2 nm[4];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(mus + j0 + 4 * k);
        nm[2 * k] = mk2(-a.x, -a.y); nm[2 * k + 1] = mk2(-a.z, -a.w);
      }
      f2 s[2][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { s[0][k] = 0ull; s[1][k] = 0ull; }
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        if (ci[i] < C) {
          bn3_sum_span(q[i][0], q[i][1], q[i][2], chs[ci[i]], nm, d[i]);
          lds_span_t(gs, g.ge, j0, ci[i], gg[i]);
          const f2 w2 = splat(wi[i]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const f2 gw = mul2(gg[i][k], w2);
            s[0][k] = add2(s[0][k], gw);
            s[1][k] = fma2(gw, d[i][k], s[1][k]);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) { d[i][k] = 0ull; gg[i][k] = 0ull; }
        }
#pragma unroll
        for (int tsr = 0; tsr < 3; ++tsr)
          ys[(tsr * kItems + i) * kThreads + tid] = make_uint4(q[i][tsr][0], q[i][tsr][1], q[i][tsr][2], q[i][tsr][3]);
      }
      pix_partials_store<2>(red, g, sp, s);
    }
    __syncthreads();
    pix_partials_fold<2>(red, red2, g);
    __syncthreads();
    if (tid < g.PIX) {
      const float r = rs[tid];                               // 0 beyond the valid pixels: du = 0 there
      const float m1 = pix_total<2>(red2, g, 0, tid) * g.invC;
      const float m2 = r * pix_total<2>(red2, g, 1, tid) * g.invC;
      cst[tid] = r; cst[g.PIX + tid] = -r * m1; cst[2 * g.PIX + tid] = -r * r * m2;
    }
    __syncthreads();
    // phase C: du, dlnw / dlnb and the BatchNorm sums
    {
      f2 r2[4], al[4], be[4];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(cst + j0 + 4 * k);
        const float4 b = *reinterpret_cast<const float4*>(cst + g.PIX + j0 + 4 * k);
        const float4 e = *reinterpret_cast<const float4*>(cst + 2 * g.PIX + j0 + 4 * k);
        r2[2 * k] = mk2(a.x, a.y); r2[2 * k + 1] = mk2(a.z, a.w);
        al[2 * k] = mk2(b.x, b.y); al[2 * k + 1] = mk2(b.z, b.w);
        be[2 * k] = mk2(e.x, e.y); be[2 * k + 1] = mk2(e.z, e.w);
      }
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        if (ci[i] < C && nv > 0) {
          const f2 w2 = splat(wi[i]);
          uint32_t ob[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const f2 dv = fma2(mul2(r2[k], w2), gg[i][k], fma2(be[k], d[i][k], al[k]));
            A0[i] = fma2(gg[i][k], mul2(d[i][k], r2[k]), A0[i]);
            A1[i] = add2(A1[i], gg[i][k]);
            ob[k] = pack2(dv);
          }
          stg_span<LW>(du + ((size_t)ti.n * C + ci[i]) * HW + ti.p0 + j0, nv, ob);
          const uint4 q1 = ys[(0 * kItems + i) * kThreads + tid], q2 = ys[(1 * kItems + i) * kThreads + tid],
                      q3 = ys[(2 * kItems + i) * kThreads + tid];
          const uint32_t* p1 = reinterpret_cast<const uint32_t*>(&q1);
          const uint32_t* p2 = reinterpret_cast<const uint32_t*>(&q2);
          const uint32_t* p3 = reinterpret_cast<const uint32_t*>(&q3);
#pragma unroll
          for (int k = 0; k < 4; ++k) {                       // the bf16-rounded du is what BatchNorm backward will see
            fh_add(a2[i], ob[k]);
            fh_acc(a3[i], ob[k], p1[k]); fh_acc(a4[i], ob[k], p2[k]); fh_acc(a5[i], ob[k], p3[k]);
          }
        }
      }
    }
  }
  // fold the spans of a channel (adjacent lanes), fixed order
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    float v[6], lo, hi;
    un2(A0[i], lo, hi); v[0] = lo + hi;
    un2(A1[i], lo, hi); v[1] = lo + hi;
    v[2] = a2[i]; v[3] = a3[i]; v[4] = a4[i]; v[5] = a5[i];
    for (int o = 1; o < spr; o <<= 1) {
#pragma unroll
      for (int qn = 0; qn < 6; ++qn) v[qn] += __shfl_xor_sync(0xffffffffu, v[qn], o);
    }
    if (sp == 0 && ci[i] < C) {
#pragma unroll
      for (int qn = 0; qn < 6; ++qn) part[((size_t)blockIdx.x * 6 + qn) * C + ci[i]] = v[qn];
    }
  }
}


// ====================================================================================================================
// residual_fwd / nhwc_to_nchw for planes that are not a multiple of 4 pixels (7 x 7): a tile is CH channels x the WHOLE
// plane, which is contiguous in NCHW -- every access of the NCHW side is a coalesced 4-byte (fp32) or 2-byte (bf16)
// element per lane, the NHWC rows (CH channels = 128 bytes per pixel) are staged in shared memory.
// ====================================================================================================================
__global__ void __launch_bounds__(kThreads, 4)
res_fwd_flat_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ h2, const float* __restrict__ gamma,
                    const float* __restrict__ dp, float* __restrict__ out, __nv_bfloat16* __restrict__ out_bf16, int N, int C, int HW,
                    int CH) {
  extern __shared__ __align__(16) unsigned char smem[];
  __nv_bfloat16* hs = reinterpret_cast<__nv_bfloat16*>(smem);          // [HW][CH + 2]
  float* gs_ = reinterpret_cast<float*>(smem + (((size_t)HW * (CH + 2) * 2 + 15) & ~(size_t)15));   // [CH] gamma * dp
  const int tid = threadIdx.x, pitch = CH + 2;
  const int cchunks = C / CH, tiles = N * cchunks, tile_elems = CH * HW, vpr = CH / 8;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int n = t / cchunks, c0 = (t - n * cchunks) * CH;
    const float dps = dp ? dp[n] : 1.f;
    __syncthreads();
    for (int v = tid; v < HW * vpr; v += kThreads) {
      const int p = v / vpr, k = v - p * vpr;
      const uint4 r = *reinterpret_cast<const uint4*>(h2 + ((size_t)n * HW + p) * C + c0 + 8 * k);
      uint32_t* d = reinterpret_cast<uint32_t*>(hs + (size_t)p * pitch + 8 * k);
      d[0] = r.x; d[1] = r.y; d[2] = r.z; d[3] = r.w;
    }
    for (int c = tid; c < CH; c += kThreads) gs_[c] = (gamma ? gamma[c0 + c] : 1.f) * dps;
    __syncthreads();
    const size_t base = ((size_t)n * C + c0) * HW;
#pragma unroll 4
    for (int e = tid; e < tile_elems; e += kThreads) {
      const int c = e / HW, p = e - c * HW;
      const float v = fmaf(__bfloat162float(hs[(size_t)p * pitch + c]), gs_[c], x ? x[base + e] : 0.f);
      out[base + e] = v;
      if (out_bf16) out_bf16[base + e] = __float2bfloat16_rn(v);
    }
  }
}

// ====================================================================================================================
// Downsampling layer (models/SLaK.py:194-199: channels_first LayerNorm, then Conv2d(k=2, s=2)) as LayerNorm + GEMM:
// the LayerNorm writes its output directly as the GEMM's A operand, A[token = (n, h/2, w/2)][k = ((h&1)*2 + (w&1))*C + c]
// (bf16, one contiguous run of C channels per input pixel); the convolution is then a plain [tokens, 4C] x [4C, Cout]
// GEMM on the tcgen05 kernels of mlp_tc.cu.  Same thread <-> data mapping as above with u = x (fp32 NCHW).
// ====================================================================================================================
__device__ __forceinline__ size_t patch_row(const G2& g, int n, int p) {      // element offset of pixel p's row in A
  const int h = (int)__umulhi((unsigned)p, g.magicW), w = p - h * g.W;
  const size_t token = ((size_t)n * g.Ho + (h >> 1)) * g.Wo + (w >> 1);
  return token * (size_t)(4 * g.C) + (size_t)((((h & 1) << 1) | (w & 1)) * g.C);
}
__device__ __forceinline__ void rows_ldg_patch(const __nv_bfloat16* __restrict__ src, const G2& g, const RowWalk& w, int n, int p0,
                                               int nvec, uint4 (&r)[kMaxVec]) {
  int j = w.j, c = w.c;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int v = (int)threadIdx.x + i * kThreads;
    r[i] = make_uint4(0u, 0u, 0u, 0u);
    if (v < nvec) r[i] = *reinterpret_cast<const uint4*>(src + patch_row(g, n, p0 + j) + c);
    c += w.dc; j += w.dj;
    if (c >= g.C) { c -= g.C; ++j; }
  }
}
__device__ __forceinline__ void rows_out_patch(__nv_bfloat16* __restrict__ dst, const __nv_bfloat16* gs, const G2& g, const RowWalk& w,
                                               int n, int p0, int nvec) {
  int j = w.j, c = w.c;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int v = (int)threadIdx.x + i * kThreads;
    if (v < nvec) {
      const uint32_t* s = reinterpret_cast<const uint32_t*>(gs + (size_t)j * g.ge + c);
      *reinterpret_cast<uint4*>(dst + patch_row(g, n, p0 + j) + c) = make_uint4(s[0], s[1], s[2], s[3]);
    }
    c += w.dc; j += w.dj;
    if (c >= g.C) { c -= g.C; ++j; }
  }
}

template <int LW>
__global__ void __launch_bounds__(kThreads, 2)
ln2d_patch_fwd_kernel(const float* __restrict__ x, const float* __restrict__ lnw, const float* __restrict__ lnb, float eps,
                      __nv_bfloat16* __restrict__ A, float* __restrict__ mu, float* __restrict__ rstd, G2 g) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* red = reinterpret_cast<float*>(smem);                // 4096 floats
  float* red2 = red + 4096;                                   // [256]
  float* cst = red2 + kThreads;                               // [2][PIX]  r, -mean*r
  float* lnw_s = cst + 2 * g.PIX;                             // [C]
  float* lnb_s = lnw_s + g.C;                                 // [C]
  __nv_bfloat16* gs = reinterpret_cast<__nv_bfloat16*>(lnb_s + g.C);   // [PIX][ge]
  const int tid = threadIdx.x, C = g.C, HW = g.HW;
  const int sp = tid & ((1 << g.vshift) - 1), j0 = sp * kSpan;
  int ci[kItems];
#pragma unroll
  for (int i = 0; i < kItems; ++i) ci[i] = (tid + i * kThreads) >> g.vshift;
  for (int c = tid; c < C; c += kThreads) { lnw_s[c] = lnw[c]; lnb_s[c] = lnb[c]; }
  const RowWalk rw = row_walk(C);
  f2 xr[kItems][4];
  auto load_tile = [&](int t) {
    const TileIdx ti = tile_of(g, t);
    const int nv = ti.npix - j0;
#pragma unroll
    for (int i = 0; i < kItems; ++i)
      ldg_span_f32<LW>(x + ((size_t)ti.n * C + ci[i]) * HW + ti.p0 + j0, ci[i] < C ? nv : 0, xr[i]);
  };
  int t = blockIdx.x;
  if (t < g.total_tiles) load_tile(t);
  for (; t < g.total_tiles; t += gridDim.x) {
    const TileIdx ti = tile_of(g, t);
    f2 d[kItems][4];
    f2 s[2][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { s[0][k] = 0ull; s[1][k] = 0ull; }
#pragma unroll
    for (int i = 0; i < kItems; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        d[i][k] = xr[i][k];                                   // zeros for channels beyond C
        s[0][k] = add2(s[0][k], d[i][k]); s[1][k] = fma2(d[i][k], d[i][k], s[1][k]);
      }
    if (t + gridDim.x < g.total_tiles) load_tile(t + gridDim.x);
    pix_partials_store<2>(red, g, sp, s);
    __syncthreads();
    pix_partials_fold<2>(red, red2, g);
    __syncthreads();
    if (tid < g.PIX) {
      const float m = pix_total<2>(red2, g, 0, tid) * g.invC;
      const float var = fmaxf(pix_total<2>(red2, g, 1, tid) * g.invC - m * m, 0.f);
      const float r = rsqrtf(var + eps);
      cst[tid] = r; cst[g.PIX + tid] = -m * r;
      if (tid < ti.npix) {
        const size_t pix = (size_t)ti.n * HW + ti.p0 + tid;
        mu[pix] = m; rstd[pix] = r;
      }
    }
    __syncthreads();
    {
      f2 r2[4], nm2[4];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(cst + j0 + 4 * k);
        const float4 b = *reinterpret_cast<const float4*>(cst + g.PIX + j0 + 4 * k);
        r2[2 * k] = mk2(a.x, a.y); r2[2 * k + 1] = mk2(a.z, a.w);
        nm2[2 * k] = mk2(b.x, b.y); nm2[2 * k + 1] = mk2(b.z, b.w);
      }
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        if (ci[i] < C) {
          const f2 w2 = splat(lnw_s[ci[i]]), b2 = splat(lnb_s[ci[i]]);
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k] = pack2(fma2(fma2(d[i][k], r2[k], nm2[k]), w2, b2));
          sts_span_t(gs, g.ge, j0, ci[i], o);
        }
      }
    }
    __syncthreads();
    rows_out_patch(A, gs, g, rw, ti.n, ti.p0, ti.npix * C / 8);
  }
}

// backward: dA (patch rows, bf16) -> dx (NCHW fp32); per-CTA partials [2][C]: dlnw = sum g*xhat, dlnb = sum g
template <int LW>
__global__ void __launch_bounds__(kThreads, 2)
ln2d_patch_bwd_kernel(const __nv_bfloat16* __restrict__ dA, const float* __restrict__ x, const float* __restrict__ lnw,
                      const float* __restrict__ mu, const float* __restrict__ rstd, float* __restrict__ dx,
                      float* __restrict__ part /*[grid][2][C]*/, G2 g) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* red = reinterpret_cast<float*>(smem);                // 4096 floats
  float* red2 = red + 4096;                                   // [256]
  float* cst = red2 + kThreads;                               // [3][PIX] r, alpha, beta
  float* mus = cst + 3 * g.PIX;                               // [PIX]
  float* rs = mus + g.PIX;                                    // [PIX]
  __nv_bfloat16* gs = reinterpret_cast<__nv_bfloat16*>(rs + g.PIX);    // [PIX][ge]
  const int tid = threadIdx.x, C = g.C, HW = g.HW;
  const int spr = 1 << g.vshift, sp = tid & (spr - 1), j0 = sp * kSpan;
  int ci[kItems];
  float wi[kItems];
  f2 A0[kItems], A1[kItems];
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    ci[i] = (tid + i * kThreads) >> g.vshift; wi[i] = ci[i] < C ? lnw[ci[i]] : 0.f;
    A0[i] = 0ull; A1[i] = 0ull;
  }
  const RowWalk rw = row_walk(C);
  const int tvec = g.PIX * C / 8;
  for (int t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
    const TileIdx ti = tile_of(g, t);
    const int nv = ti.npix - j0;
    uint4 gr[kMaxVec];
    f2 d[kItems][4], gg[kItems][4];
    rows_ldg_patch(dA, g, rw, ti.n, ti.p0, ti.npix * C / 8, gr);
#pragma unroll
    for (int i = 0; i < kItems; ++i)
      ldg_span_f32<LW>(x + ((size_t)ti.n * C + ci[i]) * HW + ti.p0 + j0, ci[i] < C ? nv : 0, d[i]);
    float mval = 0.f, rval = 0.f;
    if (tid < ti.npix) { mval = mu[(size_t)ti.n * HW + ti.p0 + tid]; rval = rstd[(size_t)ti.n * HW + ti.p0 + tid]; }
    rows_sts(gs, g, rw, tvec, gr);
    if (tid < g.PIX) { mus[tid] = mval; rs[tid] = rval; }
    __syncthreads();
    {
      f2 nm[4];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(mus + j0 + 4 * k);
        nm[2 * k] = mk2(-a.x, -a.y); nm[2 * k + 1] = mk2(-a.z, -a.w);
      }
      f2 s[2][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { s[0][k] = 0ull; s[1][k] = 0ull; }
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        if (ci[i] < C) {
          lds_span_t(gs, g.ge, j0, ci[i], gg[i]);
          const f2 w2 = splat(wi[i]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            d[i][k] = add2(d[i][k], nm[k]);
            const f2 gw = mul2(gg[i][k], w2);
            s[0][k] = add2(s[0][k], gw);
            s[1][k] = fma2(gw, d[i][k], s[1][k]);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) { d[i][k] = 0ull; gg[i][k] = 0ull; }
        }
      }
      pix_partials_store<2>(red, g, sp, s);
    }
    __syncthreads();
    pix_partials_fold<2>(red, red2, g);
    __syncthreads();
    if (tid < g.PIX) {
      const float r = rs[tid];
      const float m1 = pix_total<2>(red2, g, 0, tid) * g.invC;
      const float m2 = r * pix_total<2>(red2, g, 1, tid) * g.invC;
      cst[tid] = r; cst[g.PIX + tid] = -r * m1; cst[2 * g.PIX + tid] = -r * r * m2;
    }
    __syncthreads();
    {
      f2 r2[4], al[4], be[4];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(cst + j0 + 4 * k);
        const float4 b = *reinterpret_cast<const float4*>(cst + g.PIX + j0 + 4 * k);
        const float4 e = *reinterpret_cast<const float4*>(cst + 2 * g.PIX + j0 + 4 * k);
        r2[2 * k] = mk2(a.x, a.y); r2[2 * k + 1] = mk2(a.z, a.w);
        al[2 * k] = mk2(b.x, b.y); al[2 * k + 1] = mk2(b.z, b.w);
        be[2 * k] = mk2(e.x, e.y); be[2 * k + 1] = mk2(e.z, e.w);
      }
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        if (ci[i] < C && nv > 0) {
          const f2 w2 = splat(wi[i]);
          f2 o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            o[k] = fma2(mul2(r2[k], w2), gg[i][k], fma2(be[k], d[i][k], al[k]));
            A0[i] = fma2(gg[i][k], mul2(d[i][k], r2[k]), A0[i]);
            A1[i] = add2(A1[i], gg[i][k]);
          }
          stg_span_f32<LW>(dx + ((size_t)ti.n * C + ci[i]) * HW + ti.p0 + j0, nv, o);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    float lo, hi;
    un2(A0[i], lo, hi); float a = lo + hi;
    un2(A1[i], lo, hi); float b = lo + hi;
    for (int o = 1; o < spr; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (sp == 0 && ci[i] < C) {
      part[(size_t)blockIdx.x * 2 * C + ci[i]] = a;
      part[(size_t)blockIdx.x * 2 * C + C + ci[i]] = b;
    }
  }
}


// ====================================================================================================================
// Stem (models/SLaK.py:189-193: Conv2d(3, C, k=4, s=4) -> LayerNorm(channels_first)): patch rows of the image, the
// convolution as a GEMM (mlp_tc.cu), then LayerNorm over the channels of each token row with the NCHW residual stream
// (fp32 + bf16 copy) as output.  Backward: LayerNorm backward from the NCHW gradient to token rows (dY, bf16), the
// weight gradient is the split-K GEMM dY^T A; the image needs no gradient.
// ====================================================================================================================
// A[(n, ho, wo)][ci*16 + kh*4 + kw] = x[n, ci, 4*ho + kh, 4*wo + kw] (bf16), columns >= 16*Cin zero up to K = 64.
// One thread per (token, 4-pixel piece): a warp writes 256 contiguous bytes and reads whole 32-byte sectors.
__global__ void __launch_bounds__(256) patchify4_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ A, int N, int Cin,
                                                        int H, int W, size_t total /* tokens * 16 */) {
  const int Ho = H / 4, Wo = W / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int piece = (int)(i & 15);
    const size_t token = i >> 4;
    const int wo = (int)(token % Wo);
    const size_t nh = token / Wo;
    const int ho = (int)(nh % Ho), n = (int)(nh / Ho);
    const int ci = piece >> 2, kh = piece & 3;
    uint2 o = make_uint2(0u, 0u);
    if (ci < Cin) {
      const float4 v = *reinterpret_cast<const float4*>(x + (((size_t)n * Cin + ci) * H + 4 * ho + kh) * W + 4 * wo);
      o.x = tc::pack_bf16(v.x, v.y); o.y = tc::pack_bf16(v.z, v.w);
    }
    *reinterpret_cast<uint2*>(A + token * 64 + piece * 4) = o;
  }
}

template <int LW>
__global__ void __launch_bounds__(kThreads, 2)
ln_rows_fwd_kernel(const __nv_bfloat16* __restrict__ Y, const float* __restrict__ lnw, const float* __restrict__ lnb, float eps,
                   float* __restrict__ out, __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ mu, float* __restrict__ rstd,
                   G2 g) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* red = reinterpret_cast<float*>(smem);
  float* red2 = red + 4096;
  float* cst = red2 + kThreads;                               // [2][PIX]  r, -mean*r
  float* lnw_s = cst + 2 * g.PIX;
  float* lnb_s = lnw_s + g.C;
  __nv_bfloat16* gs = reinterpret_cast<__nv_bfloat16*>(lnb_s + g.C);
  const int tid = threadIdx.x, C = g.C, HW = g.HW;
  const int sp = tid & ((1 << g.vshift) - 1), j0 = sp * kSpan;
  int ci[kItems];
#pragma unroll
  for (int i = 0; i < kItems; ++i) ci[i] = (tid + i * kThreads) >> g.vshift;
  for (int c = tid; c < C; c += kThreads) { lnw_s[c] = lnw[c]; lnb_s[c] = lnb[c]; }
  const RowWalk rw = row_walk(C);
  const int tvec = g.PIX * C / 8;
  uint4 hr[kMaxVec];
  int t = blockIdx.x;
  if (t < g.total_tiles) { const TileIdx ti = tile_of(g, t); rows_ldg(Y + ((size_t)ti.n * HW + ti.p0) * C, ti.npix * C / 8, hr); }
  for (; t < g.total_tiles; t += gridDim.x) {
    const TileIdx ti = tile_of(g, t);
    const int nv = ti.npix - j0;
    __syncthreads();                                         // the previous tile's transposed reads are done
    rows_sts(gs, g, rw, tvec, hr);
    __syncthreads();
    if (t + gridDim.x < g.total_tiles) {
      const TileIdx tn = tile_of(g, t + gridDim.x);
      rows_ldg(Y + ((size_t)tn.n * HW + tn.p0) * C, tn.npix * C / 8, hr);
    }
    f2 d[kItems][4], s[2][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { s[0][k] = 0ull; s[1][k] = 0ull; }
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      if (ci[i] < C) {
        lds_span_t(gs, g.ge, j0, ci[i], d[i]);
#pragma unroll
        for (int k = 0; k < 4; ++k) { s[0][k] = add2(s[0][k], d[i][k]); s[1][k] = fma2(d[i][k], d[i][k], s[1][k]); }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) d[i][k] = 0ull;
      }
    }
    pix_partials_store<2>(red, g, sp, s);
    __syncthreads();
    pix_partials_fold<2>(red, red2, g);
    __syncthreads();
    if (tid < g.PIX) {
      const float m = pix_total<2>(red2, g, 0, tid) * g.invC;
      const float var = fmaxf(pix_total<2>(red2, g, 1, tid) * g.invC - m * m, 0.f);
      const float r = rsqrtf(var + eps);
      cst[tid] = r; cst[g.PIX + tid] = -m * r;
      if (tid < ti.npix) {
        const size_t pix = (size_t)ti.n * HW + ti.p0 + tid;
        mu[pix] = m; rstd[pix] = r;
      }
    }
    __syncthreads();
    f2 r2[4], nm2[4];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(cst + j0 + 4 * k);
      const float4 b = *reinterpret_cast<const float4*>(cst + g.PIX + j0 + 4 * k);
      r2[2 * k] = mk2(a.x, a.y); r2[2 * k + 1] = mk2(a.z, a.w);
      nm2[2 * k] = mk2(b.x, b.y); nm2[2 * k + 1] = mk2(b.z, b.w);
    }
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      if (ci[i] < C && nv > 0) {
        const f2 w2 = splat(lnw_s[ci[i]]), b2 = splat(lnb_s[ci[i]]);
        f2 o[4];
        uint32_t ob[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { o[k] = fma2(fma2(d[i][k], r2[k], nm2[k]), w2, b2); ob[k] = pack2(o[k]); }
        const size_t off = ((size_t)ti.n * C + ci[i]) * HW + ti.p0 + j0;
        stg_span_f32<LW>(out + off, nv, o);
        if (out_bf16) stg_span<LW>(out_bf16 + off, nv, ob);
      }
    }
  }
}

// part per CTA: [3][C] = dlnw, dlnb, column sums of dY (the convolution's bias gradient)
template <int LW>
__global__ void __launch_bounds__(kThreads, 2)
ln_rows_bwd_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ Y, const float* __restrict__ lnw,
                   const float* __restrict__ mu, const float* __restrict__ rstd, __nv_bfloat16* __restrict__ dY,
                   float* __restrict__ part, G2 g) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* red = reinterpret_cast<float*>(smem);
  float* red2 = red + 4096;
  float* cst = red2 + kThreads;                               // [3][PIX] r, alpha, beta
  float* mus = cst + 3 * g.PIX;
  float* rs = mus + g.PIX;
  __nv_bfloat16* gs = reinterpret_cast<__nv_bfloat16*>(rs + g.PIX);
  const int tid = threadIdx.x, C = g.C, HW = g.HW;
  const int spr = 1 << g.vshift, sp = tid & (spr - 1), j0 = sp * kSpan;
  int ci[kItems];
  float wi[kItems], a2[kItems];
  f2 A0[kItems], A1[kItems];
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    ci[i] = (tid + i * kThreads) >> g.vshift; wi[i] = ci[i] < C ? lnw[ci[i]] : 0.f;
    A0[i] = 0ull; A1[i] = 0ull; a2[i] = 0.f;
  }
  const RowWalk rw = row_walk(C);
  const int tvec = g.PIX * C / 8;
  for (int t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
    const TileIdx ti = tile_of(g, t);
    const int nv = ti.npix - j0;
    uint4 hr[kMaxVec];
    f2 gg[kItems][4], d[kItems][4];
    rows_ldg(Y + ((size_t)ti.n * HW + ti.p0) * C, ti.npix * C / 8, hr);
#pragma unroll
    for (int i = 0; i < kItems; ++i)
      ldg_span_f32<LW>(dout + ((size_t)ti.n * C + ci[i]) * HW + ti.p0 + j0, ci[i] < C ? nv : 0, gg[i]);
    float mval = 0.f, rval = 0.f;
    if (tid < ti.npix) { mval = mu[(size_t)ti.n * HW + ti.p0 + tid]; rval = rstd[(size_t)ti.n * HW + ti.p0 + tid]; }
    __syncthreads();                                         // the previous tile's rows have been copied out
    rows_sts(gs, g, rw, tvec, hr);
    if (tid < g.PIX) { mus[tid] = mval; rs[tid] = rval; }
    __syncthreads();
    {
      f2 nm[4];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(mus + j0 + 4 * k);
        nm[2 * k] = mk2(-a.x, -a.y); nm[2 * k + 1] = mk2(-a.z, -a.w);
      }
      f2 s[2][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { s[0][k] = 0ull; s[1][k] = 0ull; }
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        if (ci[i] < C) {
          lds_span_t(gs, g.ge, j0, ci[i], d[i]);
          const f2 w2 = splat(wi[i]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            d[i][k] = add2(d[i][k], nm[k]);
            const f2 gw = mul2(gg[i][k], w2);
            s[0][k] = add2(s[0][k], gw);
            s[1][k] = fma2(gw, d[i][k], s[1][k]);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) d[i][k] = 0ull;
        }
      }
      pix_partials_store<2>(red, g, sp, s);
    }
    __syncthreads();
    pix_partials_fold<2>(red, red2, g);
    __syncthreads();
    if (tid < g.PIX) {
      const float r = rs[tid];
      const float m1 = pix_total<2>(red2, g, 0, tid) * g.invC;
      const float m2 = r * pix_total<2>(red2, g, 1, tid) * g.invC;
      cst[tid] = r; cst[g.PIX + tid] = -r * m1; cst[2 * g.PIX + tid] = -r * r * m2;
    }
    __syncthreads();
    {
      f2 r2[4], al[4], be[4];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(cst + j0 + 4 * k);
        const float4 b = *reinterpret_cast<const float4*>(cst + g.PIX + j0 + 4 * k);
        const float4 e = *reinterpret_cast<const float4*>(cst + 2 * g.PIX + j0 + 4 * k);
        r2[2 * k] = mk2(a.x, a.y); r2[2 * k + 1] = mk2(a.z, a.w);
        al[2 * k] = mk2(b.x, b.y); al[2 * k + 1] = mk2(b.z, b.w);
        be[2 * k] = mk2(e.x, e.y); be[2 * k + 1] = mk2(e.z, e.w);
      }
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        if (ci[i] < C) {                                       // pixels beyond the tile: r = alpha = beta = 0 and g = 0 -> zeros
          const f2 w2 = splat(wi[i]);
          uint32_t ob[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            ob[k] = pack2(fma2(mul2(r2[k], w2), gg[i][k], fma2(be[k], d[i][k], al[k])));
            A0[i] = fma2(gg[i][k], mul2(d[i][k], r2[k]), A0[i]);
            A1[i] = add2(A1[i], gg[i][k]);
            fh_add(a2[i], ob[k]);
          }
          sts_span_t(gs, g.ge, j0, ci[i], ob);                 // in place: this thread owns these 8 elements
        }
      }
    }
    __syncthreads();
    rows_out(dY + ((size_t)ti.n * HW + ti.p0) * C, gs, g, rw, ti.npix * C / 8);
  }
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    float lo, hi;
    un2(A0[i], lo, hi); float a = lo + hi;
    un2(A1[i], lo, hi); float b = lo + hi;
    float c3 = a2[i];
    for (int o = 1; o < spr; o <<= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c3 += __shfl_xor_sync(0xffffffffu, c3, o);
    }
    if (sp == 0 && ci[i] < C) {
      part[(size_t)blockIdx.x * 3 * C + ci[i]] = a;
      part[(size_t)blockIdx.x * 3 * C + C + ci[i]] = b;
      part[(size_t)blockIdx.x * 3 * C + 2 * C + ci[i]] = c3;
    }
  }
}

// ====================================================================================================================
// host side
// ====================================================================================================================
static bool enabled() {                    // read on every call (host side, nanoseconds): tests and tools switch it at run time
  const char* e = getenv("SLAK_GLUE_V1");
  return !(e && atoi(e) != 0);
}
static bool make_g2(int N, int C, int HW, G2* g) {
  if (!enabled() || C % 8 != 0 || C < 8 || HW < 1) return false;
  int pix = kTileElems / C;
  if (pix < 8) return false;                                  // C > 768
  pix = pix >= 128 ? 128 : (pix >= 64 ? 64 : (pix >= 32 ? 32 : (pix >= 16 ? 16 : 8)));
  while (pix > 8 && pix / 2 >= HW) pix /= 2;
  g->N = N; g->C = C; g->HW = HW; g->PIX = pix;
  int vs = 0;
  while ((8 << vs) < pix) ++vs;
  g->vshift = vs;
  g->tiles_per_img = (HW + pix - 1) / pix;
  g->total_tiles = g->tiles_per_img * N;
  g->ge = row_pitch(C);
  g->invC = 1.f / (float)C;
  g->W = g->Wo = g->Ho = 0; g->magicW = 0;
  return true;
}
static bool make_g2_patch(int N, int C, int H, int W, G2* g) {
  if (H < 2 || W < 2 || (H & 1) || (W & 1) || W >= 1024 || (long long)H * W >= (1 << 20)) return false;
  if (!make_g2(N, C, H * W, g)) return false;
  g->W = W; g->Wo = W / 2; g->Ho = H / 2;
  g->magicW = (unsigned)((1ull << 32) / (unsigned)W) + 1u;
  return true;
}
bool supported(int N, int C, int HW) { G2 g; return make_g2(N, C, HW, &g); }
static int lw_of(int HW, uintptr_t bf16_ptrs, uintptr_t f32_ptrs) {
  if (HW % 8 == 0 && (bf16_ptrs & 15) == 0 && (f32_ptrs & 15) == 0) return 8;
  if (HW % 4 == 0 && (bf16_ptrs & 7) == 0 && (f32_ptrs & 15) == 0) return 4;
  return 1;
}
static int grid_of(const G2& g, int per_sm) {
  int grid = sm_count() * per_sm;
  if (grid > g.total_tiles) grid = g.total_tiles;
  return grid < 1 ? 1 : grid;
}
#define SLAK_LW_DISPATCH(lw, CALL)  \
  do {                              \
    if ((lw) == 8) { CALL(8); }     \
    else if ((lw) == 4) { CALL(4); }\
    else { CALL(1); }               \
  } while (0)

static size_t gs_bytes(const G2& g) { return (size_t)g.PIX * g.ge * sizeof(__nv_bfloat16); }

int ln_fwd(const void* y1, const void* y2, const void* y3, const float* scale, const float* shift, const float* lnw,
           const float* lnb, float eps, void* xn, float* mu, float* rstd, int N, int C, int HW, cudaStream_t st) {
  G2 g;
  if (!make_g2(N, C, HW, &g) || (reinterpret_cast<uintptr_t>(xn) & 15) != 0) return SLAK_G2_UNSUPPORTED;
  // (the 2-byte-aligned span loads read whole 8-byte words: the planes' tensors must start on 8 bytes)
  if (((reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(y2) | reinterpret_cast<uintptr_t>(y3)) & 7) != 0) return SLAK_G2_UNSUPPORTED;
  const size_t smem = (4096 + kThreads + 2 * (size_t)g.PIX) * sizeof(float) + (size_t)C * (sizeof(float4) + 2 * sizeof(float)) + gs_bytes(g);
  const int lw = lw_of(HW, (uintptr_t)y1 | (uintptr_t)y2 | (uintptr_t)y3, 0);
  const int grid = grid_of(g, 2);
#define CALL(V)                                                                                                      \
  SLAK_SET_MAX_SMEM(ln_fwd2_kernel<V>, smem);                                                                       \
  ln_fwd2_kernel<V><<<grid, kThreads, smem, st>>>((const __nv_bfloat16*)y1, (const __nv_bfloat16*)y2, (const __nv_bfloat16*)y3, \
      scale, shift, lnw, lnb, eps, (__nv_bfloat16*)xn, mu, rstd, g)
  SLAK_LW_DISPATCH(lw, CALL);
#undef CALL
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

int res_fwd(const float* x, const void* h2, const float* gamma, const float* dp, float* out, void* out_bf16, int N, int C,
            int HW, cudaStream_t st) {
  G2 g;
  if (!make_g2(N, C, HW, &g) || (reinterpret_cast<uintptr_t>(h2) & 15) != 0) return SLAK_G2_UNSUPPORTED;
  const size_t smem = gs_bytes(g);
  const int lw = lw_of(HW, (uintptr_t)out_bf16, (uintptr_t)x | (uintptr_t)out);
  static const int flat_lw = [] { const char* e = getenv("SLAK_RES_FLAT_LW"); return e ? atoi(e) : 4; }();   // planes of this alignment class and below go flat (14 x 14: 34 -> 32 us, 31 -> 24 us as a pure transposition)
  if (lw <= flat_lw) {
    // planes that are not 16-byte aligned (7 x 7, 14 x 14): tiles of CH channels x the whole (contiguous) plane instead of
    // spans: every NCHW access is a coalesced element per lane
    const int CHmax = lw == 1 ? 64 : 32;
    const int CH = (C % 64 == 0 && CHmax >= 64) ? 64 : (C % 32 == 0 ? 32 : (C % 16 == 0 ? 16 : 8));
    const size_t fsm = (((size_t)HW * (CH + 2) * 2 + 15) & ~(size_t)15) + (size_t)CH * sizeof(float);
    if (fsm <= 48 * 1024) {
      int fgrid = N * (C / CH);
      if (fgrid > 8 * sm_count()) fgrid = 8 * sm_count();
      res_fwd_flat_kernel<<<fgrid, kThreads, fsm, st>>>(x, (const __nv_bfloat16*)h2, gamma, dp, out, (__nv_bfloat16*)out_bf16, N, C, HW, CH);
      SLAK_CUDA_TRY(cudaGetLastError());
      return SLAK_OK;
    }
    if (x != nullptr) return SLAK_G2_UNSUPPORTED;           // (the first-generation kernel takes it; the pure transposition goes on below)
  }
  const int grid = grid_of(g, 3);
#define CALL(V)                                                                                                      \
  SLAK_SET_MAX_SMEM(res_fwd2_kernel<V>, smem);                                                                      \
  res_fwd2_kernel<V><<<grid, kThreads, smem, st>>>(x, (const __nv_bfloat16*)h2, gamma, dp, out, (__nv_bfloat16*)out_bf16, g)
  SLAK_LW_DISPATCH(lw, CALL);
#undef CALL
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

int res_bwd_parts(int N, int C, int HW) {
  G2 g;
  return make_g2(N, C, HW, &g) ? grid_of(g, 3) : 0;
}
int res_bwd(const float* dout, const void* h2, const float* gamma, const float* dp, void* dh2, float* part, int N, int C,
            int HW, cudaStream_t st) {
  G2 g;
  if (!make_g2(N, C, HW, &g) || ((reinterpret_cast<uintptr_t>(h2) | reinterpret_cast<uintptr_t>(dh2)) & 15) != 0)
    return SLAK_G2_UNSUPPORTED;      // (a null h2 is the pure-transposition mode)
  const size_t smem = gs_bytes(g);
  const int lw = lw_of(HW, 0, (uintptr_t)dout);
  const int grid = grid_of(g, 3);
#define CALL(V)                                                                                                      \
  SLAK_SET_MAX_SMEM(res_bwd2_kernel<V>, smem);                                                                      \
  res_bwd2_kernel<V><<<grid, kThreads, smem, st>>>(dout, (const __nv_bfloat16*)h2, gamma, dp, (__nv_bfloat16*)dh2, part, g)
  SLAK_LW_DISPATCH(lw, CALL);
#undef CALL
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

int ln_bwd_parts(int N, int C, int HW) {
  G2 g;
  return make_g2(N, C, HW, &g) ? grid_of(g, 2) : 0;
}
int ln_bwd(const void* dxn, const void* y1, const void* y2, const void* y3, const float* scale, const float* shift,
           const float* lnw, const float* mu, const float* rstd, void* du, float* part, int N, int C, int HW, cudaStream_t st) {
  G2 g;
  if (!make_g2(N, C, HW, &g) || (reinterpret_cast<uintptr_t>(dxn) & 15) != 0) return SLAK_G2_UNSUPPORTED;
  if (((reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(y2) | reinterpret_cast<uintptr_t>(y3)) & 7) != 0) return SLAK_G2_UNSUPPORTED;
  const size_t smem = (size_t)3 * kItems * kThreads * sizeof(uint4) + (4096 + kThreads + 5 * (size_t)g.PIX) * sizeof(float) +
                      (size_t)C * sizeof(float4) + gs_bytes(g);
  const int lw = lw_of(HW, (uintptr_t)y1 | (uintptr_t)y2 | (uintptr_t)y3 | (uintptr_t)du, 0);
  const int grid = grid_of(g, 2);
#define CALL(V)                                                                                                      \
  SLAK_SET_MAX_SMEM(ln_bwd2_kernel<V>, smem);                                                                       \
  ln_bwd2_kernel<V><<<grid, kThreads, smem, st>>>((const __nv_bfloat16*)dxn, (const __nv_bfloat16*)y1, (const __nv_bfloat16*)y2, \
      (const __nv_bfloat16*)y3, scale, shift, lnw, mu, rstd, (__nv_bfloat16*)du, part, g)
  SLAK_LW_DISPATCH(lw, CALL);
#undef CALL
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}


// ---- downsampling layer: LayerNorm -> patch rows, and its backward ---------------------------------------------------
int ln2d_patch_fwd(const float* x, const float* lnw, const float* lnb, float eps, void* A, float* mu, float* rstd, int N, int C,
                   int H, int W, cudaStream_t st) {
  G2 g;
  SLAK_REQUIRE(make_g2_patch(N, C, H, W, &g) && (reinterpret_cast<uintptr_t>(A) & 15) == 0, SLAK_ERR_UNSUPPORTED,
               "ln2d_patch_fwd: C=%d (multiple of 8, <= 768), H=%d, W=%d (even) unsupported", C, H, W);
  const size_t smem = (4096 + kThreads + 2 * (size_t)g.PIX + 2 * (size_t)C) * sizeof(float) + gs_bytes(g);
  const int lw = lw_of(H * W, 0, (uintptr_t)x);
  const int grid = grid_of(g, 2);
#define CALL(V)                                                                                                      \
  SLAK_SET_MAX_SMEM(ln2d_patch_fwd_kernel<V>, smem);                                                                \
  ln2d_patch_fwd_kernel<V><<<grid, kThreads, smem, st>>>(x, lnw, lnb, eps, (__nv_bfloat16*)A, mu, rstd, g)
  SLAK_LW_DISPATCH(lw, CALL);
#undef CALL
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}
int ln2d_patch_bwd_parts(int N, int C, int H, int W) {
  G2 g;
  return make_g2_patch(N, C, H, W, &g) ? grid_of(g, 2) : 0;
}
int ln2d_patch_bwd(const void* dA, const float* x, const float* lnw, const float* mu, const float* rstd, float* dx, float* part,
                   int N, int C, int H, int W, cudaStream_t st) {
  G2 g;
  SLAK_REQUIRE(make_g2_patch(N, C, H, W, &g) && (reinterpret_cast<uintptr_t>(dA) & 15) == 0, SLAK_ERR_UNSUPPORTED,
               "ln2d_patch_bwd: C=%d (multiple of 8, <= 768), H=%d, W=%d (even) unsupported", C, H, W);
  const size_t smem = (4096 + kThreads + 5 * (size_t)g.PIX) * sizeof(float) + gs_bytes(g);
  const int lw = lw_of(H * W, 0, (uintptr_t)x | (uintptr_t)dx);
  const int grid = grid_of(g, 2);
#define CALL(V)                                                                                                      \
  SLAK_SET_MAX_SMEM(ln2d_patch_bwd_kernel<V>, smem);                                                                \
  ln2d_patch_bwd_kernel<V><<<grid, kThreads, smem, st>>>((const __nv_bfloat16*)dA, x, lnw, mu, rstd, dx, part, g)
  SLAK_LW_DISPATCH(lw, CALL);
#undef CALL
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}
// pure layout changes between the token-major (NHWC, bf16) side of a GEMM and the NCHW residual stream
int nhwc_to_nchw(const void* h, float* out, void* out_bf16, int N, int C, int HW, cudaStream_t st) {
  const int rc = res_fwd(nullptr, h, nullptr, nullptr, out, out_bf16, N, C, HW, st);
  SLAK_REQUIRE(rc != SLAK_G2_UNSUPPORTED, SLAK_ERR_UNSUPPORTED, "nhwc_to_nchw: C=%d must be a multiple of 8 and <= 768", C);
  return rc;
}
int nchw_to_nhwc_parts(int N, int C, int HW) { return res_bwd_parts(N, C, HW); }
int nchw_to_nhwc(const float* src, void* dst_bf16, float* part, int N, int C, int HW, cudaStream_t st) {
  const int rc = res_bwd(src, nullptr, nullptr, nullptr, dst_bf16, part, N, C, HW, st);
  SLAK_REQUIRE(rc != SLAK_G2_UNSUPPORTED, SLAK_ERR_UNSUPPORTED, "nchw_to_nhwc: C=%d must be a multiple of 8 and <= 768", C);
  return rc;
}


// ---- stem -------------------------------------------------------------------------------------------------------------
int patchify4(const float* x, void* A, int N, int Cin, int H, int W, cudaStream_t st) {
  SLAK_REQUIRE(Cin >= 1 && Cin <= 4 && H % 4 == 0 && W % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(A) & 15) == 0, SLAK_ERR_UNSUPPORTED,
               "patchify4: Cin=%d (<= 4), H=%d, W=%d (multiples of 4), 16-byte aligned tensors", Cin, H, W);
  const size_t total = (size_t)N * (H / 4) * (W / 4) * 16;
  size_t grid = (total + 255) / 256;
  if (grid > (size_t)sm_count() * 16) grid = (size_t)sm_count() * 16;
  patchify4_kernel<<<(int)grid, 256, 0, st>>>(x, (__nv_bfloat16*)A, N, Cin, H, W, total);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}
int ln_rows_fwd(const void* Y, const float* lnw, const float* lnb, float eps, float* out, void* out_bf16, float* mu, float* rstd,
                int N, int C, int HW, cudaStream_t st) {
  G2 g;
  SLAK_REQUIRE(make_g2(N, C, HW, &g) && (reinterpret_cast<uintptr_t>(Y) & 15) == 0, SLAK_ERR_UNSUPPORTED,
               "ln_rows_fwd: C=%d must be a multiple of 8 and <= 768", C);
  const size_t smem = (4096 + kThreads + 2 * (size_t)g.PIX + 2 * (size_t)C) * sizeof(float) + gs_bytes(g);
  const int lw = lw_of(HW, (uintptr_t)out_bf16, (uintptr_t)out);
  const int grid = grid_of(g, 2);
#define CALL(V)                                                                                                      \
  SLAK_SET_MAX_SMEM(ln_rows_fwd_kernel<V>, smem);                                                                   \
  ln_rows_fwd_kernel<V><<<grid, kThreads, smem, st>>>((const __nv_bfloat16*)Y, lnw, lnb, eps, out, (__nv_bfloat16*)out_bf16, mu, rstd, g)
  SLAK_LW_DISPATCH(lw, CALL);
#undef CALL
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}
int ln_rows_bwd_parts(int N, int C, int HW) {
  G2 g;
  return make_g2(N, C, HW, &g) ? grid_of(g, 2) : 0;
}
int ln_rows_bwd(const float* dout, const void* Y, const float* lnw, const float* mu, const float* rstd, void* dY, float* part,
                int N, int C, int HW, cudaStream_t st) {
  G2 g;
  SLAK_REQUIRE(make_g2(N, C, HW, &g) && ((reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(dY)) & 15) == 0,
               SLAK_ERR_UNSUPPORTED, "ln_rows_bwd: C=%d must be a multiple of 8 and <= 768", C);
  const size_t smem = (4096 + kThreads + 5 * (size_t)g.PIX) * sizeof(float) + gs_bytes(g);
  const int lw = lw_of(HW, 0, (uintptr_t)dout);
  const int grid = grid_of(g, 2);
#define CALL(V)                                                                                                      \
  SLAK_SET_MAX_SMEM(ln_rows_bwd_kernel<V>, smem);                                                                   \
  ln_rows_bwd_kernel<V><<<grid, kThreads, smem, st>>>(dout, (const __nv_bfloat16*)Y, lnw, mu, rstd, (__nv_bfloat16*)dY, part, g)
  SLAK_LW_DISPATCH(lw, CALL);
#undef CALL
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

}  // namespace g2
}  // namespace blk
}  // namespace slak
