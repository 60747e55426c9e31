// Depthwise conv2d (stride 1, same padding) on CUDA cores, fp32 accumulate.
//
// Two families of kernels, all NCHW and channel-resident (one CTA = one channel
// c and a strided set of images n, so the taps of c sit in shared memory once):
//
//  * "fast" kernels for K x KS / KS x K filters with a short side KS in {3,5,7}
//    (SLaK's 51x5, 5x51, 5x5 ...).  A whole (n,c) plane lives zero-padded in
//    shared memory.  Lanes of a warp lie along the SHORT-kernel axis, each thread
//    slides a register window along the LONG-kernel axis and keeps KS partial
//    sums per output (one per short-axis tap); the KS partial sums are combined
//    across neighbouring lanes with warp shuffles at the end.  One x value read
//    from shared memory feeds 7*KS FMAs; taps are read as broadcast float4.
//    dgrad is the same kernel with the taps flipped.  wgrad keeps a 13 x KS tile
//    of tap accumulators per thread across all the planes a CTA visits and
//    reduces lanes -> warps -> CTAs in a fixed order (deterministic; the
//    reference scatters with atomicAdd, dwconv2d_direct_epilogue_simt.h:160-185).
//
//  * "generic" kernels for every other odd kh x kw (square 13x13, 31x31, 51x51
//    of the non-Decom layout, models/SLaK.py:82-83): tiled direct convolution.
//
// Semantics follow cutlass/examples/19_large_depthwise_conv2d_torch_extension/
// forward_fp32.cu:135-144,227 (cross-correlation, pad = k/2, stride 1).
#include "common.cuh"

namespace slak {

constexpr int TP = 7;         // outputs per thread along the slide axis
constexpr int UN = 4;         // taps (fwd) / positions (wgrad) per unrolled chunk
constexpr int PADT = 10;      // zero border along the slide axis (fwd/dgrad)
constexpr int PADT_WG = 16;   // zero border along the slide axis (wgrad)
constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;
constexpr int RB = 13;        // long-axis taps per wgrad accumulator tile

struct FastGeom {
  int H, W, KL;
  int SL, TL;          // lane-axis / slide-axis extent of a plane
  int rows, pitch;     // padded smem plane
  int roff, coff;      // position of element (0,0) in the padded plane
  int ppw, nwin, nseg; // planes per warp window, windows per slot row, segments / tap blocks
  int G;               // planes per CTA iteration
  int items;           // warp work items per iteration
  int opitch;          // fp32 output staging pitch
  int wfloats;         // floats reserved for the taps
  int zchunks;         // wgrad: item chunks spread over gridDim.z
};

// ---------------------------------------------------------------------------------
// cooperative plane load: global [H][W] of T -> zero-bordered fp32 smem plane
// ---------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void load_plane(const T* __restrict__ src, float* __restrict__ dst,
                                           int H, int W, int pitch, int roff, int coff,
                                           int tid, int nthreads) {
  const int HW = H * W;
  constexpr int V = 16 / sizeof(T);
  const bool vec = ((HW % V) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (vec) {
    const int nv = HW / V;
    const int4* s4 = reinterpret_cast<const int4*>(src);
    for (int v = tid; v < nv; v += nthreads) {
      int4 raw = __ldg(s4 + v);
      const T* e = reinterpret_cast<const T*>(&raw);
      int idx = v * V;
      int r = idx / W, c = idx - r * W;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        dst[(r + roff) * pitch + coff + c] = to_f32<T>(e[k]);
        if (++c == W) { c = 0; ++r; }
      }
    }
  } else {
    for (int idx = tid; idx < HW; idx += nthreads) {
      int r = idx / W, c = idx - r * W;
      dst[(r + roff) * pitch + coff + c] = to_f32<T>(src[idx]);
    }
  }
}

// fp32 staging [H][opitch] -> global [H][W] of T
template <typename T>
__device__ __forceinline__ void store_plane(const float* __restrict__ src, T* __restrict__ dst,
                                            int H, int W, int opitch, int tid, int nthreads) {
  const int HW = H * W;
  constexpr int V = 16 / sizeof(T);
  const bool vec = ((HW % V) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  if (vec) {
    const int nv = HW / V;
    int4* d4 = reinterpret_cast<int4*>(dst);
    for (int v = tid; v < nv; v += nthreads) {
      int4 raw;
      T* e = reinterpret_cast<T*>(&raw);
      int idx = v * V;
      int r = idx / W, c = idx - r * W;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        e[k] = from_f32<T>(src[r * opitch + c]);
        if (++c == W) { c = 0; ++r; }
      }
      d4[v] = raw;
    }
  } else {
    for (int idx = tid; idx < HW; idx += nthreads) {
      int r = idx / W, c = idx - r * W;
      dst[idx] = from_f32<T>(src[r * opitch + c]);
    }
  }
}

// ---------------------------------------------------------------------------------
// fast forward / dgrad
// ---------------------------------------------------------------------------------
template <typename T, typename WT, int KS, bool VERT>
__global__ void __launch_bounds__(NTHREADS, 2)
dw_fwd_fast_kernel(const T* __restrict__ x, const WT* __restrict__ w, T* __restrict__ y,
                   int N, int C, FastGeom g, int flip, int n_per_cta) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PL = KS / 2;
  float* wsm = smem;                                   // [KLpad][KS]
  float* xin = wsm + g.wfloats;                        // [G][rows][pitch]
  float* yout = xin + g.G * g.rows * g.pitch;          // [G][H][opitch]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x;
  const int n_begin = blockIdx.y * n_per_cta;
  const int n_end = min(N, n_begin + n_per_cta);
  const int HW = g.H * g.W;
  const int plane_sz = g.rows * g.pitch;
  const int KL = g.KL, pad = KL / 2;

  // taps -> smem as [long tap r][short tap s], rounded to T, flipped for dgrad
  for (int i = tid; i < g.wfloats; i += NTHREADS) {
    float v = 0.f;
    int r = i / KS, s = i - r * KS;
    if (r < KL) {
      int rr = flip ? (KL - 1 - r) : r, ss = flip ? (KS - 1 - s) : s;
      int gi = VERT ? (rr * KS + ss) : (ss * KL + rr);
      v = round_to<T>(to_f32<WT>(w[(size_t)c * KL * KS + gi]));
    }
    wsm[i] = v;
  }
  for (int i = tid; i < g.G * plane_sz; i += NTHREADS) xin[i] = 0.f;
  __syncthreads();

  const int strideL = VERT ? 1 : g.pitch;
  const int strideT = VERT ? g.pitch : 1;

  for (int n0 = n_begin; n0 < n_end; n0 += g.G) {
    const int gcur = min(g.G, n_end - n0);
    for (int gi = 0; gi < gcur; ++gi)
      load_plane<T>(x + ((size_t)(n0 + gi) * C + c) * HW, xin + gi * plane_sz, g.H, g.W, g.pitch,
                    g.roff, g.coff, tid, NTHREADS);
    __syncthreads();

    for (int item = warp; item < g.items; item += NWARPS) {
      const int per_pg = g.nwin * g.nseg;
      const int pg = item / per_pg;
      const int rem = item - pg * per_pg;
      const int win = rem / g.nseg, seg = rem - win * g.nseg;
      int plane, l;
      bool out_ok;
      if (g.ppw > 1) {
        const int slotw = g.SL + PL;
        const int sub = lane / slotw;
        l = lane - sub * slotw;
        plane = pg * g.ppw + sub;
        const bool pv = (sub < g.ppw) && (plane < gcur);
        out_ok = pv && (l < g.SL);
        if (!pv) { plane = 0; l = g.SL; }
        if (l > g.SL) l = g.SL;  // zero gap column
      } else {
        plane = pg;
        l = win * (32 - 2 * PL) + lane - PL;
        out_ok = (plane < gcur) && (lane >= PL) && (lane < 32 - PL) && (l < g.SL);
        if (plane >= gcur) plane = 0;
        if (l > g.SL + PL - 1) l = g.SL + PL - 1;
      }
      const int t0 = seg * TP;
      const float* xp = xin + plane * plane_sz + g.roff * g.pitch + g.coff + l * strideL;

      float acc[KS][TP];
#pragma unroll
      for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int i = 0; i < TP; ++i) acc[s][i] = 0.f;

      const int r_lo = max(0, pad - t0 - (TP - 1));
      const int r_hi = min(KL - 1, pad + g.TL - 1 - t0);
      int r = r_lo & ~(UN - 1);
      int tb = t0 + r - pad;
      float xw[TP + UN - 1];
#pragma unroll
      for (int i = 0; i < TP - 1; ++i) xw[i] = xp[(tb + i) * strideT];
      for (; r <= r_hi; r += UN) {
#pragma unroll
        for (int u = 0; u < UN; ++u) xw[TP - 1 + u] = xp[(tb + TP - 1 + u) * strideT];
        float wr[UN * KS];
        const float4* wv = reinterpret_cast<const float4*>(wsm + r * KS);
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          float4 q = wv[k];
          wr[4 * k + 0] = q.x; wr[4 * k + 1] = q.y; wr[4 * k + 2] = q.z; wr[4 * k + 3] = q.w;
        }
#pragma unroll
        for (int u = 0; u < UN; ++u)
#pragma unroll
          for (int s = 0; s < KS; ++s)
#pragma unroll
            for (int i = 0; i < TP; ++i) acc[s][i] = fmaf(xw[i + u], wr[u * KS + s], acc[s][i]);
#pragma unroll
        for (int i = 0; i < TP - 1; ++i) xw[i] = xw[i + UN];
        tb += UN;
      }

      // combine the KS partial sums across neighbouring lanes
      float* yo = yout + plane * (g.H * g.opitch);
#pragma unroll
      for (int i = 0; i < TP; ++i) {
        float v = acc[PL][i];
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          if (s == PL) continue;
          const int srcl = lane + s - PL;
          float o = __shfl_sync(0xffffffffu, acc[s][i], srcl & 31);
          v += (srcl >= 0 && srcl < 32) ? o : 0.f;
        }
        const int t = t0 + i;
        if (out_ok && t < g.TL) {
          if (VERT) yo[t * g.opitch + l] = v;
          else      yo[l * g.opitch + t] = v;
        }
      }
    }
    __syncthreads();
    for (int gi = 0; gi < gcur; ++gi)
      store_plane<T>(yout + gi * (g.H * g.opitch), y + ((size_t)(n0 + gi) * C + c) * HW, g.H, g.W,
                     g.opitch, tid, NTHREADS);
    // the next iteration's load only touches xin; its barrier orders the staging reuse
  }
}

// ---------------------------------------------------------------------------------
// fast wgrad: partial[part][c][kh*kw]
// ---------------------------------------------------------------------------------
template <typename T, int KS, bool VERT>
__global__ void __launch_bounds__(NTHREADS, 2)
dw_wgrad_fast_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ partial,
                     int N, int C, FastGeom g, int n_per_cta) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PL = KS / 2;
  const int plane_sz = g.rows * g.pitch;
  float* xin = smem;                         // [G][rows][pitch]
  float* din = xin + g.G * plane_sz;         // [G][rows][pitch]
  float* red = din + g.G * plane_sz;         // [NWARPS][RB*KS]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x;
  const int part = blockIdx.y;
  const int n_begin = part * n_per_cta;
  const int n_end = min(N, n_begin + n_per_cta);
  const int HW = g.H * g.W;
  const int KL = g.KL, pad = KL / 2;
  const int strideL = VERT ? 1 : g.pitch;
  const int strideT = VERT ? g.pitch : 1;

  for (int i = tid; i < 2 * g.G * plane_sz; i += NTHREADS) smem[i] = 0.f;

  // item of this warp (fixed for the whole kernel so accumulators stay in registers)
  const int per_pg = g.nwin * g.nseg;          // nseg = number of tap blocks here
  const int item = blockIdx.z * NWARPS + warp; // global item id within an iteration
  const int pg = item / per_pg;
  const int rem = item - pg * per_pg;
  const int win = rem / g.nseg, blk = rem - win * g.nseg;
  const bool item_ok = item < g.items;
  const int r0 = blk * RB;

  int sub = 0, l;
  bool lane_real;
  if (g.ppw > 1) {
    sub = lane / g.SL;
    l = lane - sub * g.SL;
    lane_real = sub < g.ppw;
  } else {
    l = win * 32 + lane;
    lane_real = l < g.SL;
  }
  if (!lane_real) { sub = 0; l = g.SL; }  // points at a zero border line

  float acc[RB][KS];
#pragma unroll
  for (int j = 0; j < RB; ++j)
#pragma unroll
    for (int s = 0; s < KS; ++s) acc[j][s] = 0.f;

  const int p_lo = max(0, pad - r0 - (RB - 1));
  const int p_hi = min(g.TL - 1, g.TL - 1 + pad - r0);

  __syncthreads();
  for (int n0 = n_begin; n0 < n_end; n0 += g.G) {
    const int gcur = min(g.G, n_end - n0);
    for (int gi = 0; gi < g.G; ++gi) {
      if (gi < gcur) {
        load_plane<T>(x + ((size_t)(n0 + gi) * C + c) * HW, xin + gi * plane_sz, g.H, g.W, g.pitch,
                      g.roff, g.coff, tid, NTHREADS);
        load_plane<T>(dy + ((size_t)(n0 + gi) * C + c) * HW, din + gi * plane_sz, g.H, g.W, g.pitch,
                      g.roff, g.coff, tid, NTHREADS);
      } else {
        // tail: planes beyond the batch contribute nothing -> zero dy interior
        for (int idx = tid; idx < HW; idx += NTHREADS) {
          int r = idx / g.W, cc = idx - r * g.W;
          din[gi * plane_sz + (r + g.roff) * g.pitch + g.coff + cc] = 0.f;
        }
      }
    }
    __syncthreads();
    if (item_ok) {
      const int plane = pg * g.ppw + sub;
      const float* xp = xin + plane * plane_sz + g.roff * g.pitch + g.coff + l * strideL;
      const float* dp = din + plane * plane_sz + g.roff * g.pitch + g.coff;
      // dy line for short tap s sits at lane-axis coordinate l + PL - s
      float xw[RB + UN - 1];
      int p = p_lo;
      int tb = p + r0 - pad;
#pragma unroll
      for (int j = 0; j < RB - 1; ++j) xw[j] = xp[(tb + j) * strideT];
      for (; p <= p_hi; p += UN) {
#pragma unroll
        for (int u = 0; u < UN; ++u) xw[RB - 1 + u] = xp[(tb + RB - 1 + u) * strideT];
        float dv[UN][KS];
#pragma unroll
        for (int u = 0; u < UN; ++u)
#pragma unroll
          for (int s = 0; s < KS; ++s) {
            int ll = l + PL - s;
            ll = min(ll, g.SL + PL - 1);
            dv[u][s] = dp[ll * strideL + (p + u) * strideT];
          }
#pragma unroll
        for (int u = 0; u < UN; ++u)
#pragma unroll
          for (int j = 0; j < RB; ++j)
#pragma unroll
            for (int s = 0; s < KS; ++s) acc[j][s] = fmaf(xw[j + u], dv[u][s], acc[j][s]);
#pragma unroll
        for (int j = 0; j < RB - 1; ++j) xw[j] = xw[j + UN];
        tb += UN;
      }
    }
    __syncthreads();
  }

  // lanes -> warp
#pragma unroll
  for (int j = 0; j < RB; ++j)
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      float v = acc[j][s];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp * (RB * KS) + j * KS + s] = item_ok ? v : 0.f;
    }
  __syncthreads();
  // warps -> CTA (fixed order), one thread per tap of each tap block owned by this CTA
  // this CTA (blockIdx.z) owns items [z*NWARPS, z*NWARPS+NWARPS); taps may be split across z
  // chunks, so every z chunk writes its own partial slot.
  const int taps = KL * KS;
  float* out = partial + (((size_t)blockIdx.z * gridDim.y + part) * C + c) * taps;
  for (int t = tid; t < taps; t += NTHREADS) {
    int r, s;
    if (VERT) { r = t / KS; s = t - r * KS; } else { s = t / KL; r = t - s * KL; }
    const int b = r / RB, j = r - b * RB;
    float v = 0.f;
    for (int wq = 0; wq < NWARPS; ++wq) {
      const int it = blockIdx.z * NWARPS + wq;
      if (it >= g.items) break;
      const int rm = it % per_pg;
      if (rm % g.nseg == b) v += red[wq * (RB * KS) + j * KS + s];
    }
    out[t] = v;
  }
}

// partial[parts][C*taps] -> dw[C*taps], fixed summation order
__global__ void dw_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                       int parts, int total) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float v = 0.f;
  for (int p = 0; p < parts; ++p) v += partial[(size_t)p * total + i];
  dw[i] = v;
}

// ---------------------------------------------------------------------------------
// generic tiled kernels (any odd kh x kw)
// ---------------------------------------------------------------------------------
constexpr int GT = 32;  // output tile edge

template <typename T, typename WT>
__global__ void __launch_bounds__(NTHREADS)
dw_fwd_generic_kernel(const T* __restrict__ x, const WT* __restrict__ w, T* __restrict__ y,
                      int N, int C, int H, int W, int kh, int kw, int flip, int n_per_cta) {
  extern __shared__ __align__(16) float smem[];
  const int ph = kh / 2, pw = kw / 2;
  const int th = GT + kh - 1, tw = GT + kw - 1;
  const int tpitch = tw | 1;
  float* wsm = smem;                  // [kh*kw]
  float* xt = smem + ((kh * kw + 3) & ~3);  // [th][tpitch]
  const int tid = threadIdx.x;
  const int c = blockIdx.x;
  const int tiles_w = (W + GT - 1) / GT, tiles_h = (H + GT - 1) / GT;
  const int n_begin = blockIdx.y * n_per_cta, n_end = min(N, n_begin + n_per_cta);
  for (int i = tid; i < kh * kw; i += NTHREADS) {
    int r = i / kw, s = i - r * kw;
    int gi = flip ? ((kh - 1 - r) * kw + (kw - 1 - s)) : i;
    wsm[i] = round_to<T>(to_f32<WT>(w[(size_t)c * kh * kw + gi]));
  }
  const int ty = tid >> 3, tx = (tid & 7) * 4;  // 32 rows x 8 column quads
  for (int n = n_begin; n < n_end; ++n) {
    const T* xp = x + ((size_t)n * C + c) * H * W;
    T* yp = y + ((size_t)n * C + c) * H * W;
    for (int tile = 0; tile < tiles_h * tiles_w; ++tile) {
      const int h0 = (tile / tiles_w) * GT, w0 = (tile % tiles_w) * GT;
      __syncthreads();
      for (int i = tid; i < th * tw; i += NTHREADS) {
        int r = i / tw, cc = i - r * tw;
        int hh = h0 + r - ph, ww = w0 + cc - pw;
        float v = 0.f;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = to_f32<T>(xp[hh * W + ww]);
        xt[r * tpitch + cc] = v;
      }
      __syncthreads();
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int r = 0; r < kh; ++r) {
        const float* xr = xt + (ty + r) * tpitch + tx;
        const float* wr = wsm + r * kw;
        float x0 = xr[0], x1 = xr[1], x2 = xr[2];
        for (int s = 0; s < kw; ++s) {
          float x3 = xr[s + 3];
          float wv = wr[s];
          a0 = fmaf(x0, wv, a0); a1 = fmaf(x1, wv, a1); a2 = fmaf(x2, wv, a2); a3 = fmaf(x3, wv, a3);
          x0 = x1; x1 = x2; x2 = x3;
        }
      }
      const int oh = h0 + ty, ow = w0 + tx;
      if (oh < H) {
        if (ow + 0 < W) yp[oh * W + ow + 0] = from_f32<T>(a0);
        if (ow + 1 < W) yp[oh * W + ow + 1] = from_f32<T>(a1);
        if (ow + 2 < W) yp[oh * W + ow + 2] = from_f32<T>(a2);
        if (ow + 3 < W) yp[oh * W + ow + 3] = from_f32<T>(a3);
      }
    }
  }
}

// generic wgrad: each thread owns taps; dy tile broadcast, x tile shifted reads.
template <typename T>
__global__ void __launch_bounds__(NTHREADS)
dw_wgrad_generic_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ partial,
                        int N, int C, int H, int W, int kh, int kw, int n_per_cta) {
  extern __shared__ __align__(16) float smem[];
  const int ph = kh / 2, pw = kw / 2;
  const int th = GT + kh - 1, tw = GT + kw - 1;
  const int tpitch = tw | 1;
  float* xt = smem;                 // [th][tpitch]
  float* dt = smem + th * tpitch;   // [GT][GT]
  const int tid = threadIdx.x;
  const int c = blockIdx.x, part = blockIdx.y;
  const int taps = kh * kw;
  const int tap_begin = blockIdx.z * NTHREADS;
  const int tap = tap_begin + tid;
  const bool tap_ok = tap < taps;
  const int r = tap_ok ? tap / kw : 0, s = tap_ok ? tap - (tap / kw) * kw : 0;
  const int tiles_w = (W + GT - 1) / GT, tiles_h = (H + GT - 1) / GT;
  const int n_begin = part * n_per_cta, n_end = min(N, n_begin + n_per_cta);
  float acc = 0.f;
  for (int n = n_begin; n < n_end; ++n) {
    const T* xp = x + ((size_t)n * C + c) * H * W;
    const T* dp = dy + ((size_t)n * C + c) * H * W;
    for (int tile = 0; tile < tiles_h * tiles_w; ++tile) {
      const int h0 = (tile / tiles_w) * GT, w0 = (tile % tiles_w) * GT;
      __syncthreads();
      for (int i = tid; i < th * tw; i += NTHREADS) {
        int rr = i / tw, cc = i - rr * tw;
        int hh = h0 + rr - ph, ww = w0 + cc - pw;
        float v = 0.f;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = to_f32<T>(xp[hh * W + ww]);
        xt[rr * tpitch + cc] = v;
      }
      for (int i = tid; i < GT * GT; i += NTHREADS) {
        int rr = i / GT, cc = i - rr * GT;
        int hh = h0 + rr, ww = w0 + cc;
        dt[i] = (hh < H && ww < W) ? to_f32<T>(dp[hh * W + ww]) : 0.f;
      }
      __syncthreads();
      if (tap_ok) {
        for (int p = 0; p < GT; ++p) {
          const float* xr = xt + (p + r) * tpitch + s;
          const float* dr = dt + p * GT;
#pragma unroll 8
          for (int q = 0; q < GT; ++q) acc = fmaf(dr[q], xr[q], acc);
        }
      }
    }
  }
  if (tap_ok) partial[((size_t)part * C + c) * taps + tap] = acc;
}

// ---------------------------------------------------------------------------------
// host-side geometry + dispatch
// ---------------------------------------------------------------------------------
static bool fast_ok(int H, int W, int kh, int kw, int* KS, bool* vert) {
  // short side must be 3/5/7; prefer the orientation whose long axis is the larger side
  auto is_short = [](int k) { return k == 3 || k == 5 || k == 7; };
  if (kw <= kh && is_short(kw)) { *KS = kw; *vert = true; }
  else if (is_short(kh)) { *KS = kh; *vert = false; }
  else return false;
  // whole plane must fit in shared memory with its border
  size_t bytes = (size_t)(H + 2 * PADT_WG + 4) * (W + 2 * PADT_WG + 4) * 4 * 2 + 8192;
  return bytes <= 200 * 1024;
}

static FastGeom make_geom_fwd(int N, int H, int W, int KL, int KS, bool vert) {
  FastGeom g{};
  const int PL = KS / 2;
  g.H = H; g.W = W; g.KL = KL;
  g.SL = vert ? W : H; g.TL = vert ? H : W;
  if (vert) { g.rows = H + 2 * PADT; g.pitch = W + 2 * PL; g.roff = PADT; g.coff = PL; }
  else      { g.rows = H + 2 * PL; g.pitch = (W + 2 * PADT) | 1; g.roff = PL; g.coff = PADT; }
  g.ppw = (g.SL + PL <= 16) ? 32 / (g.SL + PL) : 1;
  g.nwin = (g.ppw > 1) ? 1 : (g.SL + (32 - 2 * PL) - 1) / (32 - 2 * PL);
  g.nseg = (g.TL + TP - 1) / TP;
  const int per_pg = g.nwin * g.nseg;
  int groups = (2 * NWARPS + per_pg - 1) / per_pg;   // aim at >= 16 items per iteration
  if (groups < 1) groups = 1;
  g.G = g.ppw * groups;
  g.opitch = vert ? W : (W | 1);
  g.wfloats = ((KL + UN + 3) & ~3) * KS;
  g.wfloats = (g.wfloats + 3) & ~3;
  // cap smem (~96 KB) and batch
  auto bytes = [&](int G) { return (size_t)(g.wfloats + G * g.rows * g.pitch + G * H * g.opitch) * 4; };
  while (g.G > g.ppw && bytes(g.G) > 96 * 1024) g.G -= g.ppw;
  if (g.G > N) g.G = ((N + g.ppw - 1) / g.ppw) * g.ppw;
  g.items = ((g.G + g.ppw - 1) / g.ppw) * per_pg;
  g.zchunks = 1;
  return g;
}

static FastGeom make_geom_wgrad(int N, int H, int W, int KL, int KS, bool vert) {
  FastGeom g{};
  const int PL = KS / 2;
  g.H = H; g.W = W; g.KL = KL;
  g.SL = vert ? W : H; g.TL = vert ? H : W;
  if (vert) { g.rows = H + 2 * PADT_WG; g.pitch = W + 2 * PL; g.roff = PADT_WG; g.coff = PL; }
  else      { g.rows = H + 2 * PL; g.pitch = (W + 2 * PADT_WG) | 1; g.roff = PL; g.coff = PADT_WG; }
  g.ppw = (g.SL <= 16) ? 32 / g.SL : 1;
  g.nwin = (g.ppw > 1) ? 1 : (g.SL + 31) / 32;
  g.nseg = (KL + RB - 1) / RB;  // tap blocks
  const int per_pg = g.nwin * g.nseg;
  int groups = per_pg >= NWARPS ? 1 : NWARPS / per_pg;
  g.G = g.ppw * groups;
  auto bytes = [&](int G) { return (size_t)(2 * G * g.rows * g.pitch + NWARPS * RB * KS) * 4; };
  while (g.G > g.ppw && bytes(g.G) > 160 * 1024) g.G -= g.ppw;
  g.items = (g.G / g.ppw) * per_pg;
  g.zchunks = (g.items + NWARPS - 1) / NWARPS;
  g.opitch = 0; g.wfloats = 0;
  return g;
}

template <typename T, typename WT, int KS, bool VERT>
static int launch_fwd_fast(const void* x, const void* w, void* y, int N, int C, const FastGeom& g,
                           int flip, cudaStream_t st) {
  auto kern = dw_fwd_fast_kernel<T, WT, KS, VERT>;
  size_t smem = (size_t)(g.wfloats + g.G * g.rows * g.pitch + g.G * g.H * g.opitch) * 4;
  SLAK_SET_MAX_SMEM(kern, smem);
  // CTAs: C x parts, ~4 waves of 2 CTAs/SM, each CTA at least 2 iterations when the batch allows
  int target = 4 * 2 * sm_count();
  int parts = (target + C - 1) / C;
  int max_parts = (N + g.G - 1) / g.G;
  if (parts > max_parts) parts = max_parts;
  if (parts < 1) parts = 1;
  int n_per_cta = (N + parts - 1) / parts;
  n_per_cta = ((n_per_cta + g.G - 1) / g.G) * g.G;
  parts = (N + n_per_cta - 1) / n_per_cta;
  dim3 grid(C, parts);
  kern<<<grid, NTHREADS, smem, st>>>((const T*)x, (const WT*)w, (T*)y, N, C, g, flip, n_per_cta);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

template <typename T, typename WT>
static int dispatch_fwd_fast(const void* x, const void* w, void* y, int N, int C, int H, int W,
                             int kh, int kw, int KS, bool vert, int flip, cudaStream_t st) {
  const int KL = vert ? kh : kw;
  FastGeom g = make_geom_fwd(N, H, W, KL, KS, vert);
#define SLAK_CASE(ks)                                                                        \
  case ks:                                                                                   \
    return vert ? launch_fwd_fast<T, WT, ks, true>(x, w, y, N, C, g, flip, st)               \
                : launch_fwd_fast<T, WT, ks, false>(x, w, y, N, C, g, flip, st);
  switch (KS) { SLAK_CASE(3) SLAK_CASE(5) SLAK_CASE(7) }
#undef SLAK_CASE
  return SLAK_ERR_UNSUPPORTED;
}

template <typename T, typename WT>
static int launch_fwd_generic(const void* x, const void* w, void* y, int N, int C, int H, int W,
                              int kh, int kw, int flip, cudaStream_t st) {
  auto kern = dw_fwd_generic_kernel<T, WT>;
  const int th = GT + kh - 1, tw = GT + kw - 1;
  size_t smem = (size_t)(((kh * kw + 3) & ~3) + th * (tw | 1)) * 4;
  SLAK_REQUIRE(smem <= 220 * 1024, SLAK_ERR_UNSUPPORTED, "kernel %dx%d too large for the generic path", kh, kw);
  SLAK_SET_MAX_SMEM(kern, smem);
  int target = 8 * sm_count();
  int parts = (target + C - 1) / C;
  if (parts > N) parts = N;
  int n_per_cta = (N + parts - 1) / parts;
  parts = (N + n_per_cta - 1) / n_per_cta;
  dim3 grid(C, parts);
  kern<<<grid, NTHREADS, smem, st>>>((const T*)x, (const WT*)w, (T*)y, N, C, H, W, kh, kw, flip, n_per_cta);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

template <typename T, typename WT>
static int conv_typed(const void* x, const void* w, void* y, int N, int C, int H, int W, int kh,
                      int kw, int flip, cudaStream_t st) {
  int KS; bool vert;
  if (fast_ok(H, W, kh, kw, &KS, &vert))
    return dispatch_fwd_fast<T, WT>(x, w, y, N, C, H, W, kh, kw, KS, vert, flip, st);
  return launch_fwd_generic<T, WT>(x, w, y, N, C, H, W, kh, kw, flip, st);
}

int dwconv_simt_fwd(const void* x, const void* w, void* y, int N, int C, int H, int W, int kh, int kw,
                    int dtype, int wdtype, int flip, cudaStream_t st) {
  if (dtype == SLAK_F32) return conv_typed<float, float>(x, w, y, N, C, H, W, kh, kw, flip, st);
  if (dtype == SLAK_F16) {
    if (wdtype == SLAK_F32) return conv_typed<__half, float>(x, w, y, N, C, H, W, kh, kw, flip, st);
    return conv_typed<__half, __half>(x, w, y, N, C, H, W, kh, kw, flip, st);
  }
  if (wdtype == SLAK_F32) return conv_typed<__nv_bfloat16, float>(x, w, y, N, C, H, W, kh, kw, flip, st);
  return conv_typed<__nv_bfloat16, __nv_bfloat16>(x, w, y, N, C, H, W, kh, kw, flip, st);
}

// ---- wgrad ------------------------------------------------------------------------
struct WgradPlan {
  bool fast; int KS; bool vert; FastGeom g; int parts; int n_per_cta; int slots; size_t ws_bytes;
};

static WgradPlan plan_wgrad(int N, int C, int H, int W, int kh, int kw) {
  WgradPlan p{};
  p.fast = fast_ok(H, W, kh, kw, &p.KS, &p.vert);
  const int taps = kh * kw;
  if (p.fast) {
    const int KL = p.vert ? kh : kw;
    p.g = make_geom_wgrad(N, H, W, KL, p.KS, p.vert);
    int target = 2 * 2 * sm_count();
    int parts = (target + C * p.g.zchunks - 1) / (C * p.g.zchunks);
    int max_parts = (N + p.g.G - 1) / p.g.G;
    if (parts > max_parts) parts = max_parts;
    if (parts < 1) parts = 1;
    int npc = (N + parts - 1) / parts;
    npc = ((npc + p.g.G - 1) / p.g.G) * p.g.G;
    p.n_per_cta = npc;
    p.parts = (N + npc - 1) / npc;
    p.slots = p.parts * p.g.zchunks;
  } else {
    int target = 4 * sm_count();
    int zc = (taps + NTHREADS - 1) / NTHREADS;
    int parts = (target + C * zc - 1) / (C * zc);
    if (parts > N) parts = N;
    if (parts < 1) parts = 1;
    p.n_per_cta = (N + parts - 1) / parts;
    p.parts = (N + p.n_per_cta - 1) / p.n_per_cta;
    p.slots = p.parts;
  }
  p.ws_bytes = (size_t)p.slots * C * taps * sizeof(float);
  return p;
}

size_t dwconv_simt_wgrad_workspace(int N, int C, int H, int W, int kh, int kw) {
  return plan_wgrad(N, C, H, W, kh, kw).ws_bytes;
}

template <typename T, int KS, bool VERT>
static int launch_wgrad_fast(const void* x, const void* dy, float* partial, int N, int C,
                             const WgradPlan& p, cudaStream_t st) {
  auto kern = dw_wgrad_fast_kernel<T, KS, VERT>;
  const FastGeom& g = p.g;
  size_t smem = (size_t)(2 * g.G * g.rows * g.pitch + NWARPS * RB * KS) * 4;
  SLAK_SET_MAX_SMEM(kern, smem);
  dim3 grid(C, p.parts, g.zchunks);
  kern<<<grid, NTHREADS, smem, st>>>((const T*)x, (const T*)dy, partial, N, C, g, p.n_per_cta);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

template <typename T>
static int wgrad_typed(const void* dy, const void* x, float* dw, int N, int C, int H, int W, int kh,
                       int kw, float* partial, cudaStream_t st) {
  WgradPlan p = plan_wgrad(N, C, H, W, kh, kw);
  const int taps = kh * kw;
  int rc = SLAK_OK;
  if (p.fast) {
#define SLAK_CASE(ks)                                                                   \
  case ks:                                                                              \
    rc = p.vert ? launch_wgrad_fast<T, ks, true>(x, dy, partial, N, C, p, st)           \
                : launch_wgrad_fast<T, ks, false>(x, dy, partial, N, C, p, st);         \
    break;
    switch (p.KS) { SLAK_CASE(3) SLAK_CASE(5) SLAK_CASE(7) default: rc = SLAK_ERR_UNSUPPORTED; }
#undef SLAK_CASE
  } else {
    auto kern = dw_wgrad_generic_kernel<T>;
    const int th = GT + kh - 1, tw = GT + kw - 1;
    size_t smem = (size_t)(th * (tw | 1) + GT * GT) * 4;
    SLAK_REQUIRE(smem <= 220 * 1024, SLAK_ERR_UNSUPPORTED, "kernel %dx%d too large for the generic path", kh, kw);
    SLAK_SET_MAX_SMEM(kern, smem);
    dim3 grid(C, p.parts, (taps + NTHREADS - 1) / NTHREADS);
    kern<<<grid, NTHREADS, smem, st>>>((const T*)x, (const T*)dy, partial, N, C, H, W, kh, kw, p.n_per_cta);
    SLAK_CUDA_TRY(cudaGetLastError());
  }
  if (rc != SLAK_OK) return rc;
  const int total = C * taps;
  dw_wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(partial, dw, p.slots, total);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

int dwconv_simt_wgrad(const void* dy, const void* x, float* dw, int N, int C, int H, int W, int kh,
                      int kw, int dtype, void* workspace, cudaStream_t st) {
  float* partial = (float*)workspace;
  if (dtype == SLAK_F32) return wgrad_typed<float>(dy, x, dw, N, C, H, W, kh, kw, partial, st);
  if (dtype == SLAK_F16) return wgrad_typed<__half>(dy, x, dw, N, C, H, W, kh, kw, partial, st);
  return wgrad_typed<__nv_bfloat16>(dy, x, dw, N, C, H, W, kh, kw, partial, st);
}

}  // namespace slak
