// Small planes (14 x 14, 7 x 7: stages 3 and 4 of SLaK, 12 of its 18 Blocks) as DENSE per-channel GEMMs on tcgen05.
//
// The banded-Toeplitz kernels of dwconv_tc_fwd/dgrad.cu tile a plane into 16 x 16 MMA operands; for planes of 392 / 98
// bytes every row of such a tile has to be fetched by 4- or 2-byte pieces and the loaders are the bound (0.03 - 0.19 of
// the HBM roofline).  Here a whole plane is ONE row of a GEMM operand instead:
//
//   forward   Y_b[n, :] = X[n, :] . M_b^T        b = K x 5, 5 x K, 5 x 5 branch;  M_b[out, in] = w_b[in - out + pad]
//   dgrad     dX[n, :]  = sum_b dY_b[n, :] . M_b (+ the fp32 shortcut gradient)
//
// per channel: A = 128 images x P pixels (P = H W <= 208), exactly what is contiguous in the NCHW tensor, so TMA loads it
// as K-major SWIZZLE_128B tiles (rows = images, row stride = C P elements) and stores the result the same way; B = the
// P x P matrix of the branch, built in shared memory from the channel's taps (a few thousand cycles, hidden behind the
// previous branch's epilogue).  The dense matrix wastes 3 - 4 x the FLOPs of the separable form, which is irrelevant:
// 39 MMAs of 128 x 208 x 16 per channel are ~4 k cycles against ~3 k cycles of HBM time for the channel's planes.
// A plane of P bf16 is only 8- (P = 196) or 2-byte (P = 49) aligned and TMA boxes must start on 16 bytes in the inner
// dimension (tools/probes/tma_unaligned_probe.cu: illegal instruction otherwise).  The tensor maps therefore see G planes as
// one row of G P elements (a multiple of 16 bytes); a plane is loaded from the 16-byte boundary below its first pixel and
// the matrix is built with its K index shifted by those delta <= 7 elements (the strangers in front meet zero rows).
// Results leave through a shared-memory slab and coalesced 8- or 2-byte stores (bf16), or TMA boxes where rows are
// 16-byte aligned (the fp32 gradient of 14 x 14 planes).
//
// One CTA = 640 threads: w0 MMA issuer (one thread) | w1 TMEM allocator, then TMA producer (one thread) | w4-7 and w16-19
// epilogue (thread = image row = TMEM lane; the forward kernel splits a tile's 32-pixel chunks between the two sets) |
// w8-15 matrix builders.  Work unit = (channel, tile of 128 images), round-robin over a persistent grid.
#include "common.cuh"
#include "tc_common.cuh"
#include <string.h>
#include <cuda.h>

namespace slak {
namespace tc {
namespace dense {

constexpr int kThreads = 640;                    // 20 warps: see the role list above
constexpr int kBox = 64 * 128;                   // one TMA box: 64 rows x 128 bytes
constexpr int kXSlab = 2 * kBox;                 // 128 images x 64 pixels (bf16)
constexpr int kMaxKB = 4;                        // K blocks of 64 pixels: P <= 208 (13 k16 steps)
constexpr int kMaxNpad = 208;
constexpr int kOffX = 0;
constexpr int kOffB = kOffX + kMaxKB * kXSlab;                       // 65536
constexpr int kOffStg = kOffB + kMaxNpad * kMaxKB * 128;             // + 106496
constexpr int kStgSlabs = 3;                                         // staging slabs of 128 rows x 128 bytes
constexpr int kOffTaps = kOffStg + kStgSlabs * kXSlab;               // + 49152
constexpr int kTapBuf = 2048;                                        // one channel's taps as bf16: (2*99*5 + 25) * 2 bytes
constexpr int kOffHW = kOffTaps + 2 * kTapBuf;                       // pixel -> (h, w) bytes, 2 x 256
constexpr int kOffRed = kOffHW + 512;                                // statistics scratch [8 warps][6]
constexpr int kOffBar = kOffRed + 256;
constexpr int kSmem = kOffBar + 256 + 1024;
static_assert(kSmem <= 232448, "shared memory budget");

struct Params {
  const float* w[3];              // taps [C][KL][5], [C][5][KL], [C][5][5] (fp32)
  __nv_bfloat16* y[3];            // forward: outputs (for the P % 8 tail pixels the boxes cannot store)
  const float* addend;            // dgrad: fp32 [N][C][P] added to the result (the shortcut gradient)
  float* dx;                      // dgrad: fp32 [N][C][P]
  float* stats;                   // forward: [C][ntile][6] (sum, sum of squares) x 3 of the fp32 results, or null
  int N, C, H, W, KL, P;
  int G, Gf;                      // planes per tensor-map row: bf16 maps, fp32 maps
  int NK, KB, NPAD;               // k16 steps, 64-pixel K blocks, MMA N (= NK * 16)
  int ntile, units;               // image tiles per channel, work units
  int dbg;                        // timing experiments (SLAK_DENSE_DBG): 1 = skip the matrix build, 2 = skip the global stores
};

// barrier indices
// barrier indices (operand buffers: three of each when a tile is a single 64-pixel K block, else one)
enum { B_X_FULL = 0, B_X_EMPTY = 3, B_B_FULL = 6, B_B_EMPTY = 9, B_ACC_FULL = 12, B_ACC_EMPTY = 14, B_H_FULL = 16, B_COUNT = 19 };

struct Ctx {
  uint8_t* sm;
  uint32_t base, bar0;
  __device__ __forceinline__ uint32_t bar(int i) const { return bar0 + 8u * i; }
};

// ---- matrix of one branch: Bm[n][k] (K-major SWIZZLE_128B, rows n < NPAD, k < kb * 64) ----------------------------------
// forward: n = output pixel, k - delta = input pixel;  dgrad: n = input pixel, k - delta = output pixel.
// Entry = w_b[in - out + pad].  The matrix is a band: zero-fill the slabs with 16-byte stores, then write the band only --
// one item = (row n, image row r of the other side): a run of <= min(kw, W) consecutive k whose taps are consecutive too
// (ascending for the forward matrix, descending for its transpose).  Taps sit in shared memory as bf16 already.
template <bool DGRAD>
__device__ __forceinline__ void build_matrix(const Ctx& cx, const Params& P, int b, int delta, int kbu, uint8_t* Bs, int tapbuf,
                                             int t0, int nthr) {
  const unsigned short* taps = reinterpret_cast<const unsigned short*>(cx.sm + kOffTaps + tapbuf * kTapBuf);
  const uint8_t* hh = cx.sm + kOffHW;
  const uint8_t* ww = hh + 256;
  const int KL = P.KL, Hh = P.H, Ww = P.W;
  const unsigned short* tb = b == 0 ? taps : (b == 1 ? taps + KL * 5 : taps + 2 * KL * 5);
  const int kh = b == 0 ? KL : 5, kw = b == 1 ? KL : 5;
  const int ph = kh / 2, pw = kw / 2;
  const int slab = P.NPAD * 128;
  {
    const int nvec = kbu * P.NPAD * 8;
    uint4* z = reinterpret_cast<uint4*>(Bs);
    for (int i = t0; i < nvec; i += nthr) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  named_bar_sync(2, nthr);
  const int R = kh < Hh ? kh : Hh;                          // rows of the other side a row n can touch
  const int items = P.P * R;
  for (int it = t0; it < items; it += nthr) {
    const int n = it / R, j = it - n * R;
    const int hn = hh[n], wn = ww[n];
    const int r = kh < Hh ? hn - ph + j : j;
    if ((unsigned)r >= (unsigned)Hh) continue;
    const int dh = DGRAD ? (hn - r + ph) : (r - hn + ph);
    if ((unsigned)dh >= (unsigned)kh) continue;
    const int wlo = max(0, wn - pw), whi = min(Ww - 1, wn + pw);
    int k = delta + r * Ww + wlo;
    const unsigned short* tp = tb + dh * kw + (DGRAD ? (wn - wlo + pw) : (wlo - wn + pw));
    uint8_t* rowp = Bs + n * 128;
    const int sw = n & 7;
    for (int w = wlo; w <= whi; ++w, ++k, tp += (DGRAD ? -1 : 1))
      *reinterpret_cast<unsigned short*>(rowp + (k >> 6) * slab + ((((k >> 3) & 7) ^ sw) << 4) + (k & 7) * 2) = *tp;
  }
}

__device__ __forceinline__ void load_taps(const Ctx& cx, const Params& P, int c, int tapbuf, int t0, int nthr) {
  unsigned short* taps = reinterpret_cast<unsigned short*>(cx.sm + kOffTaps + tapbuf * kTapBuf);
  const int n1 = P.KL * 5;
  for (int i = t0; i < 2 * n1 + 25; i += nthr) {
    float v;
    if (i < n1) v = P.w[0][(size_t)c * n1 + i];
    else if (i < 2 * n1) v = P.w[1][(size_t)c * n1 + (i - n1)];
    else v = P.w[2][(size_t)c * 25 + (i - 2 * n1)];
    taps[i] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
}

// ====================================================================================================================
// position of channel c's plane inside its tensor-map row: TMA start coordinate (16-byte aligned) and the delta in front
struct PlanePos { int cg, c0, delta, nk, kb; };
__device__ __forceinline__ PlanePos plane_pos(const Params& P, int c) {
  PlanePos pp;
  pp.cg = c / P.G;
  const int start = (c - pp.cg * P.G) * P.P;
  pp.delta = start & 7;
  pp.c0 = start - pp.delta;
  pp.nk = (pp.delta + P.P + 15) >> 4;
  pp.kb = (pp.delta + P.P + 63) >> 6;
  return pp;
}

template <bool DGRAD>
__global__ void __launch_bounds__(kThreads, 1)
dense_kernel(const __grid_constant__ CUtensorMap in0, const __grid_constant__ CUtensorMap in1, const __grid_constant__ CUtensorMap in2,
             const __grid_constant__ CUtensorMap fa_full, const __grid_constant__ CUtensorMap fa_part,
             const __grid_constant__ CUtensorMap fd_full, const __grid_constant__ CUtensorMap fd_part, Params P) {
  // forward: in0 = x.   dgrad: in_b = dy_b; fa_* / fd_* = full (32-pixel, SWIZZLE_128B) and partial (rem4-pixel, unswizzled)
  // fp32 maps of the addend and of dx, used when the fp32 rows are 16-byte aligned (P % 4 == 0)
  extern __shared__ uint8_t smem_raw[];
  Ctx cx;
  const uint32_t raw = smem_u32(smem_raw);
  cx.base = (raw + 1023u) & ~1023u;
  cx.sm = smem_raw + (cx.base - raw);
  cx.bar0 = cx.base + kOffBar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cx.sm + kOffBar + 192);
  // a tile of one 64-pixel K block (planes of <= 57 pixels): three A buffers and three matrix buffers, so that the loads
  // and the builds of a unit's three branches (and of the next unit) run ahead of the MMAs; larger planes: one of each
  const int nb = P.KB == 1 ? 3 : 1;
  const uint32_t xbuf_bytes = (uint32_t)P.KB * kXSlab, bbuf_bytes = (uint32_t)P.NPAD * P.KB * 128;

  if (tid == 0) {
    for (int i = 0; i < 3; ++i) {
      mbar_init(cx.bar(B_X_FULL + i), 1); mbar_init(cx.bar(B_X_EMPTY + i), 1);
      mbar_init(cx.bar(B_B_FULL + i), 8); mbar_init(cx.bar(B_B_EMPTY + i), 1);
      mbar_init(cx.bar(B_H_FULL + i), 1);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(cx.bar(B_ACC_FULL + a), 1); mbar_init(cx.bar(B_ACC_EMPTY + a), DGRAD ? 4 : 8); }
    mbar_fence_init();
    tma_prefetch_desc(&in0);
    if (DGRAD) { tma_prefetch_desc(&in1); tma_prefetch_desc(&in2); tma_prefetch_desc(&fa_full); tma_prefetch_desc(&fd_full); }
  }
  for (int i = tid; i < 256; i += kThreads) {                     // pixel -> (h, w)
    const int h = i / P.W;
    cx.sm[kOffHW + i] = (uint8_t)h;
    cx.sm[kOffHW + 256 + i] = (uint8_t)(i - h * P.W);
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 1) {
    // ================= TMA producer: A tiles in consumption order (forward: x once per unit; dgrad: dy_g per group) =================
    if (elect_one()) {
      int la = 0;
      for (int u = blockIdx.x; u < P.units; u += gridDim.x) {
        const int c = u / P.ntile, n0 = (u - c * P.ntile) * 128;
        const PlanePos pp = plane_pos(P, c);
        for (int g = 0; g < (DGRAD ? 3 : 1); ++g, ++la) {
          const int xb = la % nb;
          mbar_wait(cx.bar(B_X_EMPTY + xb), ((la / nb) & 1) ^ 1);
          mbar_expect_tx(cx.bar(B_X_FULL + xb), (uint32_t)pp.kb * kXSlab);
          const CUtensorMap* m = DGRAD ? (g == 0 ? &in0 : (g == 1 ? &in1 : &in2)) : &in0;
          const uint32_t dst = cx.base + kOffX + xb * xbuf_bytes;
          for (int kb = 0; kb < pp.kb; ++kb) {
            tma_load_3d(dst + kb * kXSlab, m, cx.bar(B_X_FULL + xb), pp.c0 + kb * 64, pp.cg, n0);
            tma_load_3d(dst + kb * kXSlab + kBox, m, cx.bar(B_X_FULL + xb), pp.c0 + kb * 64, pp.cg, n0 + 64);
          }
        }
      }
    }
  } else if (warp == 0) {
    // ================= MMA issuer (one thread) =================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, P.NPAD);
      int gc = 0, la = 0, uit = 0;                                // groups issued, A tiles consumed, units done
      for (int u = blockIdx.x; u < P.units; u += gridDim.x, ++uit) {
        const PlanePos pp = plane_pos(P, u / P.ntile);
        const int ab = DGRAD ? (uit & 1) : 0;
        for (int g = 0; g < 3; ++g, ++gc) {
          const int xb = la % nb, bb = gc % nb;
          if (DGRAD || g == 0) mbar_wait(cx.bar(B_X_FULL + xb), (la / nb) & 1);
          mbar_wait(cx.bar(B_B_FULL + bb), (gc / nb) & 1);
          uint32_t acc;
          if (DGRAD) {
            if (g == 0) mbar_wait(cx.bar(B_ACC_EMPTY + ab), ((uit >> 1) & 1) ^ 1);
            acc = tmem + ab * 256;
          } else {
            const int fb = gc & 1;
            mbar_wait(cx.bar(B_ACC_EMPTY + fb), ((gc >> 1) & 1) ^ 1);
            acc = tmem + fb * 256;
          }
          tc_fence_after();
          const uint32_t xa = cx.base + kOffX + xb * xbuf_bytes, ba = cx.base + kOffB + bb * bbuf_bytes;
          for (int ks = 0; ks < pp.nk; ++ks) {
            const int kb = ks >> 2, kk = ks & 3;
            umma_bf16(acc, umma_desc_k_sw128(xa + kb * kXSlab + kk * 32, 0), umma_desc_k_sw128(ba + kb * (P.NPAD * 128) + kk * 32, 0),
                      idesc, DGRAD ? ((g | ks) != 0) : (ks != 0));
          }
          umma_commit(cx.bar(B_B_EMPTY + bb));
          if (DGRAD) {
            umma_commit(cx.bar(B_X_EMPTY + xb)); ++la;
            if (g == 2) umma_commit(cx.bar(B_ACC_FULL + ab));
          } else {
            if (g == 2) { umma_commit(cx.bar(B_X_EMPTY + xb)); ++la; }
            umma_commit(cx.bar(B_ACC_FULL + (gc & 1)));
          }
        }
      }
    }
  } else if (warp >= 8 && warp < 16) {
    // ================= matrix builders (256 threads) =================
    const int t0 = tid - 256;
    int gc = 0, tb = 0;
    if (blockIdx.x < P.units) load_taps(cx, P, blockIdx.x / P.ntile, 0, t0, 256);
    for (int u = blockIdx.x; u < P.units; u += gridDim.x, tb ^= 1) {
      const PlanePos pp = plane_pos(P, u / P.ntile);
      for (int g = 0; g < 3; ++g, ++gc) {
        const int bb = gc % nb;
        mbar_wait(cx.bar(B_B_EMPTY + bb), ((gc / nb) & 1) ^ 1);   // the matrix that lived in this buffer has been consumed
        if (!(P.dbg & 1)) build_matrix<DGRAD>(cx, P, g, pp.delta, pp.kb, cx.sm + kOffB + bb * bbuf_bytes, tb, t0, 256);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(cx.bar(B_B_FULL + bb));
      }
      // the next unit's taps into the other buffer (its readers are separated from these stores by the barrier inside
      // build_matrix; the buffer being overwritten was last read two units ago)
      if (u + gridDim.x < P.units) load_taps(cx, P, (u + gridDim.x) / P.ntile, tb ^ 1, t0, 256);
    }
  } else if (warp >= 4 && (warp < 8 || !DGRAD)) {
    // ================= epilogue: thread = image row = TMEM lane; every warp stages and copies out its own 32 rows =================
    const int eset = warp >= 16 ? 1 : 0;                          // forward: second set of epilogue warps (odd chunks)
    const int e = (warp - 4) & 3, L = e * 32 + lane;
    const uint32_t stg_s = cx.base + kOffStg;
    uint8_t* stg = cx.sm + kOffStg;
    float* red = reinterpret_cast<float*>(cx.sm + kOffRed);
    const uint32_t lane_off = (uint32_t)(e * 32) << 16;
    if constexpr (!DGRAD) {
      // ---- forward: bf16 results through a per-warp slab of 32 rows x (64 + 16) bytes and coalesced stores; chunks of 32 pixels,
      // even chunks to warps 4-7, odd chunks to warps 16-19 ----
      constexpr int kPitch = 80;                                  // 16-byte row stores of 8 consecutive lanes: 8 different bank groups
      const bool wide = (P.P & 3) == 0;                           // planes 8-byte aligned: 4 pixels per lane, else 1
      uint8_t* wslab = stg + (eset * 4 + e) * (2 * 32 * kPitch);  // two buffers per warp
      int gc = 0, sc = 0;
      for (int u = blockIdx.x; u < P.units; u += gridDim.x) {
        const int c = u / P.ntile, it = u - c * P.ntile, n0 = it * 128;
        float st[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const size_t rstride = (size_t)P.C * P.P;
        const int rows_ok = min(32, P.N - n0 - 32 * e);           // valid image rows of this warp
        const int nchunks = (P.P + 31) >> 5;
#pragma unroll
        for (int g = 0; g < 3; ++g, ++gc) {
          const int fb = gc & 1;
          mbar_wait(cx.bar(B_ACC_FULL + fb), (gc >> 1) & 1);
          tc_fence_after();
          const uint32_t ta = tmem + lane_off + fb * 256;
          const int last = ((nchunks - 1 - eset) & ~1) + eset;    // this set's last chunk (may be < eset: none)
          if (last < eset) {                                      // no chunk for this set: release the accumulator right away
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(cx.bar(B_ACC_EMPTY + fb));
          }
          f2 s2 = 0ull, q2 = 0ull;
          for (int ch = eset; ch < nchunks; ch += 2, ++sc) {
            uint32_t v[32];
            tmem_ld32(ta + 32 * ch, v);                           // (columns beyond NPAD: never used below)
            tmem_ld_wait();
            if (ch == last) {                                     // this set has drained the accumulator
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(cx.bar(B_ACC_EMPTY + fb));
            }
            const int valid = min(32, P.P - 32 * ch);             // pixels of this chunk
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (2 * j < valid) {
                const f2 a = mk2u(v[2 * j], 2 * j + 1 < valid ? v[2 * j + 1] : 0u);
                s2 = add2(s2, a); q2 = fma2(a, a, q2);
              }
            }
            uint8_t* s0 = wslab + (sc & 1) * (32 * kPitch);
            __syncwarp();                                         // this warp's copy-out of two chunks ago is complete (program order)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (8 * j < valid)
                *reinterpret_cast<uint4*>(s0 + (uint32_t)lane * kPitch + 16 * j) =
                    make_uint4(pack_bf16(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                               pack_bf16(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                               pack_bf16(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                               pack_bf16(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
            __syncwarp();
            if (P.dbg & 2) continue;
            __nv_bfloat16* yb = P.y[g] + ((size_t)(n0 + 32 * e) * P.C + c) * P.P + 32 * ch;    // row r of this warp: + r * C * P
            if (wide) {                                           // 8 lanes x 8 bytes per row, four rows per instruction
              const int q = lane & 7, rh = lane >> 3;
              uint2 t[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) t[i] = *reinterpret_cast<const uint2*>(s0 + (4 * i + rh) * kPitch + 8 * q);
              if (4 * q < valid) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (4 * i + rh < rows_ok) *reinterpret_cast<uint2*>(yb + (size_t)(4 * i + rh) * rstride + 4 * q) = t[i];
              }
            } else {                                              // one pixel per lane, one instruction per row
#pragma unroll 1
              for (int r0 = 0; r0 < 32; r0 += 16) {
                unsigned short t[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) t[i] = *reinterpret_cast<const unsigned short*>(s0 + (r0 + i) * kPitch + 2 * lane);
                if (lane < valid) {
#pragma unroll
                  for (int i = 0; i < 16; ++i)
                    if (r0 + i < rows_ok) reinterpret_cast<unsigned short*>(yb + (size_t)(r0 + i) * rstride)[lane] = t[i];
                }
              }
            }
          }
          { float lo, hi; un2(s2, lo, hi); st[2 * g] = lo + hi; un2(q2, lo, hi); st[2 * g + 1] = lo + hi; }
        }
        if (P.stats) {                                            // (sum, sum of squares) x 3 of this unit
#pragma unroll
          for (int k2 = 0; k2 < 6; ++k2) {
            float sv = st[k2];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
            if (lane == 0) red[(eset * 4 + e) * 6 + k2] = sv;
          }
          named_bar_sync(1, 256);
          if (eset == 0 && e == 0 && lane < 6) {
            float t = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) t += red[w8 * 6 + lane];
            P.stats[((size_t)c * P.ntile + it) * 6 + lane] = t;
          }
          named_bar_sync(1, 256);
        }
      }
    } else if ((P.P & 3) == 0) {
      // ---- dgrad, 16-byte aligned fp32 rows: addend and dx through TMA boxes of 32 pixels (128 bytes per row) ----
      const int nfull32 = P.P >> 5;
      const int rem4 = P.P & 31;                                  // pixels of the partial box (a multiple of 4)
      const int nbox = nfull32 + (rem4 ? 1 : 0);
      int hl = 0;                                                 // boxes processed so far (slab = hl % 3)
      // prefetch cursor of the elected thread: (unit, chunk) of the next addend box to request
      int pf_u = blockIdx.x, pf_ch = 0, pf_n = 0;
      auto pf_issue = [&]() {
        if (pf_u >= P.units) return;
        const int c = pf_u / P.ntile, n0 = (pf_u - c * P.ntile) * 128;
        const int slab = pf_n % kStgSlabs;
        const uint32_t dst = stg_s + slab * kXSlab, bar = cx.bar(B_H_FULL + slab);
        if (pf_ch < nfull32) {
          mbar_expect_tx(bar, kXSlab);
          tma_load_3d(dst, &fa_full, bar, 32 * pf_ch, c, n0);
          tma_load_3d(dst + kBox, &fa_full, bar, 32 * pf_ch, c, n0 + 64);
        } else {
          mbar_expect_tx(bar, 128 * rem4 * 4);
          tma_load_3d(dst, &fa_part, bar, 32 * pf_ch, c, n0);
          tma_load_3d(dst + 64 * rem4 * 4, &fa_part, bar, 32 * pf_ch, c, n0 + 64);
        }
        ++pf_n;
        if (++pf_ch == nbox) { pf_ch = 0; pf_u += gridDim.x; }
      };
      if (e == 0 && lane == 0) { pf_issue(); pf_issue(); }
      int uit = 0;
      for (int u = blockIdx.x; u < P.units; u += gridDim.x, ++uit) {
        const int c = u / P.ntile, n0 = (u - c * P.ntile) * 128;
        const int ab = uit & 1;
        mbar_wait(cx.bar(B_ACC_FULL + ab), (uit >> 1) & 1);
        tc_fence_after();
        const uint32_t ta = tmem + lane_off + ab * 256;
        for (int ch = 0; ch < nbox; ++ch, ++hl) {
          uint32_t v[32];
          tmem_ld32(ta + 32 * ch, v);
          tmem_ld_wait();
          if (ch == nbox - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(cx.bar(B_ACC_EMPTY + ab));
          }
          const bool full = ch < nfull32;
          const int slab = hl % kStgSlabs;
          mbar_wait(cx.bar(B_H_FULL + slab), (hl / kStgSlabs) & 1);
          uint8_t* s0 = stg + slab * kXSlab;
          if (full) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4* p4 = reinterpret_cast<float4*>(s0 + (uint32_t)L * 128 + ((j ^ (L & 7)) << 4));
              float4 a = *p4;
              a.x += __uint_as_float(v[4 * j]); a.y += __uint_as_float(v[4 * j + 1]);
              a.z += __uint_as_float(v[4 * j + 2]); a.w += __uint_as_float(v[4 * j + 3]);
              *p4 = a;
            }
          } else {
            const int pitch = rem4 * 4;
#pragma unroll
            for (int j = 0; j < 7; ++j)
              if (4 * j < rem4) {
                float4* p4 = reinterpret_cast<float4*>(s0 + (uint32_t)(L & 63) * pitch + (L >> 6) * (64 * pitch) + 16 * j);
                float4 a = *p4;
                a.x += __uint_as_float(v[4 * j]); a.y += __uint_as_float(v[4 * j + 1]);
                a.z += __uint_as_float(v[4 * j + 2]); a.w += __uint_as_float(v[4 * j + 3]);
                *p4 = a;
              }
          }
          fence_proxy_async();
          named_bar_sync(1, 128);
          if (e == 0 && lane == 0) {
            const uint32_t src = stg_s + slab * kXSlab;
            if (full) {
              tma_store_3d(&fd_full, src, 32 * ch, c, n0);
              tma_store_3d(&fd_full, src + kBox, 32 * ch, c, n0 + 64);
            } else {
              tma_store_3d(&fd_part, src, 32 * ch, c, n0);
              tma_store_3d(&fd_part, src + 64 * rem4 * 4, 32 * ch, c, n0 + 64);
            }
            bulk_commit_group();
            bulk_wait_group_read<1>();          // the store before this one has let go of its slab: the addend two boxes ahead goes there
            pf_issue();
          }
        }
      }
      if (e == 0 && lane == 0) bulk_wait_group_read<0>();
    } else {
      // ---- dgrad, unaligned fp32 rows (7 x 7): per-warp slab of 32 rows x 32 pixels (fp32), lanes along the pixels ----
      constexpr int kPitchF = 144;                                // 36 floats
      uint8_t* wslab = stg + e * (32 * kPitchF);
      int uit = 0;
      for (int u = blockIdx.x; u < P.units; u += gridDim.x, ++uit) {
        const int c = u / P.ntile, n0 = (u - c * P.ntile) * 128;
        const int ab = uit & 1;
        const size_t rstride = (size_t)P.C * P.P;
        const int rows_ok = min(32, P.N - n0 - 32 * e);
        mbar_wait(cx.bar(B_ACC_FULL + ab), (uit >> 1) & 1);
        tc_fence_after();
        const uint32_t ta = tmem + lane_off + ab * 256;
        const int nchunks = (P.P + 31) >> 5;
        for (int ch = 0; ch < nchunks; ++ch) {
          uint32_t v[32];
          tmem_ld32(ta + 32 * ch, v);
          tmem_ld_wait();
          if (ch == nchunks - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(cx.bar(B_ACC_EMPTY + ab));
          }
          const int valid = min(32, P.P - 32 * ch);
          const size_t off0 = ((size_t)(n0 + 32 * e) * P.C + c) * P.P + 32 * ch;
          // addend rows of this warp, coalesced (lanes along the pixels), in flight while the slab is written
          float a[32];
#pragma unroll
          for (int r = 0; r < 32; ++r) a[r] = (lane < valid && r < rows_ok) ? __ldg(P.addend + off0 + (size_t)r * rstride + lane) : 0.f;
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(wslab + (uint32_t)lane * kPitchF + 16 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
#pragma unroll
          for (int r = 0; r < 32; ++r)
            if (lane < valid && r < rows_ok)
              P.dx[off0 + (size_t)r * rstride + lane] = a[r] + *reinterpret_cast<const float*>(wslab + r * kPitchF + 4 * lane);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

// MN-major SWIZZLE_128B operand descriptor and instruction descriptor (both operands MN-major), as in mlp_tc.cu
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t mlp_idesc_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ====================================================================================================================
// Weight gradient of the three branches for planes of <= 200 pixels and batches of <= 128 images per GPU:
//   G_b[i, o] = sum_n X[n, i] dY_b[n, o]      (one GEMM per branch and 128-row tile of i, contraction over the images)
//   dW_b[dh + ph, dw + pw] = sum over (i, o) with i - o = (dh, dw) of G_b[i, o]
// Both operands are MN-major exactly as TMA delivers them (rows = images, 64 pixels = 128 bytes per box row).  The
// diagonal sums are taken by the epilogue straight from TMEM: thread = row i, every column o adds into the tap (i - o) of a
// per-warp private array in shared memory (no two lanes of a warp share a tap for the same column); sixteen epilogue warps
// split the columns four ways; the warps' arrays are folded in a fixed order at the end of the channel: deterministic.
// Roles (640 threads): w0 MMA issuer | w1 TMEM allocator, then TMA producer | w4-19 epilogue.
// ====================================================================================================================
namespace wg {
constexpr int kOffXw = 0;                              // x: [2 image blocks][4 pixel blocks] boxes of 64 x 128 B
constexpr int kOffDy = 8 * kBox;                       // dy_b: the same, two buffers
constexpr int kOffAcc = kOffDy + 2 * 8 * kBox;         // 16 warps x 2 copies (lane parity) x one branch's taps within reach of a plane
constexpr int kAccB = 160;                             // (2*16-1) * 5 = 155 fp32 at most
constexpr int kAccStride = 2 * kAccB;
constexpr int kOffHWw = kOffAcc + 16 * kAccStride * 4;
constexpr int kOffBarW = kOffHWw + 512;
constexpr int kSmemW = kOffBarW + 256 + 1024;
enum { W_X_FULL = 0, W_X_EMPTY, W_DY_FULL, W_DY_EMPTY = W_DY_FULL + 2, W_ACC_FULL = W_DY_EMPTY + 2, W_ACC_EMPTY = W_ACC_FULL + 2 };
}  // namespace wg

struct WgradParams {
  float* dw[3];                   // [C][KL][5], [C][5][KL], [C][5][5]
  int N, C, H, W, KL, P, G;
  int mtiles, pxblocks, npadw;    // 128-row tiles of i, 64-pixel blocks, MMA N
};

__global__ void __launch_bounds__(kThreads, 1)
dense_wgrad_kernel(const __grid_constant__ CUtensorMap xm, const __grid_constant__ CUtensorMap d0, const __grid_constant__ CUtensorMap d1,
                   const __grid_constant__ CUtensorMap d2, WgradParams P) {
  using namespace wg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t bar0 = base + kOffBarW;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kOffBarW + 192);
  if (tid == 0) {
    mbar_init(BAR(W_X_FULL), 1); mbar_init(BAR(W_X_EMPTY), 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(BAR(W_DY_FULL + a), 1); mbar_init(BAR(W_DY_EMPTY + a), 1);
      mbar_init(BAR(W_ACC_FULL + a), 1); mbar_init(BAR(W_ACC_EMPTY + a), 16);
    }
    mbar_fence_init();
    tma_prefetch_desc(&xm); tma_prefetch_desc(&d0); tma_prefetch_desc(&d1); tma_prefetch_desc(&d2);
  }
  for (int i = tid; i < 256; i += kThreads) {
    const int h = i / P.W;
    sm[kOffHWw + i] = (uint8_t)h;
    sm[kOffHWw + 256 + i] = (uint8_t)(i - h * P.W);
  }
  for (int i = tid; i < 16 * kAccStride; i += kThreads) reinterpret_cast<float*>(sm + kOffAcc)[i] = 0.f;
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int npb = P.pxblocks;

  auto pos = [&](int c, int& cg, int& c0, int& delta) {
    cg = c / P.G;
    const int start = (c - cg * P.G) * P.P;
    delta = start & 7;
    c0 = start - delta;
  };

  if (warp == 1) {
    if (elect_one()) {                                    // ---- TMA producer ----
      int ld = 0, uit = 0;
      for (int c = blockIdx.x; c < P.C; c += gridDim.x, ++uit) {
        int cg, c0, delta;
        pos(c, cg, c0, delta);
        mbar_wait(BAR(W_X_EMPTY), (uit & 1) ^ 1);
        mbar_expect_tx(BAR(W_X_FULL), (uint32_t)(2 * npb) * kBox);
        for (int kb = 0; kb < 2; ++kb)
          for (int pb = 0; pb < npb; ++pb)
            tma_load_3d(base + kOffXw + (kb * 4 + pb) * kBox, &xm, BAR(W_X_FULL), c0 + 64 * pb, cg, 64 * kb);
        for (int b = 0; b < 3; ++b, ++ld) {
          const int buf = ld & 1;
          mbar_wait(BAR(W_DY_EMPTY + buf), ((ld >> 1) & 1) ^ 1);
          mbar_expect_tx(BAR(W_DY_FULL + buf), (uint32_t)(2 * npb) * kBox);
          const CUtensorMap* m = b == 0 ? &d0 : (b == 1 ? &d1 : &d2);
          for (int kb = 0; kb < 2; ++kb)
            for (int pb = 0; pb < npb; ++pb)
              tma_load_3d(base + kOffDy + (buf * 8 + kb * 4 + pb) * kBox, m, BAR(W_DY_FULL + buf), c0 + 64 * pb, cg, 64 * kb);
        }
      }
    }
  } else if (warp == 0) {
    if (elect_one()) {                                    // ---- MMA issuer ----
      const uint32_t idesc = mlp_idesc_mn(128, P.npadw);
      int ld = 0, ac = 0, uit = 0;
      for (int c = blockIdx.x; c < P.C; c += gridDim.x, ++uit) {
        mbar_wait(BAR(W_X_FULL), uit & 1);
        for (int b = 0; b < 3; ++b, ++ld) {
          const int buf = ld & 1;
          mbar_wait(BAR(W_DY_FULL + buf), (ld >> 1) & 1);
          for (int mt = 0; mt < P.mtiles; ++mt, ++ac) {
            const int ab = ac & 1;
            mbar_wait(BAR(W_ACC_EMPTY + ab), ((ac >> 1) & 1) ^ 1);
            tc_fence_after();
            for (int ks = 0; ks < 8; ++ks) {
              const int kb = ks >> 2, kk = ks & 3;
              umma_bf16(tmem + ab * 256, desc_mn_sw128(base + kOffXw + (kb * 4 + 2 * mt) * kBox + kk * 2048, kBox),
                        desc_mn_sw128(base + kOffDy + (buf * 8 + kb * 4) * kBox + kk * 2048, kBox), idesc, ks != 0);
            }
            umma_commit(BAR(W_ACC_FULL + ab));
          }
          umma_commit(BAR(W_DY_EMPTY + buf));
        }
        umma_commit(BAR(W_X_EMPTY));
      }
    }
  } else if (warp >= 4) {
    // ---- epilogue: 16 warps; set = (warp - 4) / 4 takes the 32-column chunks ch = set, set + 4, ... ----
    const int w16 = warp - 4, eset = w16 >> 2, e = w16 & 3;
    const uint32_t lane_off = (uint32_t)(e * 32) << 16;
    float* acc = reinterpret_cast<float*>(sm + kOffAcc) + w16 * kAccStride;
    const uint8_t* hh = sm + kOffHWw;
    const uint8_t* ww = hh + 256;
    const int KL = P.KL;
    const int nchunks = (P.npadw + 31) >> 5;
    int ac = 0;
    for (int c = blockIdx.x; c < P.C; c += gridDim.x) {
      int cg, c0, delta;
      pos(c, cg, c0, delta);
      for (int b = 0; b < 3; ++b) {
        // taps within reach of the plane: |dh| <= rh, |dw| <= rw; private arrays [2 rh + 1][2 rw + 1], one per lane parity:
        // two consecutive columns are then two independent read-modify-write chains (lanes that would meet in a tap for
        // columns j and j + 1 are neighbours, i.e. of different parity)
        const int rh = b == 0 ? min(KL / 2, P.H - 1) : 2, rw = b == 1 ? min(KL / 2, P.W - 1) : 2;
        const int kh = 2 * rh + 1, kw = 2 * rw + 1, ph = rh, pw = rw;
        float* ab_ = acc + (lane & 1) * kAccB;
        for (int mt = 0; mt < P.mtiles; ++mt, ++ac) {
          const int ab = ac & 1;
          mbar_wait(BAR(W_ACC_FULL + ab), (ac >> 1) & 1);
          tc_fence_after();
          const int pi = 128 * mt + e * 32 + lane - delta;        // input pixel of this row
          const bool rowok = (unsigned)pi < (unsigned)P.P;
          const int hi = rowok ? hh[pi] : 0, wi = rowok ? ww[pi] : 0;
          // image rows this warp's 32 pixels touch (for skipping chunks out of the band's reach: branches 5 x K, 5 x 5)
          const int p_lo = max(0, 128 * mt + e * 32 - delta), p_hi = min(P.P - 1, 128 * mt + e * 32 + 31 - delta);
          const int wh_lo = p_lo <= p_hi ? hh[p_lo] : 0, wh_hi = p_lo <= p_hi ? hh[p_hi] : -100;
          const uint32_t ta = tmem + lane_off + ab * 256;
          for (int ch = eset; ch < nchunks; ch += 4) {
            const int po0 = 32 * ch - delta;                      // output pixel of the chunk's first column
            const int c_lo = max(0, po0), c_hi = min(P.P - 1, po0 + 31);
            if (c_lo > c_hi) continue;
            if ((int)hh[c_hi] < wh_lo - rh || (int)hh[c_lo] > wh_hi + rh) continue;     // no (row, column) pair within reach
            uint32_t v[32];
            tmem_ld32(ta + 32 * ch, v);
            tmem_ld_wait();
            int ho = hh[c_lo], wo = ww[c_lo];                     // walked along the columns: no table reads in the loop
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              int idx[2];
              bool on[2];
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                const int po = po0 + j + t;
                on[t] = false; idx[t] = 0;
                if (po >= c_lo && po <= c_hi) {                   // (uniform across the warp)
                  const int dh = hi - ho + ph, dw = wi - wo + pw;
                  on[t] = rowok && (unsigned)dh < (unsigned)kh && (unsigned)dw < (unsigned)kw;
                  idx[t] = dh * kw + dw;
                  if (++wo == P.W) { wo = 0; ++ho; }
                }
              }
              const float a0 = on[0] ? ab_[idx[0]] : 0.f, a1 = on[1] ? ab_[idx[1]] : 0.f;
              if (on[0]) ab_[idx[0]] = a0 + __uint_as_float(v[j]);
              if (on[1]) ab_[idx[1]] = a1 + __uint_as_float(v[j + 1]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(W_ACC_EMPTY + ab));
        }
        // fold the 32 private arrays of this branch (fixed order), write the channel's gradient, clear the arrays
        named_bar_sync(1, 512);
        {
          float* all = reinterpret_cast<float*>(sm + kOffAcc);
          const int KH = b == 0 ? KL : 5, KW = b == 1 ? KL : 5;
          float* dst = P.dw[b] + (size_t)c * KH * KW;
          for (int t = tid - 128; t < KH * KW; t += 512) {
            // tap t of dw_b -> its slot in the clipped array (taps out of the plane's reach have zero gradient)
            const int i = t / KW, j = t - i * KW;
            const int r = i - KH / 2 + rh, q = j - KW / 2 + rw;
            float sacc = 0.f;
            if ((unsigned)r < (unsigned)kh && (unsigned)q < (unsigned)kw) {
              const int off = r * kw + q;
#pragma unroll
              for (int w8 = 0; w8 < 32; ++w8) sacc += all[w8 * kAccB + off];
            }
            dst[t] = sacc;
          }
        }
        named_bar_sync(1, 512);
        for (int t = tid - 128; t < 16 * kAccStride; t += 512) reinterpret_cast<float*>(sm + kOffAcc)[t] = 0.f;
        named_bar_sync(1, 512);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

// ====================================================================================================================
// host side
// ====================================================================================================================
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_enc() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)p;
  }
  return fn;
}
// [N][C][P] tensor seen as (G*P, C/G, N); box (inner, 1, 64)
static int make_map(CUtensorMap* m, const void* t, bool f32, int N, int C, int Pp, int G, int inner, bool swizzle) {
  EncodeFn enc = get_enc();
  SLAK_REQUIRE(enc != nullptr, SLAK_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int es = f32 ? 4 : 2;
  cuuint64_t dims[3] = {(cuuint64_t)G * Pp, (cuuint64_t)(C / G), (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)G * Pp * es, (cuuint64_t)C * Pp * es};
  cuuint32_t box[3] = {(cuuint32_t)inner, 1, 64};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(t), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SLAK_REQUIRE(r == CUDA_SUCCESS, SLAK_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return SLAK_OK;
}
static int group_of(int Pp, int C) {                       // planes per tensor-map row: row bytes a multiple of 16
  for (int G = 1; G <= 8; G <<= 1)
    if ((G * Pp * 2) % 16 == 0) return (C % G == 0) ? G : 0;
  return 0;
}
static bool enabled() {
  const char* e = getenv("SLAK_DENSE_PLANES");
  return !(e && atoi(e) == 0);
}
bool supported(int N, int C, int H, int W, int KL) {
  const int Pp = H * W;
  if (!enabled() || N < 1 || H < 1 || W < 1 || H > 16 || W > 16 || Pp > 208 || Pp < 8) return false;
  if (!(KL & 1) || KL < 5 || KL > 99) return false;
  return group_of(Pp, C) != 0;
}
static void fill(Params* P, int N, int C, int H, int W, int KL) {
  P->N = N; P->C = C; P->H = H; P->W = W; P->KL = KL; P->P = H * W;
  P->G = group_of(P->P, C); P->Gf = 1;
  int dmax = 0;                                               // largest element offset of a plane from the 16-byte boundary below it
  for (int j = 0; j < P->G; ++j) dmax = ((j * P->P) & 7) > dmax ? ((j * P->P) & 7) : dmax;
  P->NK = (P->P + 15) / 16; P->KB = (dmax + P->P + 63) / 64; P->NPAD = P->NK * 16;
  P->ntile = (N + 127) / 128; P->units = C * P->ntile;
  const char* d = getenv("SLAK_DENSE_DBG");
  P->dbg = d ? atoi(d) : 0;
}
int stats_slots(int N) { return (N + 127) / 128; }

int fwd(const void* x, const float* w1, const float* w2, const float* w3, void* y1, void* y2, void* y3, int N, int C, int H, int W,
        int KL, float* stats, cudaStream_t st) {
  SLAK_REQUIRE(supported(N, C, H, W, KL), SLAK_ERR_UNSUPPORTED, "dense plane kernel: shape not covered");
  Params P{};
  fill(&P, N, C, H, W, KL);
  P.w[0] = w1; P.w[1] = w2; P.w[2] = w3;
  P.y[0] = (__nv_bfloat16*)y1; P.y[1] = (__nv_bfloat16*)y2; P.y[2] = (__nv_bfloat16*)y3;
  P.stats = stats;
  SLAK_REQUIRE(((reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(y2) | reinterpret_cast<uintptr_t>(y3)) & 15) == 0,
               SLAK_ERR_BAD_ARG, "outputs must be 16-byte aligned");
  CUtensorMap in;
  memset(&in, 0, sizeof(in));
  int rc = make_map(&in, x, false, N, C, P.P, P.G, 64, true);
  if (rc) return rc;
  int grid = sm_count();
  if (grid > P.units) grid = P.units;
  auto kern = dense_kernel<false>;
  SLAK_SET_MAX_SMEM(kern, kSmem);
  kern<<<grid, kThreads, kSmem, st>>>(in, in, in, in, in, in, in, P);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

int dgrad(const void* dy1, const void* dy2, const void* dy3, const float* w1, const float* w2, const float* w3, const float* addend,
          float* dx, int N, int C, int H, int W, int KL, cudaStream_t st) {
  SLAK_REQUIRE(supported(N, C, H, W, KL), SLAK_ERR_UNSUPPORTED, "dense plane kernel: shape not covered");
  SLAK_REQUIRE(addend != nullptr && ((reinterpret_cast<uintptr_t>(addend) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0,
               SLAK_ERR_BAD_ARG, "addend and dx must be 16-byte aligned fp32 tensors");
  Params P{};
  fill(&P, N, C, H, W, KL);
  P.w[0] = w1; P.w[1] = w2; P.w[2] = w3;
  P.addend = addend; P.dx = dx;
  CUtensorMap in[3], fa[2], fd[2];
  memset(in, 0, sizeof(in)); memset(fa, 0, sizeof(fa)); memset(fd, 0, sizeof(fd));
  const void* dys[3] = {dy1, dy2, dy3};
  int rc;
  for (int b = 0; b < 3; ++b) {
    SLAK_REQUIRE((reinterpret_cast<uintptr_t>(dys[b]) & 15) == 0, SLAK_ERR_BAD_ARG, "inputs must be 16-byte aligned");
    if ((rc = make_map(&in[b], dys[b], false, N, C, P.P, P.G, 64, true))) return rc;
  }
  if ((P.P & 3) == 0) {                                       // fp32 rows of P * 4 bytes are 16-byte aligned: TMA boxes
    const int rem4 = P.P & 31;
    if ((rc = make_map(&fa[0], addend, true, N, C, P.P, 1, 32 < P.P ? 32 : P.P, true))) return rc;
    if ((rc = make_map(&fd[0], dx, true, N, C, P.P, 1, 32 < P.P ? 32 : P.P, true))) return rc;
    if ((rc = make_map(&fa[1], addend, true, N, C, P.P, 1, rem4 ? rem4 : 4, false))) return rc;
    if ((rc = make_map(&fd[1], dx, true, N, C, P.P, 1, rem4 ? rem4 : 4, false))) return rc;
  } else {
    fa[0] = fa[1] = fd[0] = fd[1] = in[0];                    // unused
  }
  int grid = sm_count();
  if (grid > P.units) grid = P.units;
  auto kern = dense_kernel<true>;
  SLAK_SET_MAX_SMEM(kern, kSmem);
  kern<<<grid, kThreads, kSmem, st>>>(in[0], in[1], in[2], fa[0], fa[1], fd[0], fd[1], P);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}


bool wgrad_supported(int N, int C, int H, int W, int KL) {
  return supported(N, C, H, W, KL) && N <= 128 && H * W <= 200;
}
int wgrad(const void* x, const void* dy1, const void* dy2, const void* dy3, float* dw1, float* dw2, float* dw3, int N, int C, int H,
          int W, int KL, cudaStream_t st) {
  SLAK_REQUIRE(wgrad_supported(N, C, H, W, KL), SLAK_ERR_UNSUPPORTED, "dense plane wgrad: shape not covered");
  Params Q{};
  fill(&Q, N, C, H, W, KL);
  int dmax = 0;
  for (int j = 0; j < Q.G; ++j) dmax = ((j * Q.P) & 7) > dmax ? ((j * Q.P) & 7) : dmax;
  WgradParams P{};
  P.dw[0] = dw1; P.dw[1] = dw2; P.dw[2] = dw3;
  P.N = N; P.C = C; P.H = H; P.W = W; P.KL = KL; P.P = Q.P; P.G = Q.G;
  P.npadw = (Q.P + dmax + 15) / 16 * 16;
  P.pxblocks = (Q.P + dmax + 63) / 64;
  P.mtiles = (Q.P + dmax + 127) / 128;
  CUtensorMap m[4];
  memset(m, 0, sizeof(m));
  const void* ts[4] = {x, dy1, dy2, dy3};
  int rc;
  for (int k = 0; k < 4; ++k) {
    SLAK_REQUIRE((reinterpret_cast<uintptr_t>(ts[k]) & 15) == 0, SLAK_ERR_BAD_ARG, "inputs must be 16-byte aligned");
    if ((rc = make_map(&m[k], ts[k], false, N, C, Q.P, Q.G, 64, true))) return rc;
  }
  int grid = sm_count();
  if (grid > C) grid = C;
  SLAK_SET_MAX_SMEM(dense_wgrad_kernel, wg::kSmemW);
  dense_wgrad_kernel<<<grid, kThreads, wg::kSmemW, st>>>(m[0], m[1], m[2], m[3], P);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

}  // namespace dense
}  // namespace tc
}  // namespace slak
