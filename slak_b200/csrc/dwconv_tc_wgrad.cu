// Fused three-branch depthwise weight gradient on the tensor cores (bf16 in, fp32 accumulate in
// TMEM for the WHOLE batch slice of a CTA, fp32 out):
//   dw1[t,s] = sum_{n,p,q} dy1[n,p,q] * x[n,p+t-pad,q+s-2]      (KL x 5)
//   dw2[r,t] = sum_{n,p,q} dy2[n,p,q] * x[n,p+r-2,q+t-pad]      (5 x KL)
//   dw3[r,t] = sum_{n,p,q} dy3[n,p,q] * x[n,p+r-2,q+t-2]        (5 x 5)
// (backward_filter of the three convs of ReparamLargeKernelConv, backward_filter_fp32.cu:199-263).
//
// Formulation: per channel, the plane-vs-plane correlation matrices
//   G_s[h,p]      = sum_{n,q} x[n,h,q+s-2] * dy1[n,p,q]          -> dw1[t,s] = sum_p G_s[p+t-pad, p]
//   D_r[(b,q),w]  = sum_{n,p} dy_b[n,p,q] * x[n,p+r-2,w], b=2,3 -> dw2[r,t] = sum_q D_r[(2,q), q+t-pad]
// are GEMMs whose contraction runs over the ROWS of 64-row smem tiles, i.e. both operands are
// MN-major SWIZZLE_128B tiles, and the 5-tap shift is a row offset of the descriptor start address.
// A tile holds 64/T planes of the channel stacked along its rows (T = 64/32/16 tile class; the zero
// rows of each T x T plane block are the padding the row shift runs into).  M=128 is filled by two
// MN atoms LBO bytes apart:
//   D_r : atoms = the dy2 tile and the dy3 tile                     (LBO = 8 KB)
//   G_s : atoms = the x^T tile shifted by s and by s+1 rows         (LBO = 128 B, overlapping)
// G_s needs x^T and dy1^T, made in shared memory by four transposer warps.  Accumulators live in
// TMEM across all planes of the CTA (3 x 64 + 5 x 64 = 512 columns) and are read once at the end;
// the diagonal sums are done from shared memory and written as per-CTA partials, reduced in a
// fixed order by wgrad3_reduce_kernel (deterministic; the reference scatters with atomicAdd).
#include "common.cuh"
#include "tc_common.cuh"
#include <string.h>

namespace slak {
namespace tc {

int make_plane_map(CUtensorMap* map, const void* x, int N, int C, int H, int W);  // dwconv_tc_fwd.cu

namespace wg {
constexpr int kStages = 4;                     // input stages (x, dy1, dy2, dy3 per stage)
constexpr int kTStages = 2;                    // transposed stages (x^T, dy1^T)
constexpr int kTile = 64 * 128;                // 8 KB
constexpr int kPad = 1024;
// stage layout: [pad][XN][pad][D2][D3][D1N]
constexpr int kOffXNs = kPad;
constexpr int kOffD2s = kOffXNs + kTile + kPad;
constexpr int kOffD3s = kOffD2s + kTile;
constexpr int kOffD1s = kOffD3s + kTile;
constexpr int kStageBytes = kOffD1s + kTile;   // 34 KB
// transposed stage: [pad][XT][pad][D1T]
constexpr int kOffXTs = kPad;
constexpr int kOffD1Ts = kOffXTs + kTile + kPad;
constexpr int kTStageBytes = kOffD1Ts + kTile; // 18 KB
constexpr int kOffT = kStages * kStageBytes;
constexpr int kOffBar = kOffT + kTStages * kTStageBytes;
constexpr int kSmemBytes = kOffBar + 256 + 1024;
constexpr int kNumTransposerWarps = 4;
constexpr int kScratchPitch = 65;              // floats, epilogue scratch [128][65]
constexpr int kThreads = 384;                  // w0 loader, w1 MMA, w2-5 transposers, w4-7 epilogue, w8-11 loaders
static_assert(128 * kScratchPitch * 4 <= kStageBytes, "epilogue scratch must fit one stage");
}  // namespace wg

struct WgradParams {
  const __nv_bfloat16* x; const __nv_bfloat16* dy1; const __nv_bfloat16* dy2; const __nv_bfloat16* dy3;
  float* pw1; float* pw2; float* pw3;          // partials [S][C][KL*5], [S][C][5*KL], [S][C][25]
  int N, C, H, W, KL, splits, units_per_c;
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// cp.async of PPT stacked planes of one tensor into a 64-row SWIZZLE_128B tile
template <int T, int CB>
__device__ __forceinline__ void load_tile_pieces(const __nv_bfloat16* __restrict__ t, uint32_t tile, int n0, int c,
                                                 int N, int C, int H, int W, int lane) {
  constexpr int PPT = 64 / T;
  const int PR = (W * 2) / CB;
  const int per_plane = H * PR;
  const size_t plane_bytes = (size_t)H * W * 2;
  for (int pl = 0; pl < PPT; ++pl) {
    const int n = n0 + pl;
    if (n >= N) break;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(t) + ((size_t)n * C + c) * plane_bytes;
    for (int e = lane; e < per_plane; e += 32) {
      const int p = e / PR, j = e - p * PR;
      const int row = pl * T + p;
      const int b = j * CB;
      const uint32_t dst = tile + row * 128 + ((((b >> 4) ^ (row & 7))) << 4) + (b & 15);
      const uint8_t* s = src + (size_t)p * W * 2 + b;
      if constexpr (CB >= 4) {
        cp_async<CB>(dst, s);
      } else {
        const uint16_t val = *reinterpret_cast<const uint16_t*>(s);
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(dst), "h"(val) : "memory");
      }
    }
  }
}
// zero the plane blocks of a tile that have no plane behind them (tail unit of the batch slice)
template <int T>
__device__ __forceinline__ void zero_missing_planes(uint8_t* tile, int n0, int n_end, int lane) {
  constexpr int PPT = 64 / T;
  for (int pl = 0; pl < PPT; ++pl)
    if (n0 + pl >= n_end)
      for (int i = lane; i < T * 128 / 16; i += 32) reinterpret_cast<uint4*>(tile + pl * T * 128)[i] = make_uint4(0, 0, 0, 0);
}

template <int T, int CB, bool TMA>
__global__ void __launch_bounds__(wg::kThreads, 1)
lk3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap d1map,
                    const __grid_constant__ CUtensorMap d2map, const __grid_constant__ CUtensorMap d3map,
                    WgradParams P) {
  using namespace wg;
  constexpr int PPT = 64 / T;                  // planes per tile (stacked along the rows)
  constexpr int KSTEPS1 = T / 16;              // k-steps of the G_s GEMMs (contraction over q < T)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x / P.splits, split = blockIdx.x % P.splits;
  const int u_begin = (int)(((long long)P.units_per_c * split) / P.splits);
  const int u_end = (int)(((long long)P.units_per_c * (split + 1)) / P.splits);
  const int n_units = u_end - u_begin;
  const int KL = P.KL, pad = KL / 2, H = P.H, W = P.W;

  constexpr int B_FULL = 0, B_EMPTY = kStages, B_TFULL = 2 * kStages, B_TEMPTY = B_TFULL + kTStages,
                B_DONE = B_TEMPTY + kTStages;
  const uint32_t bar0 = base + kOffBar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kOffBar + 192);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(BAR(B_FULL + s), 1);
      mbar_init(BAR(B_EMPTY + s), 1 + kNumTransposerWarps);
    }
    for (int s = 0; s < kTStages; ++s) {
      mbar_init(BAR(B_TFULL + s), kNumTransposerWarps);
      mbar_init(BAR(B_TEMPTY + s), 1);
    }
    mbar_init(BAR(B_DONE), 1);
    mbar_fence_init();
    if (TMA) { tma_prefetch_desc(&xmap); tma_prefetch_desc(&d1map); tma_prefetch_desc(&d2map); tma_prefetch_desc(&d3map); }
  }
  {  // everything starts as zeros: pads and tile padding are never written afterwards
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < kOffBar / 16; i += kThreads) reinterpret_cast<uint4*>(sm)[i] = z;
  }
  fence_proxy_async();
  if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const bool is_loader = (warp == 0) || (!TMA && warp >= 8);
  if (is_loader) {
    if constexpr (TMA) {
      if (elect_one()) {
        for (int i = 0; i < n_units; ++i) {
          const int st = i % kStages, ph = (i / kStages) & 1;
          mbar_wait(BAR(B_EMPTY + st), ph ^ 1);
          const uint32_t sb = base + st * kStageBytes;
          const int plane = (u_begin + i) * P.C + c;   // PPT == 1
          mbar_expect_tx(BAR(B_FULL + st), 4 * kTile);
          tma_load_3d(sb + kOffXNs, &xmap, BAR(B_FULL + st), 0, 0, plane);
          tma_load_3d(sb + kOffD2s, &d2map, BAR(B_FULL + st), 0, 0, plane);
          tma_load_3d(sb + kOffD3s, &d3map, BAR(B_FULL + st), 0, 0, plane);
          tma_load_3d(sb + kOffD1s, &d1map, BAR(B_FULL + st), 0, 0, plane);
        }
      }
    } else {
      // four cp.async loader warps, loader j owns stage j
      const int lj = (warp == 0) ? 0 : (warp - 7);           // 0..3
      if (lj < kStages) {
        for (int i = lj; i < n_units; i += kStages) {
          const int st = lj, ph = (i / kStages) & 1;
          mbar_wait(BAR(B_EMPTY + st), ph ^ 1);
          const uint32_t sb = base + st * kStageBytes;
          const int n0 = PPT * (u_begin + i);
          if (n0 + PPT > P.N) {                               // tail: stale planes must not contribute
            zero_missing_planes<T>(sm + st * kStageBytes + kOffD2s, n0, P.N, lane);
            zero_missing_planes<T>(sm + st * kStageBytes + kOffD3s, n0, P.N, lane);
            zero_missing_planes<T>(sm + st * kStageBytes + kOffD1s, n0, P.N, lane);
          }
          load_tile_pieces<T, CB>(P.x, sb + kOffXNs, n0, c, P.N, P.C, H, W, lane);
          load_tile_pieces<T, CB>(P.dy2, sb + kOffD2s, n0, c, P.N, P.C, H, W, lane);
          load_tile_pieces<T, CB>(P.dy3, sb + kOffD3s, n0, c, P.N, P.C, H, W, lane);
          load_tile_pieces<T, CB>(P.dy1, sb + kOffD1s, n0, c, P.N, P.C, H, W, lane);
          cp_async_commit();
          cp_async_wait_all();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_FULL + st));
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16_mn(128, 64);
      for (int i = 0; i < n_units; ++i) {
        const int st = i % kStages, ph = (i / kStages) & 1;
        const int ts = i % kTStages, tph = (i / kTStages) & 1;
        mbar_wait(BAR(B_FULL + st), ph);
        tc_fence_after();
        const uint32_t sb = base + st * kStageBytes;
        // D_r[(b,q), w] += [dy2 | dy3](row, q) * x(row + r - 2, w), rows = stacked (plane, p)
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_desc_mn_sw128(sb + kOffD2s + k * 2048, kTile);
            const uint64_t db = umma_desc_mn_sw128(sb + kOffXNs + (r - 2) * 128 + k * 2048, 0);
            umma_bf16(tmem + 192 + 64 * r, da, db, idesc, (i | k) != 0);
          }
        umma_commit(BAR(B_EMPTY + st));
        mbar_wait(BAR(B_TFULL + ts), tph);
        tc_fence_after();
        const uint32_t tb = base + kOffT + ts * kTStageBytes;
        // G_{2j+tl}[(tl,(plane,h)), (plane',p)] += x^T(q + 2j + tl - 2, (plane,h)) * dy1^T(q, (plane',p))
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int k = 0; k < KSTEPS1; ++k) {
            const uint64_t da = umma_desc_mn_sw128(tb + kOffXTs + (2 * j - 2) * 128 + k * 2048, 128);
            const uint64_t db = umma_desc_mn_sw128(tb + kOffD1Ts + k * 2048, 0);
            umma_bf16(tmem + 64 * j, da, db, idesc, (i | k) != 0);
          }
        umma_commit(BAR(B_TEMPTY + ts));
      }
      umma_commit(BAR(B_DONE));
    }
  } else if (warp < 2 + kNumTransposerWarps) {
    // ================= transposers: x -> x^T, dy1 -> dy1^T (only the T leading columns matter) ========
    const int tw = warp - 2;
    const int m = lane >> 3, kk = lane & 7;
    constexpr int NBJ = T / 8;                  // column blocks that hold data
    constexpr int GROUPS = 2 * 8 * NBJ / 4;     // x4 groups per unit over both tensors
    for (int i = 0; i < n_units; ++i) {
      const int st = i % kStages, ph = (i / kStages) & 1;
      const int ts = i % kTStages, tph = (i / kTStages) & 1;
      mbar_wait(BAR(B_FULL + st), ph);
      mbar_wait(BAR(B_TEMPTY + ts), tph ^ 1);
      const uint32_t sb = base + st * kStageBytes;
      const uint32_t tb = base + kOffT + ts * kTStageBytes;
#pragma unroll 4
      for (int it = tw; it < GROUPS; it += kNumTransposerWarps) {
        const int blk = 4 * it + m;                         // (which, bi, bj)
        const int which = blk / (8 * NBJ), rem = blk - which * (8 * NBJ);
        const int bi = rem / NBJ, bj = rem - bi * NBJ;
        const uint32_t src0 = sb + (which ? kOffD1s : kOffXNs);
        const uint32_t dst0 = tb + (which ? kOffD1Ts : kOffXTs);
        const uint32_t src = src0 + (8 * bi + kk) * 128 + ((bj ^ kk) << 4);
        const uint32_t dst = dst0 + (8 * bj + kk) * 128 + ((bi ^ kk) << 4);
        uint32_t r0, r1, r2, r3;
        ldmatrix_x4_trans(src, r0, r1, r2, r3);
        stmatrix_x4(dst, r0, r1, r2, r3);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(BAR(B_TFULL + ts));
        mbar_arrive(BAR(B_EMPTY + st));
      }
    }
  }

  // ================= epilogue: warps 4..7 read TMEM (lane quarters 0..3), everyone sums diagonals ======
  if (warp >= 4 && warp < 8) {
    mbar_wait(BAR(B_DONE), 0);
    tc_fence_after();
  }
  float* scr = reinterpret_cast<float*>(sm);   // [128][65] fp32 in stage 0 (all MMAs complete once B_DONE fired)
  const int taps1 = KL * 5;
  float* o1 = P.pw1 + ((size_t)split * P.C + c) * taps1;
  float* o2 = P.pw2 + ((size_t)split * P.C + c) * taps1;
  float* o3 = P.pw3 + ((size_t)split * P.C + c) * 25;
  for (int acc = 0; acc < 8; ++acc) {
    __syncthreads();                       // previous round's readers are done with the scratch
    if (warp >= 4 && warp < 8) {
      const int e = warp - 4;
      const int L = e * 32 + lane;
      uint32_t v[64];
      const uint32_t t0 = tmem + ((uint32_t)(e * 32) << 16) + 64 * acc;
      tmem_ld32(t0, v);
      tmem_ld32(t0 + 32, v + 32);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 64; ++j) scr[L * kScratchPitch + j] = __uint_as_float(v[j]);
    }
    __syncthreads();
    if (acc < 3) {
      // G_s, s = 2*acc (lanes 0..63) and 2*acc+1 (lanes 64..127): dw1[t,s] = sum_{plane,p} G_s[(plane,p+t-pad)][(plane,p)]
      for (int idx = tid; idx < 2 * KL; idx += kThreads) {
        const int tl = idx / KL, t = idx - tl * KL;
        const int s = 2 * acc + tl;
        if (s < 5) {
          float a = 0.f;
          for (int pl = 0; pl < PPT; ++pl)
            for (int p = 0; p < H; ++p) {
              const int h = p + t - pad;
              if (h >= 0 && h < H) a += scr[(tl * 64 + pl * T + h) * kScratchPitch + pl * T + p];
            }
          o1[t * 5 + s] = a;
        }
      }
    } else {
      const int r = acc - 3;
      // dw2[r,t] = sum_q D_r[q][q+t-pad] ; dw3[r,t'] = sum_q D_r[64+q][q+t'-2]
      for (int idx = tid; idx < KL + 5; idx += kThreads) {
        const bool is3 = idx >= KL;
        const int t = is3 ? idx - KL : idx;
        const int off = is3 ? 2 : pad;
        float a = 0.f;
        for (int q = 0; q < W; ++q) {
          const int w = q + t - off;
          if (w >= 0 && w < W) a += scr[((is3 ? 64 : 0) + q) * kScratchPitch + w];
        }
        if (is3) o3[r * 5 + t] = a; else o2[r * KL + t] = a;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

// partial[S][total] -> out[total] in a fixed order
__global__ void wgrad3_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int S, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float v = 0.f;
  for (int s = 0; s < S; ++s) v += partial[(size_t)s * total + i];
  out[i] = v;
}

static int wgrad_units(int N, int tile) { const int ppt = 64 / tile; return (N + ppt - 1) / ppt; }

size_t lk3_wgrad_tc_workspace(int N, int C, int H, int W, int KL) {
  const TcShape s = tc_shape(H, W);
  if (s.tile == 0) return 0;
  const int S = tc_pick_splits(C, wgrad_units(N, s.tile));
  return (size_t)S * C * (2 * KL * 5 + 25) * sizeof(float);
}

template <int T, int CB, bool TMA>
static int launch_wgrad(const CUtensorMap* maps, WgradParams& P, cudaStream_t st) {
  auto kern = lk3_wgrad_tc_kernel<T, CB, TMA>;
  SLAK_SET_MAX_SMEM(kern, wg::kSmemBytes);
  kern<<<P.C * P.splits, wg::kThreads, wg::kSmemBytes, st>>>(maps[0], maps[1], maps[2], maps[3], P);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

int lk3_wgrad_tc(const void* x, const void* dy1, const void* dy2, const void* dy3, float* dw1, float* dw2,
                 float* dw3, int N, int C, int H, int W, int KL, void* workspace, cudaStream_t st) {
  const TcShape s = tc_shape(H, W);
  SLAK_REQUIRE(s.tile != 0, SLAK_ERR_UNSUPPORTED, "shape %dx%d not covered by the tensor-core path", H, W);
  CUtensorMap maps[4];
  memset(maps, 0, sizeof(maps));
  if (s.tma) {
    int rc;
    if ((rc = make_plane_map(&maps[0], x, N, C, H, W))) return rc;
    if ((rc = make_plane_map(&maps[1], dy1, N, C, H, W))) return rc;
    if ((rc = make_plane_map(&maps[2], dy2, N, C, H, W))) return rc;
    if ((rc = make_plane_map(&maps[3], dy3, N, C, H, W))) return rc;
  }
  WgradParams P;
  P.x = (const __nv_bfloat16*)x; P.dy1 = (const __nv_bfloat16*)dy1;
  P.dy2 = (const __nv_bfloat16*)dy2; P.dy3 = (const __nv_bfloat16*)dy3;
  P.N = N; P.C = C; P.H = H; P.W = W; P.KL = KL;
  P.units_per_c = wgrad_units(N, s.tile);
  P.splits = tc_pick_splits(C, P.units_per_c);
  const size_t t1 = (size_t)C * KL * 5, t3 = (size_t)C * 25;
  float* ws = (float*)workspace;
  P.pw1 = ws;
  P.pw2 = ws + (size_t)P.splits * t1;
  P.pw3 = ws + 2 * (size_t)P.splits * t1;
  int rc;
  if (s.tile == 64) rc = launch_wgrad<64, 16, true>(maps, P, st);
  else if (s.tile == 32) rc = s.cb == 8 ? launch_wgrad<32, 8, false>(maps, P, st)
                            : s.cb == 4 ? launch_wgrad<32, 4, false>(maps, P, st)
                                        : launch_wgrad<32, 2, false>(maps, P, st);
  else rc = s.cb == 4 ? launch_wgrad<16, 4, false>(maps, P, st) : launch_wgrad<16, 2, false>(maps, P, st);
  if (rc) return rc;
  wgrad3_reduce_kernel<<<(int)((t1 + 255) / 256), 256, 0, st>>>(P.pw1, dw1, P.splits, (int)t1);
  wgrad3_reduce_kernel<<<(int)((t1 + 255) / 256), 256, 0, st>>>(P.pw2, dw2, P.splits, (int)t1);
  wgrad3_reduce_kernel<<<(int)((t3 + 255) / 256), 256, 0, st>>>(P.pw3, dw3, P.splits, (int)t3);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

}  // namespace tc
}  // namespace slak
