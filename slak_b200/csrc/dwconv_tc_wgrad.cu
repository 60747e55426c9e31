// Fused three-branch depthwise weight gradient on the tensor cores (bf16 in, fp32 accumulate in
// TMEM over a channel's whole batch slice, fp32 out):
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
// A unit is one 64 x 64 tile per tensor holding (64/T)^2 planes of the channel: 64/T stacked along the
// rows (they add up in the contraction) x 64/T side by side in T-column bands (the MMA also forms the
// cross-band products; only the diagonal T x T blocks are read back).  The zero rows / columns of each
// T x T plane block are the padding the row shift runs into.  M=128 is filled by two MN atoms LBO
// bytes apart:
//   D_r : atoms = the dy2 tile and the dy3 tile                     (LBO = 8 KB)
//   G_s : atoms = the x^T tile shifted by s and by s+1 rows         (LBO = 128 B, overlapping)
// G_s needs x^T and dy1^T, made in shared memory by four transposer warps.  The 8 accumulators
// (3 x 64 + 5 x 64 = 512 TMEM columns) stay in TMEM over all units of a channel; at the channel's end
// four epilogue warps copy them to shared memory, take the diagonal sums and write per-(channel, CTA)
// partials that wgrad3_reduce_kernel adds in a fixed order (deterministic; the reference scatters
// with atomicAdd).  CTAs are persistent over (channel, unit) items like the forward kernel.
//
// Warp roles (416 threads): w0 loader | w1 MMA | w2-5 transposers (w2 owns TMEM) | w6-9 epilogue |
// w10-12 extra loaders (cp.async path).
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
constexpr int kScratchPitch = 65;              // floats, epilogue scratch [128][65]
constexpr int kOffScr = kOffT + kTStages * kTStageBytes;
constexpr int kOffBar = kOffScr + ((128 * kScratchPitch * 4 + 1023) / 1024) * 1024;
constexpr int kSmemBytes = kOffBar + 1024 + 1024;
constexpr int kNumTransposerWarps = 4;
// cp.async classes: loader warps per stage.  4-byte cp.async is bound by per-warp latency, so the 16-class (16 planes
// x 4 tensors per unit) runs two warps per stage, each taking half of the planes (w13-16 join w0, w10-12)
__host__ __device__ constexpr int loader_split(int T) { return T == 16 ? 2 : 1; }
__host__ __device__ constexpr int threads(int T) { return T == 16 ? 544 : 416; }
}  // namespace wg

struct WgradParams {
  const __nv_bfloat16* x; const __nv_bfloat16* dy1; const __nv_bfloat16* dy2; const __nv_bfloat16* dy3;
  float* pw1; float* pw2; float* pw3;          // partials [S][C][KL*5], [S][C][5*KL], [S][C][25]
  int N, C, H, W, KL, splits, units_per_c, per_cta;
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

// zero a T x T plane block of a tile (tail unit: stale planes must not contribute)
template <int T>
__device__ __forceinline__ void zero_plane_block(uint8_t* tile, int row0, int cb0, int lane) {
  constexpr int CH = T / 8;
  for (int i = lane; i < T * CH; i += 32) {
    const int r = row0 + i / CH, ck = cb0 + (i % CH);
    *reinterpret_cast<uint4*>(tile + r * 128 + ((ck ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
  }
}

template <int T, int CB, bool TMA>
__global__ void __launch_bounds__(wg::threads(T), 1)
lk3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap d1map,
                    const __grid_constant__ CUtensorMap d2map, const __grid_constant__ CUtensorMap d3map,
                    WgradParams P) {
  using namespace wg;
  constexpr int PPT = 64 / T;                  // plane blocks per tile edge
  constexpr int PLANES = PPT * PPT;            // planes per unit (1, 4, 16)
  constexpr int kThreads = threads(T);
  constexpr int kSplit = loader_split(T);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int upc = P.units_per_c;
  long long g0, g1;
  if (T == 64) {
    const int c = blockIdx.x / P.splits, split = blockIdx.x % P.splits;
    g0 = (long long)c * upc + ((long long)upc * split) / P.splits;
    g1 = (long long)c * upc + ((long long)upc * (split + 1)) / P.splits;
  } else {
    const long long total = (long long)P.C * upc;
    g0 = (long long)blockIdx.x * P.per_cta;
    g1 = g0 + P.per_cta < total ? g0 + P.per_cta : total;
    if (g0 > total) g0 = total;
  }
  const int n_units = (int)(g1 - g0);
  const int KL = P.KL, pad = KL / 2, H = P.H, W = P.W;

  constexpr int B_FULL = 0, B_EMPTY = kStages, B_TFULL = 2 * kStages, B_TEMPTY = B_TFULL + kTStages,
                B_ACC_FULL = B_TEMPTY + kTStages, B_ACC_EMPTY = B_ACC_FULL + 1;
  const uint32_t bar0 = base + kOffBar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kOffBar + 768);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(BAR(B_FULL + s), TMA ? 1 : kSplit);
      mbar_init(BAR(B_EMPTY + s), 1 + kNumTransposerWarps);
    }
    for (int s = 0; s < kTStages; ++s) {
      mbar_init(BAR(B_TFULL + s), kNumTransposerWarps);
      mbar_init(BAR(B_TEMPTY + s), 1);
    }
    mbar_init(BAR(B_ACC_FULL), 1);             // MMA commit at a channel's last unit
    mbar_init(BAR(B_ACC_EMPTY), 4);            // the four epilogue warps drained TMEM
    mbar_fence_init();
    if (TMA) { tma_prefetch_desc(&xmap); tma_prefetch_desc(&d1map); tma_prefetch_desc(&d2map); tma_prefetch_desc(&d3map); }
  }
  {  // everything starts as zeros: pads and tile padding are never written afterwards
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < kOffScr / 16; i += kThreads) reinterpret_cast<uint4*>(sm)[i] = z;
  }
  fence_proxy_async();
  if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const bool is_loader = (warp == 0) || (!TMA && warp >= 10);     // w13+: second warp of a stage (16-class only)
  if (is_loader) {
    if constexpr (TMA) {
      if (elect_one()) {
        for (int i = 0; i < n_units; ++i) {
          const long long g = g0 + i;
          const int c = (int)(g / upc), u = (int)(g - (long long)c * upc);
          const int st = i % kStages, ph = (i / kStages) & 1;
          mbar_wait(BAR(B_EMPTY + st), ph ^ 1);
          const uint32_t sb = base + st * kStageBytes;
          const int plane = u * P.C + c;           // PLANES == 1
          mbar_expect_tx(BAR(B_FULL + st), 4 * kTile);
          tma_load_3d(sb + kOffXNs, &xmap, BAR(B_FULL + st), 0, 0, plane);
          tma_load_3d(sb + kOffD2s, &d2map, BAR(B_FULL + st), 0, 0, plane);
          tma_load_3d(sb + kOffD3s, &d3map, BAR(B_FULL + st), 0, 0, plane);
          tma_load_3d(sb + kOffD1s, &d1map, BAR(B_FULL + st), 0, 0, plane);
        }
      }
    } else {
      // four cp.async loader warps, loader j owns stage j
      const int lj = (warp == 0) ? 0 : (warp < 13 ? warp - 9 : warp - 13);   // stage 0..3
      const int hf = warp >= 13 ? 1 : 0;
      constexpr int QN = PLANES / kSplit;
      const int qlo = hf * QN, qhi = qlo + QN;
      PieceMap<CB> pm;
      pm.init(H, W, lane);
      const size_t plane_bytes = (size_t)H * W * 2;
      for (int i = lj; i < n_units; i += kStages) {
        const long long g = g0 + i;
        const int c = (int)(g / upc), u = (int)(g - (long long)c * upc);
        const int st = lj, ph = (i / kStages) & 1;
        mbar_wait(BAR(B_EMPTY + st), ph ^ 1);
        const uint32_t sb = base + st * kStageBytes;
        uint8_t* sbp = sm + st * kStageBytes;
        const int n0 = PLANES * u;
        for (int q0 = qlo; q0 < qhi; q0 += 4) {
          if (CB == 2 && pm.count >= 0 && pm.count <= 2) {
            const uint8_t* sx[4]; const uint8_t* s1[4]; const uint8_t* s2[4]; const uint8_t* s3[4]; int r0s[4], c0s[4];
            int cnt = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int q = q0 + j;
              sx[j] = s1[j] = s2[j] = s3[j] = reinterpret_cast<const uint8_t*>(P.x); r0s[j] = 0; c0s[j] = 0;
              if (q < qhi && n0 + q < P.N) {
                const size_t off = ((size_t)(n0 + q) * P.C + c) * plane_bytes;
                sx[j] = reinterpret_cast<const uint8_t*>(P.x) + off; s1[j] = reinterpret_cast<const uint8_t*>(P.dy1) + off;
                s2[j] = reinterpret_cast<const uint8_t*>(P.dy2) + off; s3[j] = reinterpret_cast<const uint8_t*>(P.dy3) + off;
                r0s[j] = (q % PPT) * T; c0s[j] = (q / PPT) * (T / 8);
                cnt = j + 1;
              }
            }
            if constexpr (CB == 2) {
              load_plane_blocks_cb2<4>(pm, sx, sb + kOffXNs, r0s, c0s, cnt, lane);
              load_plane_blocks_cb2<4>(pm, s2, sb + kOffD2s, r0s, c0s, cnt, lane);
              load_plane_blocks_cb2<4>(pm, s3, sb + kOffD3s, r0s, c0s, cnt, lane);
              load_plane_blocks_cb2<4>(pm, s1, sb + kOffD1s, r0s, c0s, cnt, lane);
            }
          } else {
            for (int q = q0; q < q0 + 4 && q < qhi; ++q)
              if (n0 + q < P.N) {
                const size_t off = ((size_t)(n0 + q) * P.C + c) * plane_bytes;
                const int r0 = (q % PPT) * T, c0 = (q / PPT) * (T / 8);
                load_plane_block<CB>(pm, reinterpret_cast<const uint8_t*>(P.x) + off, sb + kOffXNs, r0, c0, lane);
                load_plane_block<CB>(pm, reinterpret_cast<const uint8_t*>(P.dy2) + off, sb + kOffD2s, r0, c0, lane);
                load_plane_block<CB>(pm, reinterpret_cast<const uint8_t*>(P.dy3) + off, sb + kOffD3s, r0, c0, lane);
                load_plane_block<CB>(pm, reinterpret_cast<const uint8_t*>(P.dy1) + off, sb + kOffD1s, r0, c0, lane);
              }
          }
        }
        if (n0 + PLANES > P.N)                               // tail: planes that do not exist must read as zero gradients
          for (int q = qlo; q < qhi; ++q)
            if (n0 + q >= P.N) {
              const int r0 = (q % PPT) * T, c0 = (q / PPT) * (T / 8);
              zero_plane_block<T>(sbp + kOffD2s, r0, c0, lane);
              zero_plane_block<T>(sbp + kOffD3s, r0, c0, lane);
              zero_plane_block<T>(sbp + kOffD1s, r0, c0, lane);
            }
        cp_async_commit();
        cp_async_wait_all();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_FULL + st));
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16_mn(128, 64);
      int chan_idx = 0;                              // channels finished by this CTA so far
      for (int i = 0; i < n_units; ++i) {
        const long long g = g0 + i;
        const int c = (int)(g / upc);
        const bool first = (i == 0) || ((int)((g - 1) / upc) != c);
        const bool last = (i == n_units - 1) || ((int)((g + 1) / upc) != c);
        if (first && chan_idx > 0) {                 // the previous channel's accumulators must be drained first
          mbar_wait(BAR(B_ACC_EMPTY), (chan_idx - 1) & 1);
          tc_fence_after();
        }
        const int st = i % kStages, ph = (i / kStages) & 1;
        const int ts = i % kTStages, tph = (i / kTStages) & 1;
        mbar_wait(BAR(B_FULL + st), ph);
        tc_fence_after();
        const uint32_t sb = base + st * kStageBytes;
        // D_r[(b,q), w] += [dy2 | dy3](row, q) * x(row + r - 2, w), rows = stacked (plane, p), columns in bands
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_desc_mn_sw128(sb + kOffD2s + k * 2048, kTile);
            const uint64_t db = umma_desc_mn_sw128(sb + kOffXNs + (r - 2) * 128 + k * 2048, 0);
            umma_bf16(tmem + 192 + 64 * r, da, db, idesc, (!first || k > 0) ? 1u : 0u);
          }
        umma_commit(BAR(B_EMPTY + st));
        mbar_wait(BAR(B_TFULL + ts), tph);
        tc_fence_after();
        const uint32_t tb = base + kOffT + ts * kTStageBytes;
        // G_{2j+tl}[(tl,(plane,h)), (plane',p)] += x^T((band,q) + 2j + tl - 2, (plane,h)) * dy1^T((band,q), (plane',p))
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_desc_mn_sw128(tb + kOffXTs + (2 * j - 2) * 128 + k * 2048, 128);
            const uint64_t db = umma_desc_mn_sw128(tb + kOffD1Ts + k * 2048, 0);
            umma_bf16(tmem + 64 * j, da, db, idesc, (!first || k > 0) ? 1u : 0u);
          }
        umma_commit(BAR(B_TEMPTY + ts));
        if (last) { umma_commit(BAR(B_ACC_FULL)); ++chan_idx; }
      }
    }
  } else if (warp < 2 + kNumTransposerWarps) {
    // ================= transposers: x -> x^T, dy1 -> dy1^T (whole 64 x 64 tiles) ========
    const int tw = warp - 2;
    const int m = lane >> 3, kk = lane & 7;
    constexpr int ITERS = 32 / kNumTransposerWarps;          // 2 tensors x 64 blocks / 4 per group
    // block -> (source, destination) offsets inside a stage are the same for every unit: computed once
    uint32_t soff[ITERS], doff[ITERS];
#pragma unroll
    for (int q = 0; q < ITERS; ++q) {
      const int blk = 4 * (tw + q * kNumTransposerWarps) + m;
      const int which = blk >> 6, rem = blk & 63;
      const int bi = rem >> 3, bj = rem & 7;
      soff[q] = (which ? kOffD1s : kOffXNs) + (8 * bi + kk) * 128 + ((bj ^ kk) << 4);
      doff[q] = (which ? kOffD1Ts : kOffXTs) + (8 * bj + kk) * 128 + ((bi ^ kk) << 4);
    }
    for (int i = 0; i < n_units; ++i) {
      const int st = i % kStages, ph = (i / kStages) & 1;
      const int ts = i % kTStages, tph = (i / kTStages) & 1;
      mbar_wait(BAR(B_FULL + st), ph);
      const uint32_t sb = base + st * kStageBytes;
      const uint32_t tb = base + kOffT + ts * kTStageBytes;
      uint32_t r[ITERS][4];
#pragma unroll
      for (int q = 0; q < ITERS; ++q) ldmatrix_x4_trans(sb + soff[q], r[q][0], r[q][1], r[q][2], r[q][3]);
      mbar_wait(BAR(B_TEMPTY + ts), tph ^ 1);              // the loads above do not depend on the target slot
#pragma unroll
      for (int q = 0; q < ITERS; ++q) stmatrix_x4(tb + doff[q], r[q][0], r[q][1], r[q][2], r[q][3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(BAR(B_TFULL + ts));
        mbar_arrive(BAR(B_EMPTY + st));
      }
    }
  } else if (warp < 10) {
    // ================= epilogue (w6..w9 -> TMEM lane quarters 2,3,0,1): once per channel =================
    const int e = warp & 3;
    const int et = (warp - 6) * 32 + lane;            // 0..127 thread index inside the epilogue group
    const int L = e * 32 + lane;                      // TMEM lane
    float* scr = reinterpret_cast<float*>(sm + kOffScr);   // [128][65]
    const int taps1 = KL * 5;
    int chan_idx = 0;
    for (int i = 0; i < n_units; ++i) {
      const long long g = g0 + i;
      const int c = (int)(g / upc);
      const bool last = (i == n_units - 1) || ((int)((g + 1) / upc) != c);
      if (!last) continue;
      // partial slot of this CTA inside channel c
      int slot;
      if (T == 64) slot = blockIdx.x % P.splits;
      else slot = (int)(blockIdx.x - ((long long)c * upc) / P.per_cta);
      float* o1 = P.pw1 + ((size_t)slot * P.C + c) * taps1;
      float* o2 = P.pw2 + ((size_t)slot * P.C + c) * taps1;
      float* o3 = P.pw3 + ((size_t)slot * P.C + c) * 25;
      mbar_wait(BAR(B_ACC_FULL), chan_idx & 1);
      tc_fence_after();
      for (int acc = 0; acc < 8; ++acc) {
        uint32_t v[64];
        const uint32_t t0 = tmem + ((uint32_t)(e * 32) << 16) + 64 * acc;
        tmem_ld32(t0, v);
        tmem_ld32(t0 + 32, v + 32);
        tmem_ld_wait();
        if (acc == 7) {                                // TMEM is free for the next channel
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY));
        }
        named_bar_sync(1, 128);                        // previous round's readers are done with the scratch
#pragma unroll
        for (int j = 0; j < 64; ++j) scr[L * kScratchPitch + j] = __uint_as_float(v[j]);
        named_bar_sync(1, 128);
        if (acc < 3) {
          // G_s, s = 2*acc (lanes 0..63) and 2*acc+1 (lanes 64..127):
          //   dw1[t,s] = sum over plane blocks (pl) and p of G_s[(pl, p+t-pad)][(pl, p)]
          for (int idx = et; idx < 2 * KL; idx += 128) {
            const int tl = idx / KL, t = idx - tl * KL;
            const int s = 2 * acc + tl;
            if (s < 5) {
              float a = 0.f;
              for (int blk = 0; blk < PPT; ++blk)
                for (int p = 0; p < H; ++p) {
                  const int h = p + t - pad;
                  if (h >= 0 && h < H) a += scr[(tl * 64 + blk * T + h) * kScratchPitch + blk * T + p];
                }
              o1[t * 5 + s] = a;
            }
          }
        } else {
          const int r = acc - 3;
          // dw2[r,t] = sum over bands and q of D_r[(band,q)][(band, q+t-pad)] ; dw3 likewise from lanes 64..127
          for (int idx = et; idx < KL + 5; idx += 128) {
            const bool is3 = idx >= KL;
            const int t = is3 ? idx - KL : idx;
            const int off = is3 ? 2 : pad;
            float a = 0.f;
            for (int blk = 0; blk < PPT; ++blk)
              for (int q = 0; q < W; ++q) {
                const int w = q + t - off;
                if (w >= 0 && w < W) a += scr[((is3 ? 64 : 0) + blk * T + q) * kScratchPitch + blk * T + w];
              }
            if (is3) o3[r * 5 + t] = a; else o2[r * KL + t] = a;
          }
        }
      }
      ++chan_idx;
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

static TcPlan wgrad_plan(int N, int C, int tile) {
  const int ppt = 64 / tile;
  return tc_plan(N, C, tile, ppt * ppt);
}

size_t lk3_wgrad_tc_workspace(int N, int C, int H, int W, int KL) {
  const TcShape s = tc_shape(H, W);
  if (s.tile == 0) return 0;
  return (size_t)wgrad_plan(N, C, s.tile).splits * C * (2 * KL * 5 + 25) * sizeof(float);
}

template <int T, int CB, bool TMA>
static int launch_wgrad(const CUtensorMap* maps, WgradParams& P, int grid, cudaStream_t st) {
  auto kern = lk3_wgrad_tc_kernel<T, CB, TMA>;
  SLAK_SET_MAX_SMEM(kern, wg::kSmemBytes);
  kern<<<grid, wg::threads(T), wg::kSmemBytes, st>>>(maps[0], maps[1], maps[2], maps[3], P);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

// dwconv_tc_dense.cu: planes of up to 200 pixels, batches of up to 128 images: dense GEMMs + diagonal sums
namespace dense {
bool wgrad_supported(int N, int C, int H, int W, int KL);
int wgrad(const void* x, const void* dy1, const void* dy2, const void* dy3, float* dw1, float* dw2, float* dw3, int N, int C, int H,
          int W, int KL, cudaStream_t st);
}

int lk3_wgrad_tc(const void* x, const void* dy1, const void* dy2, const void* dy3, float* dw1, float* dw2,
                 float* dw3, int N, int C, int H, int W, int KL, void* workspace, cudaStream_t st) {
  if (dense::wgrad_supported(N, C, H, W, KL)) return dense::wgrad(x, dy1, dy2, dy3, dw1, dw2, dw3, N, C, H, W, KL, st);
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
  const TcPlan plan = wgrad_plan(N, C, s.tile);
  WgradParams P;
  P.x = (const __nv_bfloat16*)x; P.dy1 = (const __nv_bfloat16*)dy1;
  P.dy2 = (const __nv_bfloat16*)dy2; P.dy3 = (const __nv_bfloat16*)dy3;
  P.N = N; P.C = C; P.H = H; P.W = W; P.KL = KL;
  P.units_per_c = plan.units_per_c;
  P.per_cta = plan.per_cta;
  P.splits = plan.splits;
  const size_t t1 = (size_t)C * KL * 5, t3 = (size_t)C * 25;
  float* ws = (float*)workspace;
  P.pw1 = ws;
  P.pw2 = ws + (size_t)P.splits * t1;
  P.pw3 = ws + 2 * (size_t)P.splits * t1;
  // not every (slot, channel) partial is written in the multi-channel partition
  SLAK_CUDA_TRY(cudaMemsetAsync(ws, 0, (size_t)P.splits * (2 * t1 + t3) * sizeof(float), st));
  int rc;
  if (s.tile == 64) rc = launch_wgrad<64, 16, true>(maps, P, plan.grid, st);
  else if (s.tile == 32) rc = s.cb == 8 ? launch_wgrad<32, 8, false>(maps, P, plan.grid, st)
                            : s.cb == 4 ? launch_wgrad<32, 4, false>(maps, P, plan.grid, st)
                                        : launch_wgrad<32, 2, false>(maps, P, plan.grid, st);
  else rc = s.cb == 4 ? launch_wgrad<16, 4, false>(maps, P, plan.grid, st) : launch_wgrad<16, 2, false>(maps, P, plan.grid, st);
  if (rc) return rc;
  wgrad3_reduce_kernel<<<(int)((t1 + 255) / 256), 256, 0, st>>>(P.pw1, dw1, P.splits, (int)t1);
  wgrad3_reduce_kernel<<<(int)((t1 + 255) / 256), 256, 0, st>>>(P.pw2, dw2, P.splits, (int)t1);
  wgrad3_reduce_kernel<<<(int)((t3 + 255) / 256), 256, 0, st>>>(P.pw3, dw3, P.splits, (int)t3);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

}  // namespace tc
}  // namespace slak
