// Fused three-branch large-kernel depthwise forward on the 5th-gen tensor cores (bf16 in,
// fp32 accumulate in TMEM, bf16 out):
//     y1 = dwconv_{KL x 5}(x)   y2 = dwconv_{5 x KL}(x)   y3 = dwconv_{5 x 5}(x)
// (the three convolutions of ReparamLargeKernelConv.forward, models/SLaK.py:89-100), x read ONCE.
//
// Formulation (banded-Toeplitz GEMMs, per channel):
//   long axis of a branch  -> contraction (K) against a Toeplitz matrix of the taps, built in
//                             shared memory per channel (B operand, K-major, SWIZZLE_128B)
//   short axis (5 taps)    -> five accumulating MMAs whose A operand (the image planes) starts
//                             (t-2) ROWS later: a row shift is a +128-byte descriptor offset
//   y2|y3 [(plane,p), q]  += X [(plane,p+r-2), w]   * [T2_r ; T3_r][q, w]    M=128 N=2T K=T
//   y1^T  [(plane,q), p]  += X^T[(plane,q+s-2), h]  * T1_s[p, h]             M=128 N=T  K=T
//   A plane occupies a T x T block (T = 64, 32 or 16 >= H+2, W+2) that is zero beyond H x W.  A unit is one
//   128-row x 64-column SWIZZLE_128B tile: 128/T row groups x 64/T column bands = 2 / 8 / 32 planes
//   of one channel (~12.5 KB of input whatever the class); the zero rows double as the "same"
//   padding between stacked planes.
//   X is staged by TMA (T=64, W%8==0: 3-D tiled map, OOB zero fill) or by cp.async pieces; X^T is
//   made in shared memory with ldmatrix.trans/stmatrix.
//
// A CTA walks a contiguous range of (channel, unit) work items.  For T=64 a range stays inside one
// channel (one Toeplitz set, 120 KB); for the small classes the CTAs are persistent over many
// channels and a builder warp prepares the next channel's Toeplitz set (double-buffered) while the
// pipeline runs, so short channels do not pay a pipeline drain.
//
// Warp roles: w0 loader | w1 MMA issuer | w2-3 transposers (w2 owns TMEM alloc) | w4-7 and w8-11 two epilogue
// warpgroups, one per accumulator buffer, taking alternate units (TMEM -> registers -> bf16 -> shared staging tile ->
// global: TMA tile stores in the T = 64 class; y1 is transposed on the way) | w12-13 two more transposers (TMA class,
// 448 threads) or cp.async loaders, with w14 the Toeplitz builder and w15-17 three more loaders (small classes, 576
// threads: every X slot is filled by two warps, 4-byte cp.async is bound by per-warp latency).
#include "common.cuh"
#include "tc_common.cuh"
#include <string.h>

// -DSLAK_ROLE_PROFILE: CTA 0 prints, per warp role, the cycles spent in its main loop and the part of them spent
// waiting on mbarriers / named barriers (debug aid for finding the critical role; never defined in the product build)
#ifdef SLAK_ROLE_PROFILE
#define MBW(bar, par) do { const long long _t0 = clock64(); mbar_wait((bar), (par)); prof_wait += clock64() - _t0; } while (0)
#define NBS(id, n) do { const long long _t0 = clock64(); named_bar_sync((id), (n)); prof_wait += clock64() - _t0; } while (0)
#else
#define MBW(bar, par) mbar_wait((bar), (par))
#define NBS(id, n) named_bar_sync((id), (n))
#endif

namespace slak {
namespace tc {

constexpr int kStages = 3;                       // X (natural) slots in flight
constexpr int kAccBufs = 2;
constexpr int kUnitBytes = 128 * 128;            // 128 rows x 64 bf16
constexpr int kPad = 1024;                       // zero rows before/after a unit tile
constexpr int kXSlot = kPad + kUnitBytes;        // 17 KB: [zero pad][tile]; the next slot's pad closes this one
__host__ __device__ constexpr int fwd_transposers(bool tma) { return tma ? 4 : 2; }
constexpr int kEpiGroups = 2;
// cp.async classes: loader warps per X slot (each takes a share of the planes).  4-byte cp.async is bound by per-warp
// latency, so the 16-class (32 planes of <= 392 B per unit) runs two warps per slot; for the 32-class the extra warps
// cost more (register cap of a 576-thread CTA) than they bring
__host__ __device__ constexpr int fwd_loader_split(int T) { return T == 16 ? 2 : 1; }
__host__ __device__ constexpr int fwd_threads(int T, bool tma) { return tma ? 448 : (T == 16 ? 576 : 480); }

template <int T> struct FwdCfg {
  static constexpr int PPU = 128 / T;            // row groups per unit
  static constexpr int KSTEPS = T / 16;
  static constexpr int UPS = 64 / T;             // column bands per unit
  static constexpr int PLANES = PPU * UPS;       // planes per unit (2, 8, 32)
  static constexpr int NT = (T == 64) ? 1 : 2;   // Toeplitz sets (double-buffered for the multi-channel classes)
  static constexpr int kToep1 = 5 * T * 128;
  static constexpr int kToep23 = 5 * 2 * T * 128;
  static constexpr int kToepSet = kToep1 + kToep23;
  static constexpr int kOffToep = 0;
  static constexpr int kOffXN = kOffToep + NT * kToepSet;
  static constexpr int kOffXT = kOffXN + kStages * kXSlot;
  static constexpr int kOffY1 = kOffXT + kXSlot + kPad;      // one y1 staging tile per epilogue group
  // fp32 tap staging (4 KB): the builder warp's own for the multi-channel classes; for T = 64 the one Toeplitz set is
  // built before the pipeline starts, so the staging aliases the (not yet used) y1 tiles
  static constexpr int kOffW = (NT == 1) ? kOffY1 : kOffY1 + kEpiGroups * kUnitBytes;
  static constexpr int kOffBar = kOffY1 + kEpiGroups * kUnitBytes + (NT == 1 ? 0 : 4096);
  static constexpr int kSmem = kOffBar + 1024 + 1024;
  static_assert(kSmem <= 232448, "shared memory budget");
  static constexpr int kAccCols = 3 * T * UPS;   // 192 for every class
  static constexpr int kTmemCols = 512;
};

struct FwdParams {
  const __nv_bfloat16* x;
  const float* w1; const float* w2; const float* w3;       // fp32 taps [C,KL,5] [C,5,KL] [C,5,5]
  __nv_bfloat16* y1; __nv_bfloat16* y2; __nv_bfloat16* y3;
  int N, C, H, W, KL;
  int units_per_c;       // ceil(N / PLANES)
  int per_cta;           // work items per CTA; items are (channel, unit) in channel-major order.  For T=64
                         // per_cta divides the channel into `splits` ranges (grid = C * splits)
  int splits;            // T=64: CTAs per channel; small classes: CTAs that can touch one channel
  float* stats;          // optional [C][splits][2][6]: per (CTA, epilogue group) (sum, sum of squares) of y1, y2, y3
                         // (fp32, before rounding)
};

// Banded Toeplitz operands of one channel (K-major SWIZZLE_128B): five T1_s tiles at tp, then five [T2_r ; T3_r] tiles.
template <int T>
__device__ __forceinline__ void build_toeplitz(uint8_t* tp, const float* w1s, const float* w2s, const float* w3s,
                                               int KL, int pad, int H, int W, int t0, int nthr) {
  using Cfg = FwdCfg<T>;
  constexpr int CH = T / 8;                                  // 16-byte chunks per row that are ever read
  // T1_s[p][h] = w1[h-p+pad][s]
  for (int ch = t0; ch < 5 * T * CH; ch += nthr) {
    const int s = ch / (T * CH), rem = ch - s * (T * CH), p = rem / CH, k8 = rem - p * CH;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = (k8 * 8 + j) - p + pad;
      v[j] = (t >= 0 && t < KL && p < H) ? w1s[t * 5 + s] : 0.f;     // output rows beyond the plane stay exactly zero
    }
    *reinterpret_cast<uint4*>(tp + s * (T * 128) + p * 128 + ((k8 ^ (p & 7)) << 4)) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
  // T23_r rows 0..T-1: T2_r[q][w] = w2[r][w-q+pad] ; rows T..2T-1: T3_r[q][w] = w3[r][w-q+2]
  for (int ch = t0; ch < 5 * 2 * T * CH; ch += nthr) {
    const int r = ch / (2 * T * CH), rem = ch - r * (2 * T * CH), row = rem / CH, k8 = rem - row * CH;
    float v[8];
    if (row < T) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int t = (k8 * 8 + j) - row + pad;
        v[j] = (t >= 0 && t < KL && row < W) ? w2s[r * KL + t] : 0.f;   // output columns beyond the plane: zero
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int t = (k8 * 8 + j) - (row - T) + 2;
        v[j] = (t >= 0 && t < 5 && row - T < W) ? w3s[r * 5 + t] : 0.f;
      }
    }
    *reinterpret_cast<uint4*>(tp + Cfg::kToep1 + r * (2 * T * 128) + row * 128 + ((k8 ^ (row & 7)) << 4)) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

template <int T, int CB, bool TMA>
__global__ void __launch_bounds__(fwd_threads(T, TMA), 1)
lk3_fwd_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap y1map,
                  const __grid_constant__ CUtensorMap y2map, const __grid_constant__ CUtensorMap y3map, FwdParams P) {
  using Cfg = FwdCfg<T>;
  constexpr int PPU = Cfg::PPU, KSTEPS = Cfg::KSTEPS, E = CB / 2, UPS = Cfg::UPS, PLANES = Cfg::PLANES, NT = Cfg::NT;
  constexpr int kNumLoaders = TMA ? 1 : 3;
  constexpr int kThreads = fwd_threads(T, TMA);
  constexpr int kLoaderSplit = fwd_loader_split(T);
  constexpr int kNumTransposerWarps = fwd_transposers(TMA);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int upc = P.units_per_c;
  // work range of this CTA: global item g = c * upc + u
  long long g0, g1;
  int slot0;                                  // stats slot of this CTA inside its first channel
  if (T == 64) {
    const int c = blockIdx.x / P.splits, split = blockIdx.x % P.splits;
    g0 = (long long)c * upc + ((long long)upc * split) / P.splits;
    g1 = (long long)c * upc + ((long long)upc * (split + 1)) / P.splits;
    slot0 = split;
  } else {
    const long long total = (long long)P.C * upc;
    g0 = (long long)blockIdx.x * P.per_cta;
    g1 = g0 + P.per_cta < total ? g0 + P.per_cta : total;
    if (g0 > total) g0 = total;
    slot0 = 0;
  }
  const int n_units = (int)(g1 - g0);
  const int c_first = n_units > 0 ? (int)(g0 / upc) : 0;
  const int c_last = n_units > 0 ? (int)((g1 - 1) / upc) : -1;
  const int KL = P.KL, pad = KL / 2, H = P.H, W = P.W;

  constexpr int B_XN_FULL = 0, B_XN_EMPTY = kStages, B_XT_FULL = 2 * kStages, B_XT_EMPTY = B_XT_FULL + 1,
                B_ACC_FULL = B_XT_EMPTY + 1, B_ACC_EMPTY = B_ACC_FULL + kAccBufs, B_TP_FULL = B_ACC_EMPTY + kAccBufs,
                B_TP_EMPTY = B_TP_FULL + NT;
  const uint32_t bar0 = base + Cfg::kOffBar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + Cfg::kOffBar + 768);
  auto XN_ADDR = [&](int s) { return base + Cfg::kOffXN + s * kXSlot + kPad; };

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(BAR(B_XN_FULL + s), TMA ? 1 : kLoaderSplit);     // TMA expect_tx arrive / lane 0 of each loader warp of the slot
      mbar_init(BAR(B_XN_EMPTY + s), 1 + kNumTransposerWarps);   // MMA commit + transposers done reading
    }
    mbar_init(BAR(B_XT_FULL), kNumTransposerWarps);              // transposers wrote X^T
    mbar_init(BAR(B_XT_EMPTY), 1);                               // MMA commit
    for (int a = 0; a < kAccBufs; ++a) {
      mbar_init(BAR(B_ACC_FULL + a), 1);                         // MMA commit
      mbar_init(BAR(B_ACC_EMPTY + a), 4);                        // one arrival per warp of the group that owns it
    }
    for (int s = 0; s < NT; ++s) {
      mbar_init(BAR(B_TP_FULL + s), 1);                          // builder warp
      mbar_init(BAR(B_TP_EMPTY + s), 1);                         // MMA commit after the channel's last MMA
    }
    mbar_fence_init();
    if (TMA) { tma_prefetch_desc(&xmap); tma_prefetch_desc(&y1map); tma_prefetch_desc(&y2map); tma_prefetch_desc(&y3map); }
  }
  {  // X / X^T slots start as zeros: pads and tile padding are never written afterwards
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < ((kStages + 1) * kXSlot + kPad) / 16; i += kThreads)
      reinterpret_cast<uint4*>(sm + Cfg::kOffXN)[i] = z;
  }
  if (NT == 1 && n_units > 0) {   // single-channel range: every thread helps building the one Toeplitz set
    float* w1s = reinterpret_cast<float*>(sm + Cfg::kOffW);
    float* w2s = w1s + KL * 5;
    float* w3s = w2s + KL * 5;
    for (int i = tid; i < KL * 5; i += kThreads) {
      w1s[i] = P.w1[(size_t)c_first * KL * 5 + i];
      w2s[i] = P.w2[(size_t)c_first * KL * 5 + i];
    }
    if (tid < 25) w3s[tid] = P.w3[(size_t)c_first * 25 + tid];
    __syncthreads();
    build_toeplitz<T>(sm + Cfg::kOffToep, w1s, w2s, w3s, KL, pad, H, W, tid, kThreads);
  }
  fence_proxy_async();
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
#ifdef SLAK_ROLE_PROFILE
  long long prof_wait = 0;
  const long long prof_t0 = clock64();
#endif

  // warps 12-13: two more transposers in the TMA class, the two extra cp.async loaders otherwise
  const bool is_loader = (warp == 0) || (!TMA && (warp == 12 || warp == 13 || warp >= 15));   // w15+: only with a split
  const bool is_transposer = (warp == 2 || warp == 3 || (TMA && (warp == 12 || warp == 13)));
  if (is_loader) {
    if constexpr (TMA) {
      // ================= TMA producer (T = 64: two planes per unit) =================
      if (elect_one()) {
        for (int i = 0; i < n_units; ++i) {
          const long long g = g0 + i;
          const int c = (int)(g / upc), u = (int)(g - (long long)c * upc);
          const int st = i % kStages, ph = (i / kStages) & 1;
          MBW(BAR(B_XN_EMPTY + st), ph ^ 1);
          const int n0 = PLANES * u;
          const uint32_t dst = XN_ADDR(st);
          mbar_expect_tx(BAR(B_XN_FULL + st), kUnitBytes);
#pragma unroll
          for (int pl = 0; pl < PPU; ++pl)
            tma_load_3d(dst + pl * (T * 128), &xmap, BAR(B_XN_FULL + st), 0, 0, min(n0 + pl, P.N - 1) * P.C + c);
        }
      }
    } else {
      // ================= cp.async loaders: loader j owns slot j (one unit = PLANES planes in flight each) ====
      // slot lj = 0, 1, 2 is filled by two warps: (w0, w15), (w12, w16), (w13, w17); each takes half of the planes
      const int lj = (warp == 0) ? 0 : (warp < 15 ? warp - 11 : warp - 15);
      const int hf = warp >= 15 ? 1 : 0;
      constexpr int QN = PLANES / kLoaderSplit;
      const int qlo = hf * QN, qhi = qlo + QN;
      PieceMap<CB> pm;
      pm.init(H, W, lane);
      const size_t plane_bytes = (size_t)H * W * 2;
      for (int i = lj; i < n_units; i += kNumLoaders) {
        const long long g = g0 + i;
        const int c = (int)(g / upc), u = (int)(g - (long long)c * upc);
        const int s = lj, ph = (i / kStages) & 1;             // lj == i % kStages
        MBW(BAR(B_XN_EMPTY + s), ph ^ 1);
        const uint32_t tile = XN_ADDR(s);
        const int n0 = PLANES * u;
        if (CB == 2 && pm.count >= 0 && pm.count <= 2) {
          // batches of 8 planes: loads first, then stores
          for (int q0 = qlo; q0 < qhi; q0 += 8) {
            const uint8_t* srcs[8]; int r0s[8], c0s[8];
            int cnt = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int q = q0 + j;
              srcs[j] = reinterpret_cast<const uint8_t*>(P.x); r0s[j] = 0; c0s[j] = 0;
              if (q < qhi && n0 + q < P.N) {
                srcs[j] = reinterpret_cast<const uint8_t*>(P.x) + ((size_t)(n0 + q) * P.C + c) * plane_bytes;
                r0s[j] = (q % PPU) * T; c0s[j] = (q / PPU) * (T / 8);
                cnt = j + 1;
              }
            }
            if constexpr (CB == 2) load_plane_blocks_cb2<8>(pm, srcs, tile, r0s, c0s, cnt, lane);
          }
        } else {
          for (int q = qlo; q < qhi; ++q)
            if (n0 + q < P.N)
              load_plane_block<CB>(pm, reinterpret_cast<const uint8_t*>(P.x) + ((size_t)(n0 + q) * P.C + c) * plane_bytes,
                                   tile, (q % PPU) * T, (q / PPU) * (T / 8), lane);
        }
        cp_async_commit();
        cp_async_wait_all();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_XN_FULL + s));
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one()) {
      constexpr uint32_t idesc23 = umma_idesc_bf16(128, 2 * T);
      constexpr uint32_t idesc1 = umma_idesc_bf16(128, T);
      int cur_c = -1, k = -1;                         // k = index of the current channel inside this CTA's range
      for (int i = 0; i < n_units; ++i) {
        const int c = (int)((g0 + i) / upc);
        if (c != cur_c) {
          if (k >= 0) umma_commit(BAR(B_TP_EMPTY + (k % NT)));   // previous channel's Toeplitz set is free once its MMAs retire
          cur_c = c; ++k;
          if (NT > 1) MBW(BAR(B_TP_FULL + (k % NT)), (k / NT) & 1);
        }
        const uint32_t toep = base + Cfg::kOffToep + (k % NT) * Cfg::kToepSet;
        const int st = i % kStages, ph = (i / kStages) & 1;
        const int ab = i % kAccBufs, aph = (i / kAccBufs) & 1;
        MBW(BAR(B_ACC_EMPTY + ab), aph ^ 1);   // epilogue drained this accumulator buffer
        MBW(BAR(B_XN_FULL + st), ph);          // X landed
        tc_fence_after();
        const uint32_t xn = XN_ADDR(st);
        const uint32_t xt = base + Cfg::kOffXT + kPad;
        const uint32_t acc = tmem + ab * Cfg::kAccCols;
#pragma unroll
        for (int g = 0; g < UPS; ++g)                // column band g
#pragma unroll
          for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int kk = 0; kk < KSTEPS; ++kk) {
              const uint32_t a = xn + (r - 2) * 128 + g * (T * 2) + kk * 32;
              const uint32_t b = toep + Cfg::kToep1 + r * (2 * T * 128) + kk * 32;
              umma_bf16(acc + g * 3 * T + T, umma_desc_k_sw128(a, 0), umma_desc_k_sw128(b, 0), idesc23, (r | kk) != 0);
            }
        umma_commit(BAR(B_XN_EMPTY + st));           // X slot free (with the transposers' arrivals)
        MBW(BAR(B_XT_FULL), i & 1);            // X^T written
        tc_fence_after();
#pragma unroll
        for (int g = 0; g < UPS; ++g)
#pragma unroll
          for (int s = 0; s < 5; ++s)
#pragma unroll
            for (int kk = 0; kk < KSTEPS; ++kk) {
              const uint32_t a = xt + (s - 2) * 128 + g * (T * 2) + kk * 32;
              const uint32_t b = toep + s * (T * 128) + kk * 32;
              umma_bf16(acc + g * 3 * T, umma_desc_k_sw128(a, 0), umma_desc_k_sw128(b, 0), idesc1, (s | kk) != 0);
            }
        umma_commit(BAR(B_XT_EMPTY));                // X^T slot free
        umma_commit(BAR(B_ACC_FULL + ab));           // accumulators ready
      }
    }
  } else if (is_transposer) {
    // ================= transposers: X (natural) -> X^T, 8x8 blocks =================
    const int tw = warp < 4 ? warp - 2 : warp - 10;        // 0..3
    const int m = lane >> 3, kk = lane & 7;     // matrix id within the x4, row within the 8x8 block
    constexpr int NB = T / 8;                   // blocks per plane edge
    constexpr int ITERS = 32 / kNumTransposerWarps;
    // the block -> (source, destination) offsets are the same for every unit: computed once
    uint32_t soff[ITERS], doff[ITERS];
#pragma unroll
    for (int q = 0; q < ITERS; ++q) {           // 128 8x8 blocks: every T x T block in place
      const int blk = 4 * (tw + q * kNumTransposerWarps) + m;
      const int g = blk / (2 * T), rem0 = blk - g * (2 * T);      // column band
      const int pl = rem0 / (NB * NB), rem = rem0 - pl * (NB * NB);
      const int bi = rem / NB, bj = rem - bi * NB;
      soff[q] = (pl * T + 8 * bi + kk) * 128 + (((g * NB + bj) ^ kk) << 4);
      doff[q] = (pl * T + 8 * bj + kk) * 128 + (((g * NB + bi) ^ kk) << 4);
    }
    const uint32_t xt = base + Cfg::kOffXT + kPad;
    for (int i = 0; i < n_units; ++i) {
      const int st = i % kStages, ph = (i / kStages) & 1;
      MBW(BAR(B_XN_FULL + st), ph);       // X landed
      const uint32_t xn = XN_ADDR(st);
      uint32_t r[ITERS][4];
#pragma unroll
      for (int q = 0; q < ITERS; ++q) ldmatrix_x4_trans(xn + soff[q], r[q][0], r[q][1], r[q][2], r[q][3]);
      MBW(BAR(B_XT_EMPTY), (i & 1) ^ 1);  // previous X^T consumed (the loads above do not depend on it)
#pragma unroll
      for (int q = 0; q < ITERS; ++q) stmatrix_x4(xt + doff[q], r[q][0], r[q][1], r[q][2], r[q][3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(BAR(B_XT_FULL));            // X^T ready
        mbar_arrive(BAR(B_XN_EMPTY + st));      // done reading X
      }
    }
  } else if (warp < 12) {
    // ================= epilogue: group wg drains accumulator buffer wg, i.e. units i = wg (mod 2) =================
    const int wg = (warp - 4) >> 2, e = (warp - 4) & 3;
    const int L = e * 32 + lane;                // TMEM lane = (row group, row)
    const int pl = L / T, row = L % T;
    const size_t plane_elems = (size_t)H * W;
    uint8_t* y1s = sm + Cfg::kOffY1 + wg * kUnitBytes;
    const int nb = 1 + wg;                      // named barrier of this group
    const int PR = W / E;                       // pieces per output row
    float st_s[3] = {0.f, 0.f, 0.f}, st_q[3] = {0.f, 0.f, 0.f};   // BatchNorm statistics of this thread's elements
    const bool want_stats = P.stats != nullptr;
    int cur_c = c_first;
    // write the statistics of channel `ch` gathered by this CTA (lanes -> warp -> the four epilogue warps)
    auto flush_stats = [&](int ch) {
      float* red = reinterpret_cast<float*>(y1s);   // staging is free between units (after the TMA class has drained it)
      if constexpr (TMA) {
        if (e == 0 && lane == 0) bulk_wait_group_read<0>();
        NBS(nb, 128);
      }
#pragma unroll
      for (int k2 = 0; k2 < 3; ++k2) {
        float s = st_s[k2], q = st_q[k2];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
        if (lane == 0) { red[e * 6 + 2 * k2] = s; red[e * 6 + 2 * k2 + 1] = q; }
        st_s[k2] = 0.f; st_q[k2] = 0.f;
      }
      NBS(nb, 128);
      // slot: T=64 -> split; multi-channel -> 0 for channels this CTA starts, 1.. when the channel began in an earlier CTA
      int slot = slot0;
      if (T != 64) {
        const long long cta_of_first = ((long long)ch * upc) / P.per_cta;   // CTA holding the channel's first unit
        slot = (int)(blockIdx.x - cta_of_first);
      }
      if (e == 0 && lane < 6 && slot < P.splits)
        P.stats[(((size_t)ch * P.splits + slot) * kEpiGroups + wg) * 6 + lane] = red[lane] + red[6 + lane] + red[12 + lane] + red[18 + lane];
      NBS(nb, 128);
    };
    for (int i = wg; i < n_units; i += kEpiGroups) {
      const long long gidx = g0 + i;
      const int c = (int)(gidx / upc), u = (int)(gidx - (long long)c * upc);
      if (want_stats && c != cur_c) { flush_stats(cur_c); cur_c = c; }
      const int ab = wg, aph = (i / kAccBufs) & 1;
      MBW(BAR(B_ACC_FULL + ab), aph);
      tc_fence_after();
      uint32_t v[T];
      if constexpr (TMA) {
        // T = 64: every output tile goes registers -> swizzled staging tile -> one TMA tile store per plane (the box
        // is clipped to H x W); row-per-thread global stores would hit 32 different sectors per instruction
        const int n = PLANES * u + pl;
        const bool ok = (n < P.N) && (row < H);
        const uint32_t t0 = tmem + ((uint32_t)(e * 32) << 16) + ab * Cfg::kAccCols;
        const uint32_t stg = base + Cfg::kOffY1 + wg * kUnitBytes;
        // the previous tile store of this group must have finished reading the staging tile before it is rewritten;
        // waiting here (not right after issuing it) lets that read overlap the TMEM load and the statistics
        auto staging_free = [&]() {
          if (e == 0 && lane == 0) bulk_wait_group_read<0>();
          NBS(nb, 128);
        };
        auto store_tile = [&](const CUtensorMap* map) {
          fence_proxy_async();                      // this thread's staging writes -> visible to the TMA engine
          NBS(nb, 128);
          if (e == 0 && lane == 0) {
#pragma unroll
            for (int pq = 0; pq < PPU; ++pq)
              if (PLANES * u + pq < P.N) tma_store_3d(map, stg + pq * (T * 128), 0, 0, (PLANES * u + pq) * P.C + c);
            bulk_commit_group();
          }
        };
        uint32_t* va = v;
        uint32_t* vb = v + 32;
        auto chunk_nat = [&](int br, int h, const uint32_t* w) {      // columns 32h .. 32h+31 of y2 / y3, row L
          if (want_stats && ok) {
            // columns >= W of the accumulator are exact zeros (zero Toeplitz rows): no per-element predicate, and
            // four independent chains instead of one 32-long dependent one
            // packed fp32 (FADD2 / FFMA2): two columns per instruction, four independent chains
            f2 s2[2] = {splat(0.f), splat(0.f)}, q2[2] = {splat(0.f), splat(0.f)};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const f2 pr = mk2u(w[2 * j], w[2 * j + 1]);
              s2[j & 1] = add2(s2[j & 1], pr);
              q2[j & 1] = fma2(pr, pr, q2[j & 1]);
            }
            float sa, sb, qa, qb;
            un2(add2(s2[0], s2[1]), sa, sb);
            un2(add2(q2[0], q2[1]), qa, qb);
            st_s[1 + br] += sa + sb; st_q[1 + br] += qa + qb;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(y1s + L * 128 + (((4 * h + j) ^ (L & 7)) << 4)) =
                make_uint4(pack_bf16(__uint_as_float(w[8 * j]), __uint_as_float(w[8 * j + 1])),
                           pack_bf16(__uint_as_float(w[8 * j + 2]), __uint_as_float(w[8 * j + 3])),
                           pack_bf16(__uint_as_float(w[8 * j + 4]), __uint_as_float(w[8 * j + 5])),
                           pack_bf16(__uint_as_float(w[8 * j + 6]), __uint_as_float(w[8 * j + 7])));
        };
        auto chunk_t = [&](int h, const uint32_t* w) {                // y1^T: column q = row, p = 32h .. 32h+31
          if (want_stats && n < P.N && row < W) {       // rows p >= H of y1 are exact zeros as well
            f2 s2[2] = {splat(0.f), splat(0.f)}, q2[2] = {splat(0.f), splat(0.f)};
#pragma unroll
            for (int p = 0; p < 16; ++p) {
              const f2 pr = mk2u(w[2 * p], w[2 * p + 1]);
              s2[p & 1] = add2(s2[p & 1], pr);
              q2[p & 1] = fma2(pr, pr, q2[p & 1]);
            }
            float sa, sb, qa, qb;
            un2(add2(s2[0], s2[1]), sa, sb);
            un2(add2(q2[0], q2[1]), qa, qb);
            st_s[0] += sa + sb; st_q[0] += qa + qb;
          }
#pragma unroll
          for (int p = 0; p < 32; ++p) {
            const uint32_t r = (uint32_t)(pl * T + 32 * h + p);
            const uint32_t off = r * 128 + ((((uint32_t)row >> 3)) ^ (r & 7)) * 16 + (row & 7) * 2;
            *reinterpret_cast<__nv_bfloat16*>(y1s + off) = __float2bfloat16_rn(__uint_as_float(w[p]));
          }
        };
        tmem_ld32(t0 + T, va); tmem_ld32(t0 + T + 32, vb);
        tmem_ld_wait();
        staging_free();
        chunk_nat(0, 0, va); chunk_nat(0, 1, vb);
        store_tile(&y2map);
        tmem_ld32(t0 + 2 * T, va); tmem_ld32(t0 + 2 * T + 32, vb);
        tmem_ld_wait();
        staging_free();
        chunk_nat(1, 0, va); chunk_nat(1, 1, vb);
        store_tile(&y3map);
        tmem_ld32(t0, va); tmem_ld32(t0 + 32, vb);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + ab));   // accumulators drained
        staging_free();
        chunk_t(0, va); chunk_t(1, vb);
        store_tile(&y1map);
        continue;
      }
#pragma unroll
      for (int g = 0; g < UPS; ++g) {               // column band g: plane g*PPU + pl
        const int n = PLANES * u + g * PPU + pl;
        const bool ok = (n < P.N) && (row < H);
        const size_t rbase = ((size_t)(n < P.N ? n : 0) * P.C + c) * plane_elems + (size_t)(row < H ? row : 0) * W;
        const uint32_t t0 = tmem + ((uint32_t)(e * 32) << 16) + ab * Cfg::kAccCols + g * 3 * T;
        // ---- y2 (cols T..2T-1 of the group) and y3 (cols 2T..3T-1): natural orientation ----
#pragma unroll
        for (int br = 0; br < 2; ++br) {
          tmem_ld_cols<T>(t0 + T + br * T, v);
          tmem_ld_wait();
          if (ok) {
            __nv_bfloat16* yo = (br == 0 ? P.y2 : P.y3) + rbase;
#pragma unroll
            for (int j = 0; j < T / E; ++j)          // static register indices
              if (j < PR) store_bf16_piece<E>(yo + j * E, v + j * E);
            if (want_stats) {
              float s = 0.f, q = 0.f;
#pragma unroll
              for (int j = 0; j < T; ++j)
                if (j < W) { const float f = __uint_as_float(v[j]); s += f; q = fmaf(f, f, q); }
              st_s[1 + br] += s; st_q[1 + br] += q;
            }
          }
        }
        // ---- y1^T (cols 0..T-1): this thread holds column `row`(=q) for p = 0..T-1 -> staging[(pl,p)][band g, q] ----
        tmem_ld_cols<T>(t0, v);
        tmem_ld_wait();
#pragma unroll
        for (int p = 0; p < T; ++p) {
          const uint32_t r = (uint32_t)(pl * T + p);
          const uint32_t off = r * 128 + (((uint32_t)(g * (T / 8)) + ((uint32_t)row >> 3)) ^ (r & 7)) * 16 + (row & 7) * 2;
          *reinterpret_cast<__nv_bfloat16*>(y1s + off) = __float2bfloat16_rn(__uint_as_float(v[p]));
        }
        if (want_stats && n < P.N && row < W) {       // this thread holds column q = row of y1 for p < H
          float s = 0.f, q = 0.f;
#pragma unroll
          for (int p = 0; p < T; ++p)
            if (p < H) { const float f = __uint_as_float(v[p]); s += f; q = fmaf(f, f, q); }
          st_s[0] += s; st_q[0] += q;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + ab));   // accumulators drained
      NBS(nb, 128);
#pragma unroll
      for (int g = 0; g < UPS; ++g) {
        const int n = PLANES * u + g * PPU + pl;
        if ((n < P.N) && (row < H)) {
          __nv_bfloat16* yo = P.y1 + ((size_t)n * P.C + c) * plane_elems + (size_t)row * W;
          const uint32_t r = (uint32_t)(pl * T + row);
#pragma unroll
          for (int j = 0; j < T / E; ++j)
            if (j < PR) {
              const uint32_t b = j * CB;
              copy_piece<E>(yo + j * E, y1s + r * 128 + ((((uint32_t)(g * (T / 8)) + (b >> 4)) ^ (r & 7)) << 4) + (b & 15));
            }
        }
      }
      NBS(nb, 128);                    // staging free for the next unit
    }
    if (want_stats && n_units > 0) flush_stats(cur_c);
    if constexpr (TMA) {
      if (e == 0 && lane == 0) bulk_wait_group_read<0>();   // shared memory must outlive the last tile store
    }
  } else if (warp == 14) {
    // ================= Toeplitz builder: one set per channel of the range, NT sets in flight =================
    float* w1s = reinterpret_cast<float*>(sm + Cfg::kOffW);   // [KL][5]
    float* w2s = w1s + KL * 5;                                // [5][KL]
    float* w3s = w2s + KL * 5;                                // [5][5]
    for (int c = (NT == 1 ? c_last + 1 : c_first), k = 0; c <= c_last; ++c, ++k) {   // NT == 1: built by all threads below
      const int set = k % NT;
      MBW(BAR(B_TP_EMPTY + set), ((k / NT) & 1) ^ 1);
      uint8_t* tp = sm + Cfg::kOffToep + set * Cfg::kToepSet;
      for (int i = lane; i < KL * 5; i += 32) {
        w1s[i] = P.w1[(size_t)c * KL * 5 + i];
        w2s[i] = P.w2[(size_t)c * KL * 5 + i];
      }
      if (lane < 25) w3s[lane] = P.w3[(size_t)c * 25 + lane];
      __syncwarp();
      build_toeplitz<T>(tp, w1s, w2s, w3s, KL, pad, H, W, lane, 32);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_TP_FULL + set));
    }
  }

#ifdef SLAK_ROLE_PROFILE
  if (blockIdx.x == 0 && lane == 0)
    printf("T=%d warp %2d: loop %lld cycles, waiting %lld, units %d\n", T, warp, clock64() - prof_t0, prof_wait, n_units);
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem);
}

// ---- host side --------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// x viewed as (W, H, N*C) bf16; box 64 x 64 x 1, SWIZZLE_128B, out-of-bounds -> zeros
int make_plane_map(CUtensorMap* map, const void* x, int N, int C, int H, int W) {
  EncodeTiledFn enc = get_encode();
  SLAK_REQUIRE(enc != nullptr, SLAK_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N * C};
  cuuint64_t strides[2] = {(cuuint64_t)W * 2, (cuuint64_t)H * W * 2};
  cuuint32_t box[3] = {64, 64, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SLAK_REQUIRE(r == CUDA_SUCCESS, SLAK_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return SLAK_OK;
}

// Tile class and piece size for a plane shape; tile 0 = not covered by the tensor-core kernels.
TcShape tc_shape(int H, int W) {
  TcShape s{0, 0, false};
  const int m = H > W ? H : W;
  if (H < 1 || W < 1 || m > 62) return s;
  s.tile = m <= 14 ? 16 : (m <= 30 ? 32 : 64);
  const int rb = W * 2;
  s.cb = (rb % 16 == 0) ? 16 : (rb % 8 == 0) ? 8 : (rb % 4 == 0) ? 4 : 2;
  if (s.tile == 64) {
    s.tma = (s.cb == 16);
    if (!s.tma) s.tile = 0;          // 64-class planes go through TMA only
  } else {
    if (s.cb == 16) s.cb = 8;        // instantiated piece sizes for the small classes: 8 / 4 / 2 bytes
    if (s.tile == 16 && s.cb == 8) s.cb = 4;
  }
  return s;
}

bool lk3_tc_supported(int N, int C, int H, int W, int KL) {
  (void)N; (void)C;
  return tc_shape(H, W).tile != 0 && (KL & 1) && KL >= 5 && KL <= 99;
}

int tc_pick_splits(int C, int units) {
  const int sms = sm_count();
  int best = 1; double best_eff = 0.0;
  const int max_s = units >= 4 ? units / 2 : 1;
  for (int s = 1; s <= max_s && s <= 64; ++s) {
    const long long ctas = (long long)C * s;
    const long long waves = (ctas + sms - 1) / sms;
    const int per = (units + s - 1) / s;
    const double eff = (double)C * units / ((double)waves * sms * per) * (per / (per + 1.5));
    if (eff > best_eff) { best_eff = eff; best = s; }
  }
  return best;
}

// Work partition of the channel-walking kernels: grid size, items per CTA, statistics slots per channel.
TcPlan tc_plan(int N, int C, int tile, int planes_per_unit) {
  TcPlan p{};
  p.units_per_c = (N + planes_per_unit - 1) / planes_per_unit;
  if (tile == 64) {
    p.splits = tc_pick_splits(C, p.units_per_c);
    p.grid = C * p.splits;
    p.per_cta = (p.units_per_c + p.splits - 1) / p.splits;
  } else {
    const long long total = (long long)C * p.units_per_c;
    long long grid = sm_count();
    if (grid > total) grid = total;
    p.per_cta = (int)((total + grid - 1) / grid);
    p.grid = (int)((total + p.per_cta - 1) / p.per_cta);
    p.splits = (p.units_per_c + p.per_cta - 1) / p.per_cta + 1;   // CTAs that can touch one channel
  }
  return p;
}

template <int T, int CB, bool TMA>
static int launch_fwd(const CUtensorMap& map, const CUtensorMap* ymaps, FwdParams& P, cudaStream_t st) {
  using Cfg = FwdCfg<T>;
  const TcPlan plan = tc_plan(P.N, P.C, T, Cfg::PLANES);
  P.units_per_c = plan.units_per_c;
  P.per_cta = plan.per_cta;
  P.splits = plan.splits;
  if (P.stats)   // not every slot of a channel is written in the multi-channel partition
    SLAK_CUDA_TRY(cudaMemsetAsync(P.stats, 0, (size_t)P.C * plan.splits * kEpiGroups * 6 * sizeof(float), st));
  auto kern = lk3_fwd_tc_kernel<T, CB, TMA>;
  SLAK_SET_MAX_SMEM(kern, Cfg::kSmem);
  kern<<<plan.grid, fwd_threads(T, TMA), Cfg::kSmem, st>>>(map, ymaps[0], ymaps[1], ymaps[2], P);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

// dwconv_tc_dense.cu: planes of up to 208 pixels as dense per-channel GEMMs
namespace dense {
bool supported(int N, int C, int H, int W, int KL);
int stats_slots(int N);
int fwd(const void* x, const float* w1, const float* w2, const float* w3, void* y1, void* y2, void* y3, int N, int C, int H, int W,
        int KL, float* stats, cudaStream_t st);
}

int lk3_fwd_tc_splits(int N, int C, int H, int W) {
  if (dense::supported(N, C, H, W, 5)) return dense::stats_slots(N);
  const TcShape s = tc_shape(H, W);
  if (s.tile == 0) return 0;
  return tc_plan(N, C, s.tile, (128 / s.tile) * (64 / s.tile)).splits * kEpiGroups;   // statistics slots per channel
}

int lk3_fwd_tc(const void* x, const float* w1, const float* w2, const float* w3, void* y1, void* y2, void* y3,
               int N, int C, int H, int W, int KL, float* stats, cudaStream_t st) {
  SLAK_REQUIRE(lk3_tc_supported(N, C, H, W, KL), SLAK_ERR_UNSUPPORTED, "shape %dx%d not covered by the tensor-core path", H, W);
  SLAK_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, SLAK_ERR_BAD_ARG, "x must be 16-byte aligned");
  SLAK_REQUIRE((2 * KL * 5 + 25) * 4 <= 4096, SLAK_ERR_UNSUPPORTED, "kernel side %d too large", KL);
  if (dense::supported(N, C, H, W, KL)) return dense::fwd(x, w1, w2, w3, y1, y2, y3, N, C, H, W, KL, stats, st);
  const TcShape s = tc_shape(H, W);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  CUtensorMap ymaps[3];
  memset(ymaps, 0, sizeof(ymaps));
  if (s.tma) {
    int rc = make_plane_map(&map, x, N, C, H, W);
    if (rc) return rc;
    void* ys[3] = {y1, y2, y3};
    for (int k = 0; k < 3; ++k) {
      SLAK_REQUIRE((reinterpret_cast<uintptr_t>(ys[k]) & 15) == 0, SLAK_ERR_BAD_ARG, "outputs must be 16-byte aligned");
      rc = make_plane_map(&ymaps[k], ys[k], N, C, H, W);
      if (rc) return rc;
    }
  }
  FwdParams P;
  P.x = (const __nv_bfloat16*)x;
  P.w1 = w1; P.w2 = w2; P.w3 = w3;
  P.y1 = (__nv_bfloat16*)y1; P.y2 = (__nv_bfloat16*)y2; P.y3 = (__nv_bfloat16*)y3;
  P.N = N; P.C = C; P.H = H; P.W = W; P.KL = KL;
  P.stats = stats;
  if (s.tile == 64) return launch_fwd<64, 16, true>(map, ymaps, P, st);
  if (s.tile == 32) {
    if (s.cb == 8) return launch_fwd<32, 8, false>(map, ymaps, P, st);
    if (s.cb == 4) return launch_fwd<32, 4, false>(map, ymaps, P, st);
    return launch_fwd<32, 2, false>(map, ymaps, P, st);
  }
  if (s.cb == 4) return launch_fwd<16, 4, false>(map, ymaps, P, st);
  return launch_fwd<16, 2, false>(map, ymaps, P, st);
}

}  // namespace tc
}  // namespace slak
