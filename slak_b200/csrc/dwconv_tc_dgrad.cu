// Depthwise data-gradient on the tensor cores (bf16 in, fp32 accumulate in TMEM, bf16 or fp32 out):
//     out = conv(in_t, W_t [KL x 5])  +  conv(in_n, W_n [5 x KN])  +  addend
// with the taps optionally flipped (flip=1: backward_data of the forward convs,
// backward_data_fp32.cu:199-263; dx = sum over the three branches of models/SLaK.py:89-100).
// The same banded-Toeplitz formulation, tile classes (T = 64/32/16), units (one 128-row x 64-column
// tile = 2/8/32 planes), persistent channel walk and Toeplitz builder warp as dwconv_tc_fwd.cu:
//   natural path     D_n  [(plane,p), q] += IN_n [(plane,p+r-2), w] * Tn_r[q, w]
//   transposed path  D_t^T[(plane,q), p] += IN_t^T[(plane,q+s-2), h] * Tt_s[p, h]
// and the epilogue adds D_n + transpose(D_t^T) (+ addend rows read from global) before rounding.
// dx of a Block needs two launches: (in_n = dy3, W_n = 5x5) -> tmp, then
// (in_t = dy1, in_n = dy2, addend = tmp) -> dx; all three Toeplitz sets do not fit in shared memory
// next to a multi-stage input pipeline at T = 64.
//
// Warp roles (352 threads): w0 loader | w1 MMA | w2-3 transposers (w2 owns TMEM) | w4-7 epilogue |
// w8-9 extra loaders (cp.async path) | w10 Toeplitz builder.
#include "common.cuh"
#include "tc_common.cuh"
#include <string.h>

namespace slak {
namespace tc {

int make_plane_map(CUtensorMap* map, const void* x, int N, int C, int H, int W);  // dwconv_tc_fwd.cu

namespace dg {
constexpr int kNStages = 3;                    // natural-path input slots
constexpr int kSStages = 2;                    // transposed-path source slots (natural layout)
constexpr int kAccBufs = 2;
constexpr int kUnit = 128 * 128;
constexpr int kPad = 1024;
constexpr int kSlot = kPad + kUnit + kPad;     // 18 KB
constexpr int kNumTransposerWarps = 2;
// cp.async classes: loader warps per natural-path slot.  A second warp per slot (w11-13 next to w0, w8-9, as the
// forward and weight-gradient kernels do for the 16-class) is wired up but not enabled: the 448-thread CTA caps the
// epilogue at 128 registers and its prefetched addend rows spill (measured 165 -> 222 us at 14x14)
__host__ __device__ constexpr int loader_split(int T) { (void)T; return 1; }
__host__ __device__ constexpr int threads(int T) { return loader_split(T) == 2 ? 448 : 352; }
template <int T> struct Cfg {
  static constexpr int PPU = 128 / T;
  static constexpr int UPS = 64 / T;
  static constexpr int PLANES = PPU * UPS;
  static constexpr int KSTEPS = T / 16;
  static constexpr int NT = (T == 64) ? 1 : 2;
  static constexpr int kToep = 5 * T * 128;                 // one path
  static constexpr int kToepSet = 2 * kToep;                // Tn then Tt
  static constexpr int kOffToep = 0;
  static constexpr int kOffXN = NT * kToepSet;
  static constexpr int kOffXS = kOffXN + kNStages * kSlot;
  static constexpr int kOffXT = kOffXS + kSStages * kUnit;
  static constexpr int kOffEx = kOffXT + kSlot;             // fp32 exchange [128][64]
  static constexpr int kOffW = kOffEx + 128 * 64 * 4;
  static constexpr int kOffBar = kOffW + 4096;
  static constexpr int kSmem = kOffBar + 1024 + 1024;
  static constexpr int kAccCols = 2 * T * UPS;              // 128: per band D_n (T cols) then D_t^T (T cols)
  static constexpr int kTmemCols = 256;
};
}  // namespace dg

struct DgradParams {
  const __nv_bfloat16* in_t; const __nv_bfloat16* in_n;
  const float* wt;                 // [C][KL][5] taps of the transposed (vertical-long) path, or nullptr
  const float* wn;                 // [C][5][KN] taps of the natural path
  const __nv_bfloat16* addend;     // [N,C,H,W] or nullptr
  const float* addend_f32;         // [N,C,H,W] fp32 or nullptr (e.g. the shortcut gradient of a Block)
  const float* bias;               // [C] fp32 or nullptr: per-channel constant added before rounding (merged BN shifts)
  __nv_bfloat16* out;              // bf16 result, or nullptr when out_f32 is given
  float* out_f32;
  int N, C, H, W, KL, KN, flip, has_t, has_n, splits, units_per_c, per_cta;
};

template <int T>
__device__ __forceinline__ void build_toeplitz_pair(uint8_t* tp, const float* wts, const float* wns, int KL, int KN,
                                                    bool has_t, bool has_n, int t0, int nthr) {
  constexpr int CH = T / 8;
  const int padn = KN / 2, padt = KL / 2;
  for (int ch = t0; ch < 5 * T * CH; ch += nthr) {
    const int s = ch / (T * CH), rem = ch - s * (T * CH), row = rem / CH, k8 = rem - row * CH;
    float vn[8], vt[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int tn = (k8 * 8 + j) - row + padn;
      vn[j] = (has_n && tn >= 0 && tn < KN) ? wns[s * KN + tn] : 0.f;
      const int tt = (k8 * 8 + j) - row + padt;
      vt[j] = (has_t && tt >= 0 && tt < KL) ? wts[tt * 5 + s] : 0.f;
    }
    const uint32_t off = s * (T * 128) + row * 128 + ((k8 ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(tp + off) =
        make_uint4(pack_bf16(vn[0], vn[1]), pack_bf16(vn[2], vn[3]), pack_bf16(vn[4], vn[5]), pack_bf16(vn[6], vn[7]));
    *reinterpret_cast<uint4*>(tp + 5 * T * 128 + off) =
        make_uint4(pack_bf16(vt[0], vt[1]), pack_bf16(vt[2], vt[3]), pack_bf16(vt[4], vt[5]), pack_bf16(vt[6], vt[7]));
  }
}
// taps of channel c -> fp32 staging, flipped when asked
__device__ __forceinline__ void stage_taps(const DgradParams& P, int c, float* wts, float* wns, int t0, int nthr) {
  const int KL = P.KL, KN = P.KN;
  if (P.has_t)
    for (int i = t0; i < KL * 5; i += nthr) {
      const int t = i / 5, s = i - t * 5;
      const int src = P.flip ? ((KL - 1 - t) * 5 + (4 - s)) : i;
      wts[i] = P.wt[(size_t)c * KL * 5 + src];
    }
  if (P.has_n)
    for (int i = t0; i < 5 * KN; i += nthr) {
      const int r = i / KN, t = i - r * KN;
      const int src = P.flip ? ((4 - r) * KN + (KN - 1 - t)) : i;
      wns[i] = P.wn[(size_t)c * 5 * KN + src];
    }
}

// GEN = false: the Block data gradient (natural family always present, no bias) -- the instantiation the training step
// runs, kept free of the general modes' branches (they cost registers: the T = 32 epilogue spilled with them);
// GEN = true: single-family modes (K x 5 alone) and the per-channel bias of the merged inference layer.
template <int T, int CB, bool TMA, bool GEN>
__global__ void __launch_bounds__(dg::threads(T), 1)
lk_dgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap nmap, DgradParams P) {
  using namespace dg;
  using Cf = Cfg<T>;
  constexpr int PPU = Cf::PPU, UPS = Cf::UPS, PLANES = Cf::PLANES, KSTEPS = Cf::KSTEPS, NT = Cf::NT, E = CB / 2;
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
  const int c_first = n_units > 0 ? (int)(g0 / upc) : 0;
  const int c_last = n_units > 0 ? (int)((g1 - 1) / upc) : -1;
  const int KL = P.KL, KN = P.KN, H = P.H, W = P.W;
  const bool has_t = P.has_t != 0, has_n = GEN ? (P.has_n != 0) : true;

  constexpr int B_N_FULL = 0, B_N_EMPTY = kNStages, B_S_FULL = 2 * kNStages, B_S_EMPTY = B_S_FULL + kSStages,
                B_T_FULL = B_S_EMPTY + kSStages, B_T_EMPTY = B_T_FULL + 1, B_ACC_FULL = B_T_EMPTY + 1,
                B_ACC_EMPTY = B_ACC_FULL + kAccBufs, B_TP_FULL = B_ACC_EMPTY + kAccBufs, B_TP_EMPTY = B_TP_FULL + NT;
  const uint32_t bar0 = base + Cf::kOffBar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + Cf::kOffBar + 768);

  if (tid == 0) {
    for (int s = 0; s < kNStages; ++s) { mbar_init(BAR(B_N_FULL + s), TMA ? 1 : kSplit); mbar_init(BAR(B_N_EMPTY + s), 1); }
    for (int s = 0; s < kSStages; ++s) { mbar_init(BAR(B_S_FULL + s), TMA ? 1 : kSplit); mbar_init(BAR(B_S_EMPTY + s), kNumTransposerWarps); }
    mbar_init(BAR(B_T_FULL), kNumTransposerWarps);
    mbar_init(BAR(B_T_EMPTY), 1);
    for (int a = 0; a < kAccBufs; ++a) { mbar_init(BAR(B_ACC_FULL + a), 1); mbar_init(BAR(B_ACC_EMPTY + a), 4); }
    for (int s = 0; s < NT; ++s) { mbar_init(BAR(B_TP_FULL + s), 1); mbar_init(BAR(B_TP_EMPTY + s), 1); }
    mbar_fence_init();
    if (TMA) { tma_prefetch_desc(&nmap); if (has_t) tma_prefetch_desc(&tmap); }
  }
  {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < (Cf::kOffEx - Cf::kOffXN) / 16; i += kThreads) reinterpret_cast<uint4*>(sm + Cf::kOffXN)[i] = z;
  }
  if (NT == 1 && n_units > 0) {   // single-channel range: every thread helps building the one Toeplitz set
    float* wts = reinterpret_cast<float*>(sm + Cf::kOffW);
    float* wns = wts + KL * 5;
    stage_taps(P, c_first, wts, wns, tid, kThreads);
    __syncthreads();
    build_toeplitz_pair<T>(sm + Cf::kOffToep, wts, wns, KL, KN, has_t, has_n, tid, kThreads);
  }
  fence_proxy_async();
  if (warp == 2) tmem_alloc<Cf::kTmemCols>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const bool is_loader = (warp == 0) || (!TMA && (warp == 8 || warp == 9 || warp >= 11));   // w11+: only with a split
  if (is_loader) {
    if constexpr (TMA) {
      if (elect_one()) {
        for (int i = 0; i < n_units; ++i) {
          const long long g = g0 + i;
          const int c = (int)(g / upc), u = (int)(g - (long long)c * upc);
          const int n0 = PLANES * u;
          if (has_n) {
            const int st = i % kNStages, ph = (i / kNStages) & 1;
            mbar_wait(BAR(B_N_EMPTY + st), ph ^ 1);
            const uint32_t dst = base + Cf::kOffXN + st * kSlot + kPad;
            mbar_expect_tx(BAR(B_N_FULL + st), kUnit);
#pragma unroll
            for (int pl = 0; pl < PPU; ++pl)
              tma_load_3d(dst + pl * (T * 128), &nmap, BAR(B_N_FULL + st), 0, 0, min(n0 + pl, P.N - 1) * P.C + c);
          }
          if (has_t) {
            const int st = i % kSStages, ph = (i / kSStages) & 1;
            mbar_wait(BAR(B_S_EMPTY + st), ph ^ 1);
            const uint32_t dst = base + Cf::kOffXS + st * kUnit;
            mbar_expect_tx(BAR(B_S_FULL + st), kUnit);
#pragma unroll
            for (int pl = 0; pl < PPU; ++pl)
              tma_load_3d(dst + pl * (T * 128), &tmap, BAR(B_S_FULL + st), 0, 0, min(n0 + pl, P.N - 1) * P.C + c);
          }
        }
      }
    } else {
      // three cp.async loader warps; loader j owns natural slot j; the source slots alternate per unit
      const int lj = (warp == 0) ? 0 : (warp < 11 ? warp - 7 : warp - 11);   // natural slot 0..2
      const int hf = warp >= 11 ? 1 : 0;
      constexpr int QN = PLANES / kSplit;
      const int qlo = hf * QN, qhi = qlo + QN;
      PieceMap<CB> pm;
      pm.init(H, W, lane);
      const size_t plane_bytes = (size_t)H * W * 2;
      // Without the natural path (K x 5 alone) nothing but the two source slots paces the loaders, and a loader that
      // shares a slot with another one could run two mbarrier phases ahead (a parity wait cannot tell phase k from
      // k + 2): then only kSStages loaders work, each owning one source slot.
      const int stride = has_n ? kNStages : kSStages;
      for (int i = (has_n || lj < kSStages) ? lj : n_units; i < n_units; i += stride) {
        const long long g = g0 + i;
        const int c = (int)(g / upc), u = (int)(g - (long long)c * upc);
        const int n0 = PLANES * u;
        const int st = lj, ph = (i / kNStages) & 1;
        if (has_n) mbar_wait(BAR(B_N_EMPTY + st), ph ^ 1);
        const int ss = i % kSStages, sph = (i / kSStages) & 1;
        if (has_t) mbar_wait(BAR(B_S_EMPTY + ss), sph ^ 1);
        const uint32_t tn = base + Cf::kOffXN + st * kSlot + kPad;
        const uint32_t ts = base + Cf::kOffXS + ss * kUnit;
        if (CB == 2 && pm.count >= 0 && pm.count <= 2) {
          for (int q0 = qlo; q0 < qhi; q0 += 8) {
            const uint8_t* sn[8]; const uint8_t* stt[8]; int r0s[8], c0s[8];
            int cnt = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int q = q0 + j;
              sn[j] = reinterpret_cast<const uint8_t*>(P.in_n); stt[j] = sn[j]; r0s[j] = 0; c0s[j] = 0;
              if (q < qhi && n0 + q < P.N) {
                const size_t off = ((size_t)(n0 + q) * P.C + c) * plane_bytes;
                sn[j] = reinterpret_cast<const uint8_t*>(P.in_n) + off;
                stt[j] = has_t ? reinterpret_cast<const uint8_t*>(P.in_t) + off : sn[j];
                r0s[j] = (q % PPU) * T; c0s[j] = (q / PPU) * (T / 8);
                cnt = j + 1;
              }
            }
            if constexpr (CB == 2) {
              if (has_n) load_plane_blocks_cb2<8>(pm, sn, tn, r0s, c0s, cnt, lane);
              if (has_t) load_plane_blocks_cb2<8>(pm, stt, ts, r0s, c0s, cnt, lane);
            }
          }
        } else {
          for (int q = qlo; q < qhi; ++q)
            if (n0 + q < P.N) {
              const size_t off = ((size_t)(n0 + q) * P.C + c) * plane_bytes;
              if (has_n)
                load_plane_block<CB>(pm, reinterpret_cast<const uint8_t*>(P.in_n) + off, tn, (q % PPU) * T, (q / PPU) * (T / 8), lane);
              if (has_t)
                load_plane_block<CB>(pm, reinterpret_cast<const uint8_t*>(P.in_t) + off, ts, (q % PPU) * T, (q / PPU) * (T / 8), lane);
            }
        }
        cp_async_commit();
        cp_async_wait_all();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (has_n) mbar_arrive(BAR(B_N_FULL + st));
          if (has_t) mbar_arrive(BAR(B_S_FULL + ss));
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, T);
      int cur_c = -1, k = -1;
      for (int i = 0; i < n_units; ++i) {
        const int c = (int)((g0 + i) / upc);
        if (c != cur_c) {
          if (k >= 0) umma_commit(BAR(B_TP_EMPTY + (k % NT)));
          cur_c = c; ++k;
          if (NT > 1) mbar_wait(BAR(B_TP_FULL + (k % NT)), (k / NT) & 1);
        }
        const uint32_t toep = base + Cf::kOffToep + (k % NT) * Cf::kToepSet;
        const int st = i % kNStages, ph = (i / kNStages) & 1;
        const int ab = i % kAccBufs, aph = (i / kAccBufs) & 1;
        mbar_wait(BAR(B_ACC_EMPTY + ab), aph ^ 1);
        const uint32_t acc = tmem + ab * Cf::kAccCols;
        if (has_n) {
          mbar_wait(BAR(B_N_FULL + st), ph);
          tc_fence_after();
          const uint32_t xn = base + Cf::kOffXN + st * kSlot + kPad;
#pragma unroll
          for (int g = 0; g < UPS; ++g)
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
              for (int kk = 0; kk < KSTEPS; ++kk)
                umma_bf16(acc + g * 2 * T, umma_desc_k_sw128(xn + (r - 2) * 128 + g * (T * 2) + kk * 32, 0),
                          umma_desc_k_sw128(toep + r * (T * 128) + kk * 32, 0), idesc, (r | kk) != 0);
          umma_commit(BAR(B_N_EMPTY + st));
        }
        if (has_t) {
          mbar_wait(BAR(B_T_FULL), i & 1);
          tc_fence_after();
          const uint32_t xt = base + Cf::kOffXT + kPad;
#pragma unroll
          for (int g = 0; g < UPS; ++g)
#pragma unroll
            for (int s = 0; s < 5; ++s)
#pragma unroll
              for (int kk = 0; kk < KSTEPS; ++kk)
                umma_bf16(acc + g * 2 * T + T, umma_desc_k_sw128(xt + (s - 2) * 128 + g * (T * 2) + kk * 32, 0),
                          umma_desc_k_sw128(toep + Cf::kToep + s * (T * 128) + kk * 32, 0), idesc, (s | kk) != 0);
          umma_commit(BAR(B_T_EMPTY));
        }
        umma_commit(BAR(B_ACC_FULL + ab));
      }
    }
  } else if (warp < 4) {
    // ================= transposers =================
    if (has_t) {
      const int tw = warp - 2;
      const int m = lane >> 3, kk = lane & 7;
      constexpr int NB = T / 8;
      constexpr int ITERS = 32 / kNumTransposerWarps;
      // block -> (source, destination) offsets are the same for every unit: computed once
      uint32_t soff[ITERS], doff[ITERS];
#pragma unroll
      for (int q = 0; q < ITERS; ++q) {
        const int blk = 4 * (tw + q * kNumTransposerWarps) + m;
        const int g = blk / (2 * T), rem0 = blk - g * (2 * T);
        const int pl = rem0 / (NB * NB), rem = rem0 - pl * (NB * NB);
        const int bi = rem / NB, bj = rem - bi * NB;
        soff[q] = (pl * T + 8 * bi + kk) * 128 + (((g * NB + bj) ^ kk) << 4);
        doff[q] = (pl * T + 8 * bj + kk) * 128 + (((g * NB + bi) ^ kk) << 4);
      }
      const uint32_t xt = base + Cf::kOffXT + kPad;
      for (int i = 0; i < n_units; ++i) {
        const int st = i % kSStages, ph = (i / kSStages) & 1;
        mbar_wait(BAR(B_S_FULL + st), ph);
        const uint32_t xs = base + Cf::kOffXS + st * kUnit;
#pragma unroll
        for (int h = 0; h < 2; ++h) {            // two halves: 8 x4 matrices in registers at a time
          uint32_t r[ITERS / 2][4];
#pragma unroll
          for (int q = 0; q < ITERS / 2; ++q)
            ldmatrix_x4_trans(xs + soff[h * (ITERS / 2) + q], r[q][0], r[q][1], r[q][2], r[q][3]);
          if (h == 0) mbar_wait(BAR(B_T_EMPTY), (i & 1) ^ 1);   // the first loads do not depend on the target slot
#pragma unroll
          for (int q = 0; q < ITERS / 2; ++q)
            stmatrix_x4(xt + doff[h * (ITERS / 2) + q], r[q][0], r[q][1], r[q][2], r[q][3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(BAR(B_T_FULL));
          mbar_arrive(BAR(B_S_EMPTY + st));
        }
      }
    }
  } else if (warp < 8) {
    // ================= epilogue =================
    const int e = warp - 4;
    const int L = e * 32 + lane;
    const int pl = L / T, row = L % T;
    const size_t plane_elems = (size_t)H * W;
    float* ex = reinterpret_cast<float*>(sm + Cf::kOffEx);   // [128 rows][64 floats]: band g uses floats [g*T, g*T+T)
    const int PR = W / E;
    constexpr int XM = T / 4 - 1;               // float4-chunk XOR mask inside a band
    constexpr bool PFB = (E >= 2);              // bf16 addend prefetched into registers (32 per thread)
    constexpr bool PFF = PFB && (T < 64);       // fp32 addend too (64 more) where the accumulator row is short
    for (int i = 0; i < n_units; ++i) {
      const long long gidx = g0 + i;
      const int c = (int)(gidx / upc), u = (int)(gidx - (long long)c * upc);
      const int ab = i % kAccBufs, aph = (i / kAccBufs) & 1;
      // addend rows of this unit are requested BEFORE the accumulators are waited for: one global round trip per
      // unit, hidden behind the MMAs, instead of one per band and addend after them
      uint32_t pfb[PFB ? UPS * (T / 2) : 1];
      uint32_t pff[PFF ? UPS * T : 1];
      if constexpr (PFB) {
#pragma unroll
        for (int g = 0; g < UPS; ++g) {
          const int n = PLANES * u + g * PPU + pl;
          if ((n < P.N) && (row < H)) {
            const size_t rbase = ((size_t)n * P.C + c) * plane_elems + (size_t)row * W;
#pragma unroll
            for (int j = 0; j < T / E; ++j)
              if (j < PR) {
                if (P.addend) ld_bf16_piece_raw<E>(pfb + g * (T / 2) + j * (E / 2), P.addend + rbase + j * E);
                if constexpr (PFF) {
                  if (P.out_f32 && P.addend_f32) ld_f32_piece_raw<E>(pff + g * T + j * E, P.addend_f32 + rbase + j * E);
                }
              }
            if constexpr (!PFF) {       // T = 64: no registers left for the fp32 addend -> at least pull it into L2
              if (P.out_f32 && P.addend_f32)
                for (int b = 0; b < W * 4; b += 128) prefetch_l2(reinterpret_cast<const uint8_t*>(P.addend_f32 + rbase) + b);
            }
          }
        }
      }
      mbar_wait(BAR(B_ACC_FULL + ab), aph);
      tc_fence_after();
      uint32_t v[T];
      if (has_t) {
        // D_t^T: this thread holds column `row`(=q) over p -> exchange[(pl,p)][band g, q] (float4-chunk XOR swizzle)
#pragma unroll
        for (int g = 0; g < UPS; ++g) {
          tmem_ld_cols<T>(tmem + ((uint32_t)(e * 32) << 16) + ab * Cf::kAccCols + g * 2 * T + T, v);
          tmem_ld_wait();
#pragma unroll
          for (int p = 0; p < T; ++p) {
            const int chunk = (row >> 2) ^ (p & XM);
            ex[(pl * T + p) * 64 + g * T + chunk * 4 + (row & 3)] = __uint_as_float(v[p]);
          }
        }
        named_bar_sync(1, 128);
      }
#pragma unroll
      for (int g = 0; g < UPS; ++g) {
        const int n = PLANES * u + g * PPU + pl;
        const bool ok = (n < P.N) && (row < H);
        const size_t rbase = ((size_t)(n < P.N ? n : 0) * P.C + c) * plane_elems + (size_t)(row < H ? row : 0) * W;
        if (has_n) {
          tmem_ld_cols<T>(tmem + ((uint32_t)(e * 32) << 16) + ab * Cf::kAccCols + g * 2 * T, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int k = 0; k < T; ++k) v[k] = 0u;
        }
        if (g == UPS - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + ab));
        }
        if (GEN && P.bias) {
          const float bv = __ldg(P.bias + c);
#pragma unroll
          for (int k = 0; k < T; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) + bv);
        }
        if (has_t) {
#pragma unroll
          for (int ck = 0; ck < T / 4; ++ck) {
            const float4 t = *reinterpret_cast<const float4*>(&ex[(pl * T + row) * 64 + g * T + ((ck ^ (row & XM)) << 2)]);
            v[4 * ck + 0] = __float_as_uint(__uint_as_float(v[4 * ck + 0]) + t.x);
            v[4 * ck + 1] = __float_as_uint(__uint_as_float(v[4 * ck + 1]) + t.y);
            v[4 * ck + 2] = __float_as_uint(__uint_as_float(v[4 * ck + 2]) + t.z);
            v[4 * ck + 3] = __float_as_uint(__uint_as_float(v[4 * ck + 3]) + t.w);
          }
        }
        if (ok) {
#pragma unroll
          for (int j = 0; j < T / E; ++j)
            if (j < PR) {
              if (P.addend) {
                if constexpr (PFB) add_bf16_raw<E>(v + j * E, pfb + g * (T / 2) + j * (E / 2));
                else add_bf16_piece<E>(v + j * E, P.addend + rbase + j * E);
              }
              if (P.out_f32) {
                if (P.addend_f32) {
                  if constexpr (PFF) add_f32_raw<E>(v + j * E, pff + g * T + j * E);
                  else add_f32_piece<E>(v + j * E, P.addend_f32 + rbase + j * E);
                }
                store_f32_piece<E>(P.out_f32 + rbase + j * E, v + j * E);
              } else {
                store_bf16_piece<E>(P.out + rbase + j * E, v + j * E);
              }
            }
        }
      }
      if (has_t) named_bar_sync(1, 128);        // exchange free for the next unit
    }
  } else if (warp == 10) {
    // ================= Toeplitz builder (multi-channel classes) =================
    float* wts = reinterpret_cast<float*>(sm + Cf::kOffW);
    float* wns = wts + KL * 5;
    for (int c = (NT == 1 ? c_last + 1 : c_first), k = 0; c <= c_last; ++c, ++k) {
      const int set = k % NT;
      mbar_wait(BAR(B_TP_EMPTY + set), ((k / NT) & 1) ^ 1);
      stage_taps(P, c, wts, wns, lane, 32);
      __syncwarp();
      build_toeplitz_pair<T>(sm + Cf::kOffToep + set * Cf::kToepSet, wts, wns, KL, KN, has_t, has_n, lane, 32);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_TP_FULL + set));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cf::kTmemCols>(tmem);
}

template <int T, int CB, bool TMA>
static int launch_dgrad(const CUtensorMap& mt, const CUtensorMap& mn, DgradParams& P, cudaStream_t st) {
  using Cf = dg::Cfg<T>;
  const TcPlan plan = tc_plan(P.N, P.C, T, Cf::PLANES);
  P.units_per_c = plan.units_per_c;
  P.per_cta = plan.per_cta;
  P.splits = plan.splits;
  if (P.has_n && !P.bias) {
    auto kern = lk_dgrad_tc_kernel<T, CB, TMA, false>;
    SLAK_SET_MAX_SMEM(kern, Cf::kSmem);
    kern<<<plan.grid, dg::threads(T), Cf::kSmem, st>>>(mt, mn, P);
  } else {
    auto kern = lk_dgrad_tc_kernel<T, CB, TMA, true>;
    SLAK_SET_MAX_SMEM(kern, Cf::kSmem);
    kern<<<plan.grid, dg::threads(T), Cf::kSmem, st>>>(mt, mn, P);
  }
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

// out = conv(in_t, wt [C,KL,5]) + conv(in_n, wn [C,5,KN]) + addend ; in_t/wt may be null together, and so may in_n/wn
int lk_conv_tc(const void* in_t, const float* wt, const void* in_n, const float* wn, const void* addend, void* out,
               const float* addend_f32, float* out_f32, int N, int C, int H, int W, int KL, int KN, int flip,
               cudaStream_t st, const float* bias) {
  const TcShape s = tc_shape(H, W);
  SLAK_REQUIRE(s.tile != 0, SLAK_ERR_UNSUPPORTED, "shape %dx%d not covered by the tensor-core path", H, W);
  SLAK_REQUIRE((KL * 5 + 5 * KN) * 4 <= 4096, SLAK_ERR_UNSUPPORTED, "kernel side %d too large", KL);
  CUtensorMap mt, mn;
  memset(&mt, 0, sizeof(mt)); memset(&mn, 0, sizeof(mn));
  SLAK_REQUIRE(in_t || in_n, SLAK_ERR_BAD_ARG, "lk_conv_tc needs at least one input");
  if (s.tma) {
    int rc;
    if ((rc = make_plane_map(&mn, in_n ? in_n : in_t, N, C, H, W))) return rc;
    if ((rc = make_plane_map(&mt, in_t ? in_t : in_n, N, C, H, W))) return rc;
  }
  DgradParams P;
  P.in_t = (const __nv_bfloat16*)in_t; P.in_n = (const __nv_bfloat16*)in_n;
  P.wt = wt; P.wn = wn; P.addend = (const __nv_bfloat16*)addend; P.out = (__nv_bfloat16*)out;
  P.addend_f32 = addend_f32; P.out_f32 = out_f32; P.bias = bias;
  P.N = N; P.C = C; P.H = H; P.W = W; P.KL = KL; P.KN = KN; P.flip = flip; P.has_t = in_t ? 1 : 0; P.has_n = in_n ? 1 : 0;
  if (s.tile == 64) return launch_dgrad<64, 16, true>(mt, mn, P, st);
  if (s.tile == 32) {
    if (s.cb == 8) return launch_dgrad<32, 8, false>(mt, mn, P, st);
    if (s.cb == 4) return launch_dgrad<32, 4, false>(mt, mn, P, st);
    return launch_dgrad<32, 2, false>(mt, mn, P, st);
  }
  if (s.cb == 4) return launch_dgrad<16, 4, false>(mt, mn, P, st);
  return launch_dgrad<16, 2, false>(mt, mn, P, st);
}

// shapes covered by the tensor-core dgrad/wgrad kernels: the same set as the forward
bool lk3_bwd_tc_supported(int N, int C, int H, int W, int KL) {
  (void)N; (void)C;
  return tc_shape(H, W).tile != 0 && (KL & 1) && KL >= 5 && KL <= 99;
}

}  // namespace tc
}  // namespace slak
