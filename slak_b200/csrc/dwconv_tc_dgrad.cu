// Depthwise data-gradient on the tensor cores (bf16 in/out, fp32 accumulate in TMEM):
//     out = conv(in_t, W_t [KL x 5])  +  conv(in_n, W_n [5 x KN])  +  addend
// with the taps optionally flipped (flip=1: backward_data of the forward convs,
// backward_data_fp32.cu:199-263; dx = sum over the three branches of models/SLaK.py:89-100).
// The same banded-Toeplitz formulation and tile classes (T = 64/32/16, 128/T stacked planes per
// unit) as dwconv_tc_fwd.cu:
//   natural path     D_n  [(plane,p), q] += IN_n [(plane,p+r-2), w] * Tn_r[q, w]
//   transposed path  D_t^T[(plane,q), p] += IN_t^T[(plane,q+s-2), h] * Tt_s[p, h]
// and the epilogue adds D_n + transpose(D_t^T) (+ addend rows read from global) before rounding.
// dx of a Block needs two launches: (in_n = dy3, W_n = 5x5) -> tmp, then
// (in_t = dy1, in_n = dy2, addend = tmp) -> dx; all three Toeplitz sets do not fit in shared memory
// next to a multi-stage input pipeline at T = 64.
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
constexpr int kThreads = 320;
template <int T> struct Cfg {
  static constexpr int PPU = 128 / T;
  static constexpr int KSTEPS = T / 16;
  static constexpr int kToep = 5 * T * 128;
  static constexpr int kOffTn = 0;
  static constexpr int kOffTt = kToep;
  static constexpr int kOffXN = 2 * kToep;
  static constexpr int kOffXS = kOffXN + kNStages * kSlot;
  static constexpr int kOffXT = kOffXS + kSStages * kUnit;
  static constexpr int kOffEx = kOffXT + kSlot;             // fp32 exchange [128][T]
  static constexpr int kExBytes = 128 * T * 4 < 4096 ? 4096 : 128 * T * 4;
  static constexpr int kOffBar = kOffEx + kExBytes;
  static constexpr int kSmem = kOffBar + 256 + 1024;
  static constexpr int kAccCols = 2 * T;                    // D_n: cols 0..T-1, D_t^T: cols T..2T-1
  static constexpr int kTmemCols = (2 * kAccCols <= 64) ? 64 : (2 * kAccCols <= 128 ? 128 : 256);
};
}  // namespace dg

struct DgradParams {
  const __nv_bfloat16* in_t; const __nv_bfloat16* in_n;
  const float* wt;                 // [C][KL][5] taps of the transposed (vertical-long) path, or nullptr
  const float* wn;                 // [C][5][KN] taps of the natural path
  const __nv_bfloat16* addend;     // [N,C,H,W] or nullptr
  const float* addend_f32;         // [N,C,H,W] fp32 or nullptr (e.g. the shortcut gradient of a Block)
  __nv_bfloat16* out;              // bf16 result, or nullptr when out_f32 is given
  float* out_f32;
  int N, C, H, W, KL, KN, flip, has_t, splits, units_per_c;
};

// same loader as the forward kernel (declared there as a template; repeated to keep the TU standalone)
template <int T, int CB>
__device__ __forceinline__ void dg_load_unit(const __nv_bfloat16* __restrict__ x, uint32_t tile, int n0, int c,
                                             int N, int C, int H, int W, int lane) {
  constexpr int PPU = 128 / T;
  const int PR = (W * 2) / CB;
  const int per_plane = H * PR;
  const size_t plane_bytes = (size_t)H * W * 2;
  for (int pl = 0; pl < PPU; ++pl) {
    const int n = n0 + pl;
    if (n >= N) break;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(x) + ((size_t)n * C + c) * plane_bytes;
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

template <int T, int CB, bool TMA>
__global__ void __launch_bounds__(dg::kThreads, 1)
lk_dgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap nmap, DgradParams P) {
  using namespace dg;
  using Cf = Cfg<T>;
  constexpr int PPU = Cf::PPU, KSTEPS = Cf::KSTEPS, E = CB / 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x / P.splits, split = blockIdx.x % P.splits;
  const int u_begin = (int)(((long long)P.units_per_c * split) / P.splits);
  const int u_end = (int)(((long long)P.units_per_c * (split + 1)) / P.splits);
  const int n_units = u_end - u_begin;
  const int KL = P.KL, KN = P.KN, H = P.H, W = P.W;
  const bool has_t = P.has_t != 0;

  constexpr int B_N_FULL = 0, B_N_EMPTY = kNStages, B_S_FULL = 2 * kNStages, B_S_EMPTY = B_S_FULL + kSStages,
                B_T_FULL = B_S_EMPTY + kSStages, B_T_EMPTY = B_T_FULL + 1, B_ACC_FULL = B_T_EMPTY + 1,
                B_ACC_EMPTY = B_ACC_FULL + kAccBufs;
  const uint32_t bar0 = base + Cf::kOffBar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + Cf::kOffBar + 192);

  if (tid == 0) {
    for (int s = 0; s < kNStages; ++s) { mbar_init(BAR(B_N_FULL + s), 1); mbar_init(BAR(B_N_EMPTY + s), 1); }
    for (int s = 0; s < kSStages; ++s) { mbar_init(BAR(B_S_FULL + s), 1); mbar_init(BAR(B_S_EMPTY + s), kNumTransposerWarps); }
    mbar_init(BAR(B_T_FULL), kNumTransposerWarps);
    mbar_init(BAR(B_T_EMPTY), 1);
    for (int a = 0; a < kAccBufs; ++a) { mbar_init(BAR(B_ACC_FULL + a), 1); mbar_init(BAR(B_ACC_EMPTY + a), 4); }
    mbar_fence_init();
    if (TMA) { tma_prefetch_desc(&nmap); if (has_t) tma_prefetch_desc(&tmap); }
  }
  {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < (Cf::kOffEx - Cf::kOffXN) / 16; i += kThreads) reinterpret_cast<uint4*>(sm + Cf::kOffXN)[i] = z;
    // taps -> fp32 staging (exchange area), flipped when asked
    float* wts = reinterpret_cast<float*>(sm + Cf::kOffEx);  // [KL][5]
    float* wns = wts + KL * 5;                               // [5][KN]
    if (has_t)
      for (int i = tid; i < KL * 5; i += kThreads) {
        const int t = i / 5, s = i - t * 5;
        const int src = P.flip ? ((KL - 1 - t) * 5 + (4 - s)) : i;
        wts[i] = P.wt[(size_t)c * KL * 5 + src];
      }
    for (int i = tid; i < 5 * KN; i += kThreads) {
      const int r = i / KN, t = i - r * KN;
      const int src = P.flip ? ((4 - r) * KN + (KN - 1 - t)) : i;
      wns[i] = P.wn[(size_t)c * 5 * KN + src];
    }
    __syncthreads();
    const int padn = KN / 2, padt = KL / 2;
    constexpr int CH = T / 8;
    for (int ch = tid; ch < 5 * T * CH; ch += kThreads) {
      const int s = ch / (T * CH), rem = ch - s * (T * CH), row = rem / CH, k8 = rem - row * CH;
      float vn[8], vt[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int tn = (k8 * 8 + j) - row + padn;
        vn[j] = (tn >= 0 && tn < KN) ? wns[s * KN + tn] : 0.f;
        const int tt = (k8 * 8 + j) - row + padt;
        vt[j] = (has_t && tt >= 0 && tt < KL) ? wts[tt * 5 + s] : 0.f;
      }
      const uint32_t off = s * (T * 128) + row * 128 + ((k8 ^ (row & 7)) << 4);
      *reinterpret_cast<uint4*>(sm + Cf::kOffTn + off) =
          make_uint4(pack_bf16(vn[0], vn[1]), pack_bf16(vn[2], vn[3]), pack_bf16(vn[4], vn[5]), pack_bf16(vn[6], vn[7]));
      *reinterpret_cast<uint4*>(sm + Cf::kOffTt + off) =
          make_uint4(pack_bf16(vt[0], vt[1]), pack_bf16(vt[2], vt[3]), pack_bf16(vt[4], vt[5]), pack_bf16(vt[6], vt[7]));
    }
  }
  fence_proxy_async();
  if (warp == 2) tmem_alloc<Cf::kTmemCols>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const bool is_loader = (warp == 0) || (!TMA && warp >= 8);
  if (is_loader) {
    if constexpr (TMA) {
      if (elect_one()) {
        for (int i = 0; i < n_units; ++i) {
          const int n0 = PPU * (u_begin + i);
          {
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
      // three cp.async loader warps; loader j owns natural stage j; the source slots alternate per unit
      const int lj = (warp == 0) ? 0 : (warp - 7);
      for (int i = lj; i < n_units; i += kNStages) {
        const int n0 = PPU * (u_begin + i);
        const int st = lj, ph = (i / kNStages) & 1;
        mbar_wait(BAR(B_N_EMPTY + st), ph ^ 1);
        dg_load_unit<T, CB>(P.in_n, base + Cf::kOffXN + st * kSlot + kPad, n0, c, P.N, P.C, H, W, lane);
        const int ss = i % kSStages, sph = (i / kSStages) & 1;
        if (has_t) {
          mbar_wait(BAR(B_S_EMPTY + ss), sph ^ 1);
          dg_load_unit<T, CB>(P.in_t, base + Cf::kOffXS + ss * kUnit, n0, c, P.N, P.C, H, W, lane);
        }
        cp_async_commit();
        cp_async_wait_all();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(BAR(B_N_FULL + st));
          if (has_t) mbar_arrive(BAR(B_S_FULL + ss));
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, T);
      for (int i = 0; i < n_units; ++i) {
        const int st = i % kNStages, ph = (i / kNStages) & 1;
        const int ab = i % kAccBufs, aph = (i / kAccBufs) & 1;
        mbar_wait(BAR(B_ACC_EMPTY + ab), aph ^ 1);
        mbar_wait(BAR(B_N_FULL + st), ph);
        tc_fence_after();
        const uint32_t xn = base + Cf::kOffXN + st * kSlot + kPad;
        const uint32_t dn = tmem + ab * Cf::kAccCols;
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k)
            umma_bf16(dn, umma_desc_k_sw128(xn + (r - 2) * 128 + k * 32, 0),
                      umma_desc_k_sw128(base + Cf::kOffTn + r * (T * 128) + k * 32, 0), idesc, (r | k) != 0);
        umma_commit(BAR(B_N_EMPTY + st));
        if (has_t) {
          mbar_wait(BAR(B_T_FULL), i & 1);
          tc_fence_after();
          const uint32_t xt = base + Cf::kOffXT + kPad;
#pragma unroll
          for (int s = 0; s < 5; ++s)
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k)
              umma_bf16(dn + T, umma_desc_k_sw128(xt + (s - 2) * 128 + k * 32, 0),
                        umma_desc_k_sw128(base + Cf::kOffTt + s * (T * 128) + k * 32, 0), idesc, (s | k) != 0);
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
      for (int i = 0; i < n_units; ++i) {
        const int st = i % kSStages, ph = (i / kSStages) & 1;
        mbar_wait(BAR(B_S_FULL + st), ph);
        mbar_wait(BAR(B_T_EMPTY), (i & 1) ^ 1);
        const uint32_t xs = base + Cf::kOffXS + st * kUnit;
        const uint32_t xt = base + Cf::kOffXT + kPad;
#pragma unroll 4
        for (int it = tw; it < T / 2; it += kNumTransposerWarps) {
          const int blk = 4 * it + m;
          const int pl = blk / (NB * NB), rem = blk - pl * (NB * NB);
          const int bi = rem / NB, bj = rem - bi * NB;
          const uint32_t src = xs + (pl * T + 8 * bi + kk) * 128 + ((bj ^ kk) << 4);
          const uint32_t dst = xt + (pl * T + 8 * bj + kk) * 128 + ((bi ^ kk) << 4);
          uint32_t r0, r1, r2, r3;
          ldmatrix_x4_trans(src, r0, r1, r2, r3);
          stmatrix_x4(dst, r0, r1, r2, r3);
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
    float* ex = reinterpret_cast<float*>(sm + Cf::kOffEx);
    const int PR = W / E;
    constexpr int XM = T / 4 - 1;               // float4-chunk XOR mask of the exchange rows
    for (int i = 0; i < n_units; ++i) {
      const int ab = i % kAccBufs, aph = (i / kAccBufs) & 1;
      const int n = PPU * (u_begin + i) + pl;
      const bool ok = (n < P.N) && (row < H);
      const size_t rbase = ((size_t)(n < P.N ? n : 0) * P.C + c) * plane_elems + (size_t)(row < H ? row : 0) * W;
      mbar_wait(BAR(B_ACC_FULL + ab), aph);
      tc_fence_after();
      const uint32_t t0 = tmem + ((uint32_t)(e * 32) << 16) + ab * Cf::kAccCols;
      uint32_t v[T];
      if (has_t) {
        // D_t^T: this thread holds column `row`(=q) over p -> exchange[(pl,p)][q] (float4-chunk XOR swizzle)
        tmem_ld_cols<T>(t0 + T, v);
        tmem_ld_wait();
#pragma unroll
        for (int p = 0; p < T; ++p) {
          const int chunk = (row >> 2) ^ (p & XM);
          ex[(pl * T + p) * T + chunk * 4 + (row & 3)] = __uint_as_float(v[p]);
        }
      }
      tmem_ld_cols<T>(t0, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + ab));
      if (has_t) {
        named_bar_sync(1, 128);
#pragma unroll
        for (int ck = 0; ck < T / 4; ++ck) {
          const float4 t = *reinterpret_cast<const float4*>(&ex[(pl * T + row) * T + ((ck ^ (row & XM)) << 2)]);
          v[4 * ck + 0] = __float_as_uint(__uint_as_float(v[4 * ck + 0]) + t.x);
          v[4 * ck + 1] = __float_as_uint(__uint_as_float(v[4 * ck + 1]) + t.y);
          v[4 * ck + 2] = __float_as_uint(__uint_as_float(v[4 * ck + 2]) + t.z);
          v[4 * ck + 3] = __float_as_uint(__uint_as_float(v[4 * ck + 3]) + t.w);
        }
        named_bar_sync(1, 128);        // exchange free for the next unit
      }
      if (ok) {
#pragma unroll
        for (int j = 0; j < T / E; ++j)
          if (j < PR) {
            if (P.addend) add_bf16_piece<E>(v + j * E, P.addend + rbase + j * E);
            if (P.out_f32) {
              if (P.addend_f32) add_f32_piece<E>(v + j * E, P.addend_f32 + rbase + j * E);
              store_f32_piece<E>(P.out_f32 + rbase + j * E, v + j * E);
            } else {
              store_bf16_piece<E>(P.out + rbase + j * E, v + j * E);
            }
          }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cf::kTmemCols>(tmem);
}

template <int T, int CB, bool TMA>
static int launch_dgrad(const CUtensorMap& mt, const CUtensorMap& mn, DgradParams& P, cudaStream_t st) {
  using Cf = dg::Cfg<T>;
  P.units_per_c = (P.N + Cf::PPU - 1) / Cf::PPU;
  P.splits = tc_pick_splits(P.C, P.units_per_c);
  auto kern = lk_dgrad_tc_kernel<T, CB, TMA>;
  SLAK_SET_MAX_SMEM(kern, Cf::kSmem);
  kern<<<P.C * P.splits, dg::kThreads, Cf::kSmem, st>>>(mt, mn, P);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

// out = conv(in_t, wt [C,KL,5]) + conv(in_n, wn [C,5,KN]) + addend ; in_t/wt may be null together
int lk_conv_tc(const void* in_t, const float* wt, const void* in_n, const float* wn, const void* addend, void* out,
               const float* addend_f32, float* out_f32, int N, int C, int H, int W, int KL, int KN, int flip,
               cudaStream_t st) {
  const TcShape s = tc_shape(H, W);
  SLAK_REQUIRE(s.tile != 0, SLAK_ERR_UNSUPPORTED, "shape %dx%d not covered by the tensor-core path", H, W);
  CUtensorMap mt, mn;
  memset(&mt, 0, sizeof(mt)); memset(&mn, 0, sizeof(mn));
  if (s.tma) {
    int rc;
    if ((rc = make_plane_map(&mn, in_n, N, C, H, W))) return rc;
    if ((rc = make_plane_map(&mt, in_t ? in_t : in_n, N, C, H, W))) return rc;
  }
  DgradParams P;
  P.in_t = (const __nv_bfloat16*)in_t; P.in_n = (const __nv_bfloat16*)in_n;
  P.wt = wt; P.wn = wn; P.addend = (const __nv_bfloat16*)addend; P.out = (__nv_bfloat16*)out;
  P.addend_f32 = addend_f32; P.out_f32 = out_f32;
  P.N = N; P.C = C; P.H = H; P.W = W; P.KL = KL; P.KN = KN; P.flip = flip; P.has_t = in_t ? 1 : 0;
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
  return tc_shape(H, W).tile != 0 && (KL & 1) && KL >= 5 && KL <= 129;
}

}  // namespace tc
}  // namespace slak
