// Depthwise data-gradient on the tensor cores (bf16 in/out, fp32 accumulate in TMEM):
//     out = conv(in_t, W_t [KL x 5])  +  conv(in_n, W_n [5 x KN])  +  addend
// with the taps optionally flipped (flip=1: backward_data of the forward convs,
// backward_data_fp32.cu:199-263; dx = sum over the three branches of models/SLaK.py:89-100).
// The same banded-Toeplitz formulation as dwconv_tc_fwd.cu:
//   natural path     D_n  [(plane,p), q] += IN_n [(plane,p+r-2), w] * Tn_r[q, w]      (5 MMAs x 4 k-steps)
//   transposed path  D_t^T[(plane,q), p] += IN_t^T[(plane,q+s-2), h] * Tt_s[p, h]
// and the epilogue adds D_n + transpose(D_t^T) (+ addend rows read from global) before rounding.
// dx of a Block needs two launches: (in_n = dy3, W_n = 5x5) -> tmp, then
// (in_t = dy1, in_n = dy2, addend = tmp) -> dx; all three Toeplitz sets do not fit in shared memory
// next to a multi-stage input pipeline.
#include "common.cuh"
#include "tc_common.cuh"

namespace slak {
namespace tc {

int make_plane_map(CUtensorMap* map, const void* x, int N, int C, int H, int W);  // dwconv_tc_fwd.cu

namespace dg {
constexpr int kNStages = 3;                    // natural-path input slots
constexpr int kSStages = 2;                    // transposed-path source slots (natural layout)
constexpr int kAccBufs = 2;
constexpr int kTile = 64 * 128;
constexpr int kUnit = 2 * kTile;
constexpr int kPad = 1024;
constexpr int kSlot = kPad + kUnit + kPad;     // 18 KB
constexpr int kOffTn = 0;                      // 5 x 8 KB
constexpr int kOffTt = kOffTn + 5 * kTile;     // 5 x 8 KB
constexpr int kOffXN = kOffTt + 5 * kTile;     // 80 KB
constexpr int kOffXS = kOffXN + kNStages * kSlot;
constexpr int kOffXT = kOffXS + kSStages * kUnit;
constexpr int kOffEx = kOffXT + kSlot;         // fp32 exchange [128][64]
constexpr int kOffBar = kOffEx + 128 * 64 * 4;
constexpr int kSmemBytes = kOffBar + 256 + 1024;
constexpr int kAccCols = 128;                  // D_n: cols 0..63, D_t^T: cols 64..127
constexpr int kNumTransposerWarps = 2;
}  // namespace dg

struct DgradParams {
  const float* wt;                 // [C][KL][5] taps of the transposed (vertical-long) path, or nullptr
  const float* wn;                 // [C][5][KN] taps of the natural path
  const __nv_bfloat16* addend;     // [N,C,H,W] or nullptr
  __nv_bfloat16* out;
  int N, C, H, W, KL, KN, flip, has_t, splits, pairs_per_c;
};

__device__ __forceinline__ uint32_t pack_bf16_(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(256, 1)
lk_dgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap nmap, DgradParams P) {
  using namespace dg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x / P.splits, split = blockIdx.x % P.splits;
  const int u_begin = (int)(((long long)P.pairs_per_c * split) / P.splits);
  const int u_end = (int)(((long long)P.pairs_per_c * (split + 1)) / P.splits);
  const int n_units = u_end - u_begin;
  const int KL = P.KL, KN = P.KN, H = P.H, W = P.W;
  const bool has_t = P.has_t != 0;

  constexpr int B_N_FULL = 0, B_N_EMPTY = kNStages, B_S_FULL = 2 * kNStages, B_S_EMPTY = B_S_FULL + kSStages,
                B_T_FULL = B_S_EMPTY + kSStages, B_T_EMPTY = B_T_FULL + 1, B_ACC_FULL = B_T_EMPTY + 1,
                B_ACC_EMPTY = B_ACC_FULL + kAccBufs;
  const uint32_t bar0 = base + kOffBar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kOffBar + 192);

  if (tid == 0) {
    for (int s = 0; s < kNStages; ++s) { mbar_init(BAR(B_N_FULL + s), 1); mbar_init(BAR(B_N_EMPTY + s), 1); }
    for (int s = 0; s < kSStages; ++s) { mbar_init(BAR(B_S_FULL + s), 1); mbar_init(BAR(B_S_EMPTY + s), kNumTransposerWarps); }
    mbar_init(BAR(B_T_FULL), kNumTransposerWarps);
    mbar_init(BAR(B_T_EMPTY), 1);
    for (int a = 0; a < kAccBufs; ++a) { mbar_init(BAR(B_ACC_FULL + a), 1); mbar_init(BAR(B_ACC_EMPTY + a), 4); }
    mbar_fence_init();
    tma_prefetch_desc(&nmap);
    if (has_t) tma_prefetch_desc(&tmap);
  }
  {
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int s = 0; s < kNStages; ++s) {
      uint8_t* slot = sm + kOffXN + s * kSlot;
      for (int i = tid; i < kPad / 16; i += 256) {
        reinterpret_cast<uint4*>(slot)[i] = z;
        reinterpret_cast<uint4*>(slot + kPad + kUnit)[i] = z;
      }
    }
    uint8_t* slot = sm + kOffXT;
    for (int i = tid; i < kPad / 16; i += 256) {
      reinterpret_cast<uint4*>(slot)[i] = z;
      reinterpret_cast<uint4*>(slot + kPad + kUnit)[i] = z;
    }
    // taps -> fp32 staging (exchange area), flipped when asked
    float* wts = reinterpret_cast<float*>(sm + kOffEx);      // [KL][5]
    float* wns = wts + KL * 5;                               // [5][KN]
    if (has_t)
      for (int i = tid; i < KL * 5; i += 256) {
        const int t = i / 5, s = i - t * 5;
        const int src = P.flip ? ((KL - 1 - t) * 5 + (4 - s)) : i;
        wts[i] = P.wt[(size_t)c * KL * 5 + src];
      }
    for (int i = tid; i < 5 * KN; i += 256) {
      const int r = i / KN, t = i - r * KN;
      const int src = P.flip ? ((4 - r) * KN + (KN - 1 - t)) : i;
      wns[i] = P.wn[(size_t)c * 5 * KN + src];
    }
    __syncthreads();
    const int padn = KN / 2, padt = KL / 2;
    for (int ch = tid; ch < 5 * 64 * 8; ch += 256) {
      const int s = ch / 512, rem = ch - s * 512, row = rem >> 3, k8 = rem & 7;
      float vn[8], vt[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int tn = (k8 * 8 + j) - row + padn;
        vn[j] = (tn >= 0 && tn < KN) ? wns[s * KN + tn] : 0.f;
        const int tt = (k8 * 8 + j) - row + padt;
        vt[j] = (has_t && tt >= 0 && tt < KL) ? wts[tt * 5 + s] : 0.f;
      }
      const uint32_t off = s * kTile + row * 128 + ((k8 ^ (row & 7)) << 4);
      *reinterpret_cast<uint4*>(sm + kOffTn + off) =
          make_uint4(pack_bf16_(vn[0], vn[1]), pack_bf16_(vn[2], vn[3]), pack_bf16_(vn[4], vn[5]), pack_bf16_(vn[6], vn[7]));
      *reinterpret_cast<uint4*>(sm + kOffTt + off) =
          make_uint4(pack_bf16_(vt[0], vt[1]), pack_bf16_(vt[2], vt[3]), pack_bf16_(vt[4], vt[5]), pack_bf16_(vt[6], vt[7]));
    }
  }
  fence_proxy_async();
  if (warp == 2) tmem_alloc<256>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      for (int i = 0; i < n_units; ++i) {
        const int n0 = 2 * (u_begin + i);
        const int pa = n0 * P.C + c, pb = min(n0 + 1, P.N - 1) * P.C + c;
        {
          const int st = i % kNStages, ph = (i / kNStages) & 1;
          mbar_wait(BAR(B_N_EMPTY + st), ph ^ 1);
          const uint32_t dst = base + kOffXN + st * kSlot + kPad;
          mbar_expect_tx(BAR(B_N_FULL + st), kUnit);
          tma_load_3d(dst, &nmap, BAR(B_N_FULL + st), 0, 0, pa);
          tma_load_3d(dst + kTile, &nmap, BAR(B_N_FULL + st), 0, 0, pb);
        }
        if (has_t) {
          const int st = i % kSStages, ph = (i / kSStages) & 1;
          mbar_wait(BAR(B_S_EMPTY + st), ph ^ 1);
          const uint32_t dst = base + kOffXS + st * kUnit;
          mbar_expect_tx(BAR(B_S_FULL + st), kUnit);
          tma_load_3d(dst, &tmap, BAR(B_S_FULL + st), 0, 0, pa);
          tma_load_3d(dst + kTile, &tmap, BAR(B_S_FULL + st), 0, 0, pb);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
      for (int i = 0; i < n_units; ++i) {
        const int st = i % kNStages, ph = (i / kNStages) & 1;
        const int ab = i % kAccBufs, aph = (i / kAccBufs) & 1;
        mbar_wait(BAR(B_ACC_EMPTY + ab), aph ^ 1);
        mbar_wait(BAR(B_N_FULL + st), ph);
        tc_fence_after();
        const uint32_t xn = base + kOffXN + st * kSlot + kPad;
        const uint32_t dn = tmem + ab * kAccCols;
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(dn, umma_desc_k_sw128(xn + (r - 2) * 128 + k * 32, 0),
                      umma_desc_k_sw128(base + kOffTn + r * kTile + k * 32, 0), idesc, (r | k) != 0);
        umma_commit(BAR(B_N_EMPTY + st));
        if (has_t) {
          mbar_wait(BAR(B_T_FULL), i & 1);
          tc_fence_after();
          const uint32_t xt = base + kOffXT + kPad;
#pragma unroll
          for (int s = 0; s < 5; ++s)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(dn + 64, umma_desc_k_sw128(xt + (s - 2) * 128 + k * 32, 0),
                        umma_desc_k_sw128(base + kOffTt + s * kTile + k * 32, 0), idesc, (s | k) != 0);
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
      for (int i = 0; i < n_units; ++i) {
        const int st = i % kSStages, ph = (i / kSStages) & 1;
        mbar_wait(BAR(B_S_FULL + st), ph);
        mbar_wait(BAR(B_T_EMPTY), (i & 1) ^ 1);
        const uint32_t xs = base + kOffXS + st * kUnit;
        const uint32_t xt = base + kOffXT + kPad;
#pragma unroll 4
        for (int it = tw; it < 32; it += kNumTransposerWarps) {
          const int h = it >> 4, bi = (it >> 1) & 7, g = it & 1;
          const int bj = 4 * g + m;
          const uint32_t src = xs + (64 * h + 8 * bi + kk) * 128 + ((bj ^ kk) << 4);
          const uint32_t dst = xt + (64 * h + 8 * bj + kk) * 128 + ((bi ^ kk) << 4);
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
  } else {
    // ================= epilogue =================
    const int e = warp - 4;
    const int L = e * 32 + lane;
    const int half = L >> 6, row = L & 63;
    const size_t plane_elems = (size_t)H * W;
    float* ex = reinterpret_cast<float*>(sm + kOffEx);
    const int wchunks = W >> 3;
    for (int i = 0; i < n_units; ++i) {
      const int ab = i % kAccBufs, aph = (i / kAccBufs) & 1;
      const int n = 2 * (u_begin + i) + half;
      const bool ok = (n < P.N) && (row < H);
      const size_t rbase = ((size_t)(n < P.N ? n : 0) * P.C + c) * plane_elems + (size_t)(row < H ? row : 0) * W;
      mbar_wait(BAR(B_ACC_FULL + ab), aph);
      tc_fence_after();
      const uint32_t t0 = tmem + ((uint32_t)(e * 32) << 16) + ab * kAccCols;
      uint32_t v[64];
      if (has_t) {
        // D_t^T: this thread holds column `row`(=q) over p -> exchange[(half,p)][q] (float4-chunk XOR swizzle)
        tmem_ld32(t0 + 64, v);
        tmem_ld32(t0 + 96, v + 32);
        tmem_ld_wait();
#pragma unroll
        for (int p = 0; p < 64; ++p) {
          const int chunk = (row >> 2) ^ (p & 15);
          ex[(half * 64 + p) * 64 + chunk * 4 + (row & 3)] = __uint_as_float(v[p]);
        }
      }
      tmem_ld32(t0, v);
      tmem_ld32(t0 + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + ab));
      if (has_t) {
        named_bar_sync(1, 128);
#pragma unroll
        for (int ck = 0; ck < 16; ++ck) {
          const float4 t = *reinterpret_cast<const float4*>(&ex[(half * 64 + row) * 64 + ((ck ^ (row & 15)) << 2)]);
          v[4 * ck + 0] = __float_as_uint(__uint_as_float(v[4 * ck + 0]) + t.x);
          v[4 * ck + 1] = __float_as_uint(__uint_as_float(v[4 * ck + 1]) + t.y);
          v[4 * ck + 2] = __float_as_uint(__uint_as_float(v[4 * ck + 2]) + t.z);
          v[4 * ck + 3] = __float_as_uint(__uint_as_float(v[4 * ck + 3]) + t.w);
        }
        named_bar_sync(1, 128);        // exchange free for the next unit
      }
      if (ok) {
#pragma unroll
        for (int ck = 0; ck < 8; ++ck) {
          if (ck < wchunks) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[8 * ck + j]);
            if (P.addend) {
              const uint4 a = *reinterpret_cast<const uint4*>(P.addend + rbase + 8 * ck);
              const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 t = __bfloat1622float2(ap[j]);
                f[2 * j] += t.x; f[2 * j + 1] += t.y;
              }
            }
            uint4 o = make_uint4(pack_bf16_(f[0], f[1]), pack_bf16_(f[2], f[3]), pack_bf16_(f[4], f[5]), pack_bf16_(f[6], f[7]));
            *reinterpret_cast<uint4*>(P.out + rbase + 8 * ck) = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<256>(tmem);
}

static int pick_splits(int C, int pairs) {
  const int sms = sm_count();
  int best = 1; double best_eff = 0.0;
  const int max_s = pairs >= 8 ? pairs / 8 : 1;
  for (int s = 1; s <= max_s && s <= 64; ++s) {
    const long long ctas = (long long)C * s;
    const long long waves = (ctas + sms - 1) / sms;
    const int per = (pairs + s - 1) / s;
    const double eff = (double)C * pairs / ((double)waves * sms * per) * (per / (per + 1.5));
    if (eff > best_eff) { best_eff = eff; best = s; }
  }
  return best;
}

// out = conv(in_t, wt [C,KL,5]) + conv(in_n, wn [C,5,KN]) + addend ; in_t/wt may be null together
int lk_conv_tc(const void* in_t, const float* wt, const void* in_n, const float* wn, const void* addend, void* out,
               int N, int C, int H, int W, int KL, int KN, int flip, cudaStream_t st) {
  CUtensorMap mt, mn;
  int rc;
  if ((rc = make_plane_map(&mn, in_n, N, C, H, W))) return rc;
  if ((rc = make_plane_map(&mt, in_t ? in_t : in_n, N, C, H, W))) return rc;
  DgradParams P;
  P.wt = wt; P.wn = wn; P.addend = (const __nv_bfloat16*)addend; P.out = (__nv_bfloat16*)out;
  P.N = N; P.C = C; P.H = H; P.W = W; P.KL = KL; P.KN = KN; P.flip = flip; P.has_t = in_t ? 1 : 0;
  P.pairs_per_c = (N + 1) / 2;
  P.splits = pick_splits(C, P.pairs_per_c);
  SLAK_CUDA_TRY(cudaFuncSetAttribute(lk_dgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dg::kSmemBytes));
  lk_dgrad_tc_kernel<<<C * P.splits, 256, dg::kSmemBytes, st>>>(mt, mn, P);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

}  // namespace tc
}  // namespace slak

namespace slak { namespace tc {
// shapes covered by the tensor-core dgrad/wgrad kernels (64-class planes staged by TMA)
bool lk3_bwd_tc_supported(int N, int C, int H, int W, int KL) {
  (void)N; (void)C;
  const TcShape s = tc_shape(H, W);
  return s.tile == 64 && s.tma && H >= 8 && W >= 8 && (KL & 1) && KL >= 5 && KL * 5 * 2 + 25 <= 4000;
}
} }
