// Pointwise MLP of a SLaK Block on the tensor cores (models/SLaK.py:157-160: pwconv1 -> GELU -> pwconv2, and its
// backward): every GEMM of the Block is a tcgen05 kernel of this file, with the elementwise work of the reference's
// separate kernels (bias add, exact-erf GELU, GELU', bias gradients, fp32 -> bf16 casts) folded into the epilogues.
//
//   gemm_nt<EPI>   D[M,N] = A[M,K] B[N,K]^T   both operands K-major (row-major, K contiguous), bf16, fp32 accumulate
//        EPI_FC1    H = D + b1 (bf16, as nn.Linear under autocast), A = gelu(H) (exact erf GELU of the ROUNDED H, as
//                   F.gelu on the bf16 tensor); X read once, H and A written once
//        EPI_BIAS   H2 = D + b2 (bf16)
//        EPI_DGELU  dH = D * gelu'(H);  per-CTA partial column sums of dH (bias gradient of pwconv1); dA = dH2 W2
//                   never goes to HBM
//        EPI_PLAIN  dXn = D (bf16)
//   gemm_tn_splitk D[Ma,Nb] = P[M,Ma]^T Q[M,Nb]  contraction over the TOKENS (weight gradients): both operands are
//                   row-major activations, i.e. MN-major MMA operands exactly as TMA delivers them; every CTA
//                   accumulates one 128 x BN output tile over its token range in TMEM and writes an fp32 partial;
//                   partials are folded in a fixed order (slak_colsum_f32): deterministic.
//
// Structure of gemm_nt: persistent, warp-specialised; 128 x BN output tiles (BN = 256 keeps the shared-memory operand
// traffic at 96 B/clk, below the 128 B/clk port; BN = 128 for narrow N), K in 64-wide SWIZZLE_128B blocks through a TMA
// ring, fp32 accumulators double-buffered in TMEM (2 x BN columns), two epilogue warpgroups that each own half of
// the tile's 64-column slabs: TMEM -> registers -> epilogue math -> swizzled staging slab -> TMA tile store.
// Warp roles (384 threads): w0 TMA producer | w1 MMA issuer | w2 TMEM allocator | w4-7, w8-11 epilogue.
#include "common.cuh"
#include "tc_common.cuh"
#include <string.h>

namespace slak {
namespace tc {

int make_plane_map(CUtensorMap* map, const void* x, int N, int C, int H, int W);  // dwconv_tc_fwd.cu: (W, H, N*C), box 64x64x1

namespace mlp {
constexpr int BM = 128, BK = 64;
constexpr int kBox = 64 * 128;                  // one 64-row x 128-byte TMA box (8 KB)
constexpr int kThreads = 384;
enum Epi { EPI_FC1 = 0, EPI_BIAS = 1, EPI_DGELU = 2, EPI_PLAIN = 3 };

template <int BN> struct Cfg {
  static constexpr int kStages = BN == 256 ? 3 : 4;
  static constexpr int kATile = 2 * kBox;                      // 128 rows x 64 k
  static constexpr int kBTile = (BN / 64) * kBox;
  static constexpr int kStage = kATile + kBTile;
  static constexpr int kOffStg = kStages * kStage;             // staging: [group][2] slabs of 128 rows x 128 B
  static constexpr int kOffCol = kOffStg + 4 * 2 * kBox;       // DGELU: per-CTA column accumulators [N <= 3072] fp32 + scratch
  static constexpr int kColBytes = 3072 * 4 + 8 * 64 * 4;
  static constexpr int kOffBar = kOffCol + kColBytes;
  static constexpr int kSmem = kOffBar + 256 + 1024;
  static constexpr int kSlabsPerGroup = BN / 128;              // 64-column slabs each epilogue group owns
  static_assert(kSmem <= 232448, "shared memory budget");
};

struct Params {
  const float* bias;                // FC1 / BIAS: [N] fp32 (rounded to bf16 on use, as b.to(bf16) in the module path)
  const __nv_bfloat16* h;           // DGELU: [M][N] bf16, the saved pre-activation
  float* colpart;                   // DGELU: [grid][N] per-CTA partial column sums of dH
  int M, N, K;
  int write_h;                      // FC1: also store H (training); 0 = only A (inference)
};

// d/dx [x Phi(x)] and x Phi(x) with Phi through erf's rational approximation (A&S 7.1.26, |err| <= 1.5e-7)
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf) {
  const float e = __expf(-0.5f * x * x);
  const float t = __fdividef(1.f, fmaf(0.3275911f * 0.70710678118654752f, fabsf(x), 1.f));
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  cdf = 0.5f + copysignf(0.5f * fmaf(-poly, e, 1.f), x);
  pdf = 0.39894228040143268f * e;
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
mlp_gemm_nt_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                   const __grid_constant__ CUtensorMap o0map, const __grid_constant__ CUtensorMap o1map, Params P) {
  using C = Cfg<BN>;
  constexpr int kStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int m_tiles = (P.M + BM - 1) / BM, n_tiles = (P.N + BN - 1) / BN;
  const int tiles = m_tiles * n_tiles;
  const int KB = (P.K + BK - 1) / BK;
  const int ksteps_last = ((P.K - (KB - 1) * BK) + 15) / 16;     // k16 steps of the last (partial) K block

  constexpr int B_FULL = 0, B_EMPTY = kStages, B_ACC_FULL = 2 * kStages, B_ACC_EMPTY = B_ACC_FULL + 2;
  const uint32_t bar0 = base + C::kOffBar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::kOffBar + 192);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(BAR(B_FULL + s), 1); mbar_init(BAR(B_EMPTY + s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(BAR(B_ACC_FULL + a), 1); mbar_init(BAR(B_ACC_EMPTY + a), 8); }
    mbar_fence_init();
    tma_prefetch_desc(&amap); tma_prefetch_desc(&bmap); tma_prefetch_desc(&o0map);
    if (EPI == EPI_FC1) tma_prefetch_desc(&o1map);
  }
  if (EPI == EPI_DGELU) {   // column accumulators start at zero
    float* col = reinterpret_cast<float*>(sm + C::kOffCol);
    for (int i = tid; i < P.N; i += kThreads) col[i] = 0.f;
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int kbc = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
        for (int kb = 0; kb < KB; ++kb, ++kbc) {
          const int s = kbc % kStages, ph = (kbc / kStages) & 1;
          mbar_wait(BAR(B_EMPTY + s), ph ^ 1);
          const uint32_t sa = base + s * C::kStage, sb = sa + C::kATile;
          mbar_expect_tx(BAR(B_FULL + s), C::kStage);
          tma_load_3d(sa, &amap, BAR(B_FULL + s), kb * BK, m0, 0);
          tma_load_3d(sa + kBox, &amap, BAR(B_FULL + s), kb * BK, m0 + 64, 0);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_3d(sb + j * kBox, &bmap, BAR(B_FULL + s), kb * BK, n0 + 64 * j, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int kbc = 0, it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const int ab = it & 1, aph = (it >> 1) & 1;
        mbar_wait(BAR(B_ACC_EMPTY + ab), aph ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem + ab * BN;
        for (int kb = 0; kb < KB; ++kb, ++kbc) {
          const int s = kbc % kStages, ph = (kbc / kStages) & 1;
          mbar_wait(BAR(B_FULL + s), ph);
          tc_fence_after();
          const uint32_t sa = base + s * C::kStage, sb = sa + C::kATile;
          const int ks = (kb == KB - 1) ? ksteps_last : BK / 16;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk)
            if (kk < ks)
              umma_bf16(acc, umma_desc_k_sw128(sa + kk * 32, 0), umma_desc_k_sw128(sb + kk * 32, 0), idesc, (kb | kk) != 0);
          umma_commit(BAR(B_EMPTY + s));
        }
        umma_commit(BAR(B_ACC_FULL + ab));
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: group g owns the 64-column slabs [g * SPG, (g + 1) * SPG) of every tile =================
    constexpr int SPG = C::kSlabsPerGroup;
    const int g = (warp - 4) >> 2, e = (warp - 4) & 3;
    const int L = e * 32 + lane;                          // row of the tile = TMEM lane
    const int nb = 1 + g;                                 // named barrier of the group
    uint8_t* stg = sm + C::kOffStg + g * 2 * kBox * 2;     // two slabs [128 rows][128 B] swizzled (each = two 64-row boxes)
    const uint32_t stg_s = base + C::kOffStg + g * 2 * kBox * 2;
    float* col = reinterpret_cast<float*>(sm + C::kOffCol);
    float* scratch = col + 3072;                          // [8 warps][64]
    int it = 0, slab_ctr = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int m0 = (t / n_tiles) * BM, nt0 = (t % n_tiles) * BN;
      const int ab = it & 1, aph = (it >> 1) & 1;
      const int m = m0 + L;
#pragma unroll 1
      for (int sl = 0; sl < SPG; ++sl, ++slab_ctr) {
        const int n0 = nt0 + 64 * (g * SPG + sl);         // first column of the slab
        if (n0 >= P.N) {                                  // slab entirely beyond N (narrow N in a wide tile)
          if (sl == SPG - 1) {
            if (sl == 0) { mbar_wait(BAR(B_ACC_FULL + ab), aph); tc_fence_after(); }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + ab));
          }
          continue;
        }
        uint4 hraw[8];                                    // DGELU: the saved pre-activation row, requested before the wait
        if constexpr (EPI == EPI_DGELU) {
          const uint4* hp = reinterpret_cast<const uint4*>(P.h + (size_t)(m < P.M ? m : 0) * P.N + n0);
#pragma unroll
          for (int j = 0; j < 8; ++j) hraw[j] = (n0 + 8 * j < P.N) ? __ldg(hp + j) : make_uint4(0, 0, 0, 0);
        }
        if (sl == 0) { mbar_wait(BAR(B_ACC_FULL + ab), aph); tc_fence_after(); }
        uint32_t v[64];
        const uint32_t ta = tmem + ((uint32_t)(e * 32) << 16) + ab * BN + 64 * (g * SPG + sl);
        tmem_ld32(ta, v); tmem_ld32(ta + 32, v + 32);
        tmem_ld_wait();
        if (sl == SPG - 1) {                              // accumulators drained (8 warps arrive)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + ab));
        }
        // staging: FC1 uses both slabs of the group (H, A) for one 64-column slab; the single-output epilogues
        // alternate between them, so only the store issued two slabs ago has to have finished reading
        const int buf = (EPI == EPI_FC1) ? 0 : (slab_ctr & 1);
        if (e == 0 && lane == 0) {
          if (EPI == EPI_FC1) bulk_wait_group_read<0>(); else bulk_wait_group_read<1>();
        }
        named_bar_sync(nb, 128);
        uint8_t* s0 = stg + buf * 2 * kBox;
        uint8_t* s1 = stg + 2 * kBox;
        float cs[(EPI == EPI_DGELU) ? 64 : 1];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float o[8], a[8];
          if constexpr (EPI == EPI_FC1 || EPI == EPI_BIAS) {
            float bb[8];
            if (n0 + 8 * j < P.N) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(P.bias + n0 + 8 * j));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(P.bias + n0 + 8 * j + 4));
              bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) bb[k] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float hv = __uint_as_float(v[8 * j + k]) + __bfloat162float(__float2bfloat16_rn(bb[k]));
              o[k] = hv;
              if constexpr (EPI == EPI_FC1) {
                const float hr = __bfloat162float(__float2bfloat16_rn(hv));     // GELU of the value that is stored
                float cdf, pdf;
                gelu_parts(hr, cdf, pdf);
                a[k] = hr * cdf;
              }
            }
          } else if constexpr (EPI == EPI_DGELU) {
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&hraw[j]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 hx = __bfloat1622float2(h2[k]);
              float c0, p0, c1, p1;
              gelu_parts(hx.x, c0, p0);
              gelu_parts(hx.y, c1, p1);
              o[2 * k] = __uint_as_float(v[8 * j + 2 * k]) * fmaf(hx.x, p0, c0);
              o[2 * k + 1] = __uint_as_float(v[8 * j + 2 * k + 1]) * fmaf(hx.y, p1, c1);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
              cs[8 * j + k] = (m < P.M) ? __bfloat162float(__float2bfloat16_rn(o[k])) : 0.f;
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = __uint_as_float(v[8 * j + k]);
          }
          const uint32_t off = (uint32_t)L * 128 + ((j ^ (L & 7)) << 4);
          if (EPI != EPI_FC1 || P.write_h)
            *reinterpret_cast<uint4*>(s0 + off) = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]),
                                                             pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
          if constexpr (EPI == EPI_FC1)
            *reinterpret_cast<uint4*>(s1 + off) = make_uint4(pack_bf16(a[0], a[1]), pack_bf16(a[2], a[3]),
                                                             pack_bf16(a[4], a[5]), pack_bf16(a[6], a[7]));
        }
        fence_proxy_async();
        if constexpr (EPI == EPI_DGELU) {
          // column sums over the warp's 32 rows by a transposing butterfly (lane l ends with column l / 32 + l)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float* c = cs + 32 * hh;
#pragma unroll
            for (int o2 = 16; o2 >= 1; o2 >>= 1) {
              const bool up = (lane & o2) != 0;
#pragma unroll
              for (int k = 0; k < o2; ++k) {
                const float send = up ? c[k] : c[k + o2], keep = up ? c[k + o2] : c[k];
                c[k] = keep + __shfl_xor_sync(0xffffffffu, send, o2);
              }
            }
            scratch[(warp - 4) * 64 + 32 * hh + lane] = c[0];
          }
        }
        named_bar_sync(nb, 128);
        if (EPI == EPI_DGELU && e == 0) {      // one warp folds the group's four row blocks into the CTA's column accumulators
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int cidx = 32 * hh + lane;
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) s += scratch[(4 * g + w) * 64 + cidx];
            if (n0 + cidx < P.N) col[n0 + cidx] += s;
          }
        }
        if (e == 0 && lane == 0) {
          const uint32_t a0 = stg_s + buf * 2 * kBox;
          if (EPI != EPI_FC1 || P.write_h) {
            tma_store_3d(&o0map, a0, n0, m0, 0);
            tma_store_3d(&o0map, a0 + kBox, n0, m0 + 64, 0);
          }
          if (EPI == EPI_FC1) {
            tma_store_3d(&o1map, stg_s + 2 * kBox, n0, m0, 0);
            tma_store_3d(&o1map, stg_s + 2 * kBox + kBox, n0, m0 + 64, 0);
          }
          bulk_commit_group();
        }
        if (EPI == EPI_DGELU) named_bar_sync(nb, 128);    // scratch is rewritten by the next slab
      }
    }
    if (e == 0 && lane == 0) bulk_wait_group_read<0>();   // shared memory must outlive the last tile store
  }

  tc_fence_before();
  __syncthreads();
  if (EPI == EPI_DGELU) {
    const float* col = reinterpret_cast<const float*>(sm + C::kOffCol);
    for (int i = tid; i < P.N; i += kThreads) P.colpart[(size_t)blockIdx.x * P.N + i] = col[i];
  }
  if (warp == 2) tmem_dealloc<512>(tmem);
}

// --------------------------------------------------------------------------------------------------------------
// Weight gradients: D[Ma,Nb] = sum over tokens m of P[m,a] Q[m,b].  Both operands MN-major SWIZZLE_128B: a K block is
// 64 tokens = a 64-row x 128-byte box per 64 columns of P / Q; a k16 step advances the descriptors by 16 rows (2 KB).
// CTA (tile, split): output tile 128 x BN, tokens [split * span, (split + 1) * span).
// Warp roles (192 threads): w0 TMA | w1 MMA + TMEM | w2-5 epilogue (after the last K block).
// --------------------------------------------------------------------------------------------------------------
constexpr int kWgThreads = 192;
template <int BN> struct WgCfg {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kATile = 2 * kBox;                      // 64 tokens x 128 columns of P
  static constexpr int kBTile = (BN / 64) * kBox;              // 64 tokens x BN columns of Q
  static constexpr int kStage = kATile + kBTile;
  static constexpr int kOffBar = kStages * kStage;
  static constexpr int kSmem = kOffBar + 256 + 1024;
  static_assert(kSmem <= 232448, "shared memory budget");
};
struct WgParams {
  float* part;                      // [splits][Ma][Nb] fp32
  int M, Ma, Nb, span;              // span = tokens per split (multiple of 64)
  int tiles_b;                      // ceil(Nb / BN)
};
__device__ __forceinline__ uint64_t umma_desc_mn_sw128_(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn_(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BN>
__global__ void __launch_bounds__(kWgThreads, 1)
mlp_gemm_tn_splitk_kernel(const __grid_constant__ CUtensorMap pmap, const __grid_constant__ CUtensorMap qmap, WgParams P) {
  using C = WgCfg<BN>;
  constexpr int kStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles = gridDim.x;                      // (a tile, b tile) pairs
  const int tile = blockIdx.x, split = blockIdx.y;
  const int a0 = (tile / P.tiles_b) * 128, b0 = (tile % P.tiles_b) * BN;
  const int m_lo = split * P.span;
  int m_hi = m_lo + P.span; if (m_hi > P.M) m_hi = P.M;
  const int KB = m_hi > m_lo ? (m_hi - m_lo + 63) / 64 : 0;
  (void)tiles;

  constexpr int B_FULL = 0, B_EMPTY = kStages, B_ACC = 2 * kStages;
  const uint32_t bar0 = base + C::kOffBar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::kOffBar + 192);
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(BAR(B_FULL + s), 1); mbar_init(BAR(B_EMPTY + s), 1); }
    mbar_init(BAR(B_ACC), 1);
    mbar_fence_init();
    tma_prefetch_desc(&pmap); tma_prefetch_desc(&qmap);
  }
  if (warp == 1) tmem_alloc<BN>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % kStages, ph = (kb / kStages) & 1;
        mbar_wait(BAR(B_EMPTY + s), ph ^ 1);
        const uint32_t sa = base + s * C::kStage, sb = sa + C::kATile;
        const int m = m_lo + kb * 64;
        mbar_expect_tx(BAR(B_FULL + s), C::kStage);
        tma_load_3d(sa, &pmap, BAR(B_FULL + s), a0, m, 0);
        tma_load_3d(sa + kBox, &pmap, BAR(B_FULL + s), a0 + 64, m, 0);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) tma_load_3d(sb + j * kBox, &qmap, BAR(B_FULL + s), b0 + 64 * j, m, 0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16_mn_(128, BN);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % kStages, ph = (kb / kStages) & 1;
        mbar_wait(BAR(B_FULL + s), ph);
        tc_fence_after();
        const uint32_t sa = base + s * C::kStage, sb = sa + C::kATile;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem, umma_desc_mn_sw128_(sa + k * 2048, kBox), umma_desc_mn_sw128_(sb + k * 2048, BN > 64 ? kBox : 0),
                    idesc, (kb | k) != 0);
        umma_commit(BAR(B_EMPTY + s));
      }
      umma_commit(BAR(B_ACC));
    }
  } else {
    // ================= epilogue: TMEM -> fp32 partial tile =================
    const int e = warp & 3;                               // TMEM lane quarter of this warp
    const int a = a0 + e * 32 + lane;                     // output row
    float* dst = P.part + ((size_t)split * P.Ma + (a < P.Ma ? a : 0)) * P.Nb + b0;
    if (KB > 0) {
      mbar_wait(BAR(B_ACC), 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      if (KB > 0) {
        tmem_ld32(tmem + ((uint32_t)(e * 32) << 16) + c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0u;
      }
      if (a < P.Ma) {
#pragma unroll
        for (int k = 0; k < 32; k += 4)
          if (b0 + c0 + k < P.Nb)
            *reinterpret_cast<float4*>(dst + c0 + k) = make_float4(__uint_as_float(v[k]), __uint_as_float(v[k + 1]),
                                                                   __uint_as_float(v[k + 2]), __uint_as_float(v[k + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BN>(tmem);
}

static int grid_for(int M, int N, int BN) {
  const long long tiles = (long long)((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const long long g = tiles < sm_count() ? tiles : sm_count();
  return (int)(g < 1 ? 1 : g);
}
static int pick_bn(int N) { return (N % 256 == 0 || (N > 128 && N <= 256)) ? 256 : 128; }
static int check_shape(int M, int N, int K) {
  SLAK_REQUIRE(M > 0 && N > 0 && K > 0, SLAK_ERR_BAD_ARG, "non-positive GEMM size");
  SLAK_REQUIRE(N % 8 == 0 && K % 8 == 0, SLAK_ERR_UNSUPPORTED, "N=%d, K=%d must be multiples of 8 (16-byte rows for the tensor maps)", N, K);
  SLAK_REQUIRE(N <= 3072, SLAK_ERR_UNSUPPORTED, "N=%d too wide (max 3072)", N);
  return SLAK_OK;
}

template <int BN, int EPI>
static int launch_nt(const CUtensorMap& am, const CUtensorMap& bm, const CUtensorMap& o0, const CUtensorMap& o1,
                     const Params& P, int grid, cudaStream_t st) {
  auto kern = mlp_gemm_nt_kernel<BN, EPI>;
  SLAK_SET_MAX_SMEM(kern, Cfg<BN>::kSmem);
  kern<<<grid, kThreads, Cfg<BN>::kSmem, st>>>(am, bm, o0, o1, P);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}
}  // namespace mlp

int mlp_parts(int M, int N) { return mlp::grid_for(M, N, mlp::pick_bn(N)); }

// epi: 0 FC1 (out0 = H or NULL, out1 = A), 1 BIAS (out0), 2 DGELU (out0 = dH, aux_h = H, colpart), 3 PLAIN (out0)
int mlp_gemm_nt(int epi, const void* a, const void* b, const float* bias, const void* aux_h, void* out0, void* out1,
                float* colpart, int M, int N, int K, cudaStream_t st) {
  using namespace mlp;
  int rc = check_shape(M, N, K);
  if (rc) return rc;
  CUtensorMap am, bm, o0, o1;
  if ((rc = make_plane_map(&am, a, 1, 1, M, K))) return rc;
  if ((rc = make_plane_map(&bm, b, 1, 1, N, K))) return rc;
  if ((rc = make_plane_map(&o0, out0 ? out0 : out1, 1, 1, M, N))) return rc;
  if ((rc = make_plane_map(&o1, out1 ? out1 : out0, 1, 1, M, N))) return rc;
  Params P{};
  P.bias = bias; P.h = (const __nv_bfloat16*)aux_h; P.colpart = colpart; P.M = M; P.N = N; P.K = K;
  P.write_h = out0 != nullptr;
  const int BN = pick_bn(N);
  const int grid = grid_for(M, N, BN);
#define SLAK_NT(E)                                                                         \
  return BN == 256 ? launch_nt<256, E>(am, bm, o0, o1, P, grid, st) : launch_nt<128, E>(am, bm, o0, o1, P, grid, st)
  switch (epi) {
    case EPI_FC1: SLAK_NT(EPI_FC1);
    case EPI_BIAS: SLAK_NT(EPI_BIAS);
    case EPI_DGELU: SLAK_NT(EPI_DGELU);
    default: SLAK_NT(EPI_PLAIN);
  }
#undef SLAK_NT
}

// split plan of the weight-gradient GEMM: (tiles, splits, span)
static void wg_plan(int M, int Ma, int Nb, int* BN, int* tiles_a, int* tiles_b, int* splits, int* span) {
  *BN = mlp::pick_bn(Nb);
  *tiles_a = (Ma + 127) / 128;
  *tiles_b = (Nb + *BN - 1) / *BN;
  const int tiles = *tiles_a * *tiles_b;
  int s = (sm_count() + tiles - 1) / tiles;               // about one wave of CTAs
  const int kblocks = (M + 63) / 64;
  if (s > kblocks) s = kblocks;
  if (s < 1) s = 1;
  int sp = ((kblocks + s - 1) / s) * 64;
  *span = sp;
  *splits = (M + sp - 1) / sp;
}
int mlp_wgrad_splits(int M, int Ma, int Nb) {
  int BN, ta, tb, s, sp;
  wg_plan(M, Ma, Nb, &BN, &ta, &tb, &s, &sp);
  return s;
}
// part[splits][Ma][Nb] = per-split partial sums of P^T Q; fold with slak_colsum_f32(part, splits, Ma * Nb, out)
int mlp_gemm_tn_splitk(const void* p, const void* q, float* part, int M, int Ma, int Nb, cudaStream_t st) {
  using namespace mlp;
  SLAK_REQUIRE(M > 0 && Ma > 0 && Nb > 0, SLAK_ERR_BAD_ARG, "non-positive GEMM size");
  SLAK_REQUIRE(Ma % 8 == 0 && Nb % 8 == 0, SLAK_ERR_UNSUPPORTED, "Ma=%d, Nb=%d must be multiples of 8", Ma, Nb);
  int BN, ta, tb, splits, span;
  wg_plan(M, Ma, Nb, &BN, &ta, &tb, &splits, &span);
  CUtensorMap pm, qm;
  int rc;
  if ((rc = make_plane_map(&pm, p, 1, 1, M, Ma))) return rc;
  if ((rc = make_plane_map(&qm, q, 1, 1, M, Nb))) return rc;
  WgParams P{};
  P.part = part; P.M = M; P.Ma = Ma; P.Nb = Nb; P.span = span; P.tiles_b = tb;
  dim3 grid(ta * tb, splits);
  if (BN == 256) {
    auto kern = mlp_gemm_tn_splitk_kernel<256>;
    SLAK_SET_MAX_SMEM(kern, WgCfg<256>::kSmem);
    kern<<<grid, kWgThreads, WgCfg<256>::kSmem, st>>>(pm, qm, P);
  } else {
    auto kern = mlp_gemm_tn_splitk_kernel<128>;
    SLAK_SET_MAX_SMEM(kern, WgCfg<128>::kSmem);
    kern<<<grid, kWgThreads, WgCfg<128>::kSmem, st>>>(pm, qm, P);
  }
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

}  // namespace tc
}  // namespace slak
