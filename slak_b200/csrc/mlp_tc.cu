// ROUND-2 DRAFT -- compiled into the library but NOT on the default path (slak_b200/block.py uses it only with
// SLAK_FUSED_MLP=1) and NOT yet run on hardware: written after the round's GPU budget was spent.  The building
// blocks (K-major SWIZZLE_128B descriptors, 64x64 TMA boxes with OOB fill / clipping, TMEM double buffering,
// staging tile + TMA tile store) are the ones the depthwise kernels use and have been verified there.
//
// Pointwise MLP of a SLaK Block on the tensor cores (models/SLaK.py:157-160: pwconv1 -> GELU -> pwconv2), the two
// GEMMs whose epilogues replace separate elementwise passes:
//
//   fc1_gelu_fwd    H = Xn W1^T + b1   (bf16, what nn.Linear returns under autocast)
//                   A = gelu(H)        (exact erf GELU of the ROUNDED H, as F.gelu on the bf16 tensor)
//                   one pass: Xn read once, H and A written once (the separate GELU kernel re-reads H)
//   fc2_dgelu_bwd   dH = (dH2 W2) * gelu'(H),  per-CTA partial column sums of dH (bias gradient of pwconv1)
//                   dA = dH2 W2 never goes to HBM
//
// Both are TN GEMMs D[M,N] = A[M,K] B[N,K]^T with K = C (96..768) or 4C: at these shapes they are HBM-bound
// (stage 1: 77 MB in, 616 MB out for fc1), so the structure is a plain persistent warp-specialised GEMM:
// 128 x 128 output tiles, K in 64-wide SWIZZLE_128B blocks through a 4-stage TMA ring, fp32 accumulators
// double-buffered in TMEM (2 x 128 columns), two epilogue warpgroups that each own 64 columns of the tile.
// Warp roles (384 threads): w0 TMA producer | w1 MMA issuer | w2 TMEM allocator | w4-7, w8-11 epilogue.
#include "common.cuh"
#include "tc_common.cuh"
#include <string.h>

namespace slak {
namespace tc {

int make_plane_map(CUtensorMap* map, const void* x, int N, int C, int H, int W);  // dwconv_tc_fwd.cu: (W, H, N*C), box 64x64x1

namespace mlp {
constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kStages = 4;
constexpr int kHalf = 64 * 128;                 // one 64-row x 128-byte box (8 KB)
constexpr int kTile = 2 * kHalf;                // 128 rows (16 KB)
constexpr int kStage = 2 * kTile;               // A tile + B tile
constexpr int kOffStg = kStages * kStage;       // staging: [group][0 = first output, 1 = second output], 16 KB each
constexpr int kOffCol = kOffStg + 4 * kTile;    // fc2: per-CTA column accumulators [N] fp32 (<= 16 KB) | scratch [8][64]
constexpr int kColBytes = 16384 + 8 * 64 * 4;
constexpr int kOffBar = kOffCol + kColBytes;
constexpr int kSmem = kOffBar + 256 + 1024;
constexpr int kThreads = 384;
static_assert(kSmem <= 232448, "shared memory budget");

struct Params {
  const float* bias;                // fc1: [N] fp32 (rounded to bf16 on use, as b1.to(bf16) in the module path)
  const __nv_bfloat16* h;           // fc2: [M][N] bf16, the saved pre-activation
  float* colpart;                   // fc2: [grid][N] per-CTA partial column sums of dH
  int M, N, K;
};

// d/dx [x Phi(x)] and x Phi(x) with Phi through erf's rational approximation (A&S 7.1.26, |err| <= 1.5e-7)
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf) {
  const float e = __expf(-0.5f * x * x);
  const float t = __fdividef(1.f, fmaf(0.3275911f * 0.70710678118654752f, fabsf(x), 1.f));
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  cdf = 0.5f + copysignf(0.5f * fmaf(-poly, e, 1.f), x);
  pdf = 0.39894228040143268f * e;
}

template <bool BWD>
__global__ void __launch_bounds__(kThreads, 1)
mlp_gemm_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                const __grid_constant__ CUtensorMap o0map, const __grid_constant__ CUtensorMap o1map, Params P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int m_tiles = (P.M + BM - 1) / BM, n_tiles = P.N / BN;
  const int tiles = m_tiles * n_tiles;
  const int KB = (P.K + BK - 1) / BK;

  constexpr int B_FULL = 0, B_EMPTY = kStages, B_ACC_FULL = 2 * kStages, B_ACC_EMPTY = B_ACC_FULL + 2;
  const uint32_t bar0 = base + kOffBar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kOffBar + 192);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(BAR(B_FULL + s), 1); mbar_init(BAR(B_EMPTY + s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(BAR(B_ACC_FULL + a), 1); mbar_init(BAR(B_ACC_EMPTY + a), 8); }
    mbar_fence_init();
    tma_prefetch_desc(&amap); tma_prefetch_desc(&bmap); tma_prefetch_desc(&o0map);
    if (!BWD) tma_prefetch_desc(&o1map);
  }
  if (BWD) {   // column accumulators start at zero
    float* col = reinterpret_cast<float*>(sm + kOffCol);
    for (int i = tid; i < P.N; i += kThreads) col[i] = 0.f;
  }
  if (warp == 2) tmem_alloc<256>(smem_u32(tmem_slot));
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
          const uint32_t sa = base + s * kStage, sb = sa + kTile;
          mbar_expect_tx(BAR(B_FULL + s), kStage);
          tma_load_3d(sa, &amap, BAR(B_FULL + s), kb * BK, m0, 0);
          tma_load_3d(sa + kHalf, &amap, BAR(B_FULL + s), kb * BK, m0 + 64, 0);
          tma_load_3d(sb, &bmap, BAR(B_FULL + s), kb * BK, n0, 0);
          tma_load_3d(sb + kHalf, &bmap, BAR(B_FULL + s), kb * BK, n0 + 64, 0);
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
          const uint32_t sa = base + s * kStage, sb = sa + kTile;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk)
            umma_bf16(acc, umma_desc_k_sw128(sa + kk * 32, 0), umma_desc_k_sw128(sb + kk * 32, 0), idesc, (kb | kk) != 0);
          umma_commit(BAR(B_EMPTY + s));
        }
        umma_commit(BAR(B_ACC_FULL + ab));
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: group g owns columns [64 g, 64 g + 64) of every tile =================
    const int g = (warp - 4) >> 2, e = (warp - 4) & 3;
    const int L = e * 32 + lane;                          // row of the tile = TMEM lane
    const int nb = 1 + g;                                 // named barrier of the group
    uint8_t* stg0 = sm + kOffStg + (2 * g) * kTile;       // first output (H / dH), [128 rows][128 B] swizzled
    uint8_t* stg1 = stg0 + kTile;                         // second output (A), forward only
    const uint32_t stg0_s = base + kOffStg + (2 * g) * kTile, stg1_s = stg0_s + kTile;
    float* col = reinterpret_cast<float*>(sm + kOffCol);
    float* scratch = col + 4096;                          // [8 warps][64]
    int it = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN + 64 * g;
      const int ab = it & 1, aph = (it >> 1) & 1;
      const int m = m0 + L;
      uint4 hraw[8];                                      // fc2: the saved pre-activation row, requested before the wait
      if (BWD) {
        const uint4* hp = reinterpret_cast<const uint4*>(P.h + (size_t)(m < P.M ? m : 0) * P.N + n0);
#pragma unroll
        for (int j = 0; j < 8; ++j) hraw[j] = __ldg(hp + j);
      }
      mbar_wait(BAR(B_ACC_FULL + ab), aph);
      tc_fence_after();
      uint32_t v[64];
      const uint32_t ta = tmem + ((uint32_t)(e * 32) << 16) + ab * BN + 64 * g;
      tmem_ld32(ta, v); tmem_ld32(ta + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + ab));  // accumulators drained (8 warps)
      // the previous tile's stores must have finished reading this group's staging tiles
      if (e == 0 && lane == 0) bulk_wait_group_read<0>();
      named_bar_sync(nb, 128);
      float cs[64];                                       // fc2: this row's contribution to the column sums
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float o[8], a[8];
        if (!BWD) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(P.bias + n0 + 8 * j));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(P.bias + n0 + 8 * j + 4));
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float hv = __uint_as_float(v[8 * j + k]) + __bfloat162float(__float2bfloat16_rn(bb[k]));
            o[k] = hv;
            const float hr = __bfloat162float(__float2bfloat16_rn(hv));     // GELU of the value that is stored
            float cdf, pdf;
            gelu_parts(hr, cdf, pdf);
            a[k] = hr * cdf;
          }
        } else {
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
        }
        const uint32_t off = (uint32_t)L * 128 + ((j ^ (L & 7)) << 4);
        *reinterpret_cast<uint4*>(stg0 + off) = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]),
                                                           pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        if (!BWD)
          *reinterpret_cast<uint4*>(stg1 + off) = make_uint4(pack_bf16(a[0], a[1]), pack_bf16(a[2], a[3]),
                                                             pack_bf16(a[4], a[5]), pack_bf16(a[6], a[7]));
      }
      fence_proxy_async();
      if (BWD) {
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
      if (BWD && e == 0) {      // one warp folds the group's four row blocks into the CTA's column accumulators
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int cidx = 32 * hh + lane;
          float s = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) s += scratch[(4 * g + w) * 64 + cidx];
          col[n0 + cidx] += s;
        }
      }
      if (e == 0 && lane == 0) {
        tma_store_3d(&o0map, stg0_s, n0, m0, 0);
        tma_store_3d(&o0map, stg0_s + kHalf, n0, m0 + 64, 0);
        if (!BWD) {
          tma_store_3d(&o1map, stg1_s, n0, m0, 0);
          tma_store_3d(&o1map, stg1_s + kHalf, n0, m0 + 64, 0);
        }
        bulk_commit_group();
      }
      if (BWD) named_bar_sync(nb, 128);                   // scratch is rewritten by the next tile
    }
    if (e == 0 && lane == 0) bulk_wait_group_read<0>();   // shared memory must outlive the last tile store
  }

  tc_fence_before();
  __syncthreads();
  if (BWD) {
    const float* col = reinterpret_cast<const float*>(sm + kOffCol);
    for (int i = tid; i < P.N; i += kThreads) P.colpart[(size_t)blockIdx.x * P.N + i] = col[i];
  }
  if (warp == 2) tmem_dealloc<256>(tmem);
}

static int grid_for(int M, int N) {
  const long long tiles = (long long)((M + BM - 1) / BM) * (N / BN);
  const long long g = tiles < sm_count() ? tiles : sm_count();
  return (int)(g < 1 ? 1 : g);
}
static int check_shape(int M, int N, int K) {
  SLAK_REQUIRE(M > 0 && N > 0 && K > 0, SLAK_ERR_BAD_ARG, "non-positive GEMM size");
  SLAK_REQUIRE(N % BN == 0, SLAK_ERR_UNSUPPORTED, "N=%d must be a multiple of %d", N, BN);
  SLAK_REQUIRE(K % 8 == 0, SLAK_ERR_UNSUPPORTED, "K=%d must be a multiple of 8 (16-byte rows for the tensor map)", K);
  SLAK_REQUIRE(N <= 4096, SLAK_ERR_UNSUPPORTED, "N=%d too wide for the column accumulators", N);
  return SLAK_OK;
}
}  // namespace mlp

int mlp_parts(int M, int N) { return mlp::grid_for(M, N); }

// H[M,N] = X[M,K] W[N,K]^T + bias, A = gelu(H); X, W, H, A bf16 row-major
int mlp_fc1_gelu_fwd(const void* x, const void* w, const float* bias, void* h, void* a, int M, int N, int K, cudaStream_t st) {
  using namespace mlp;
  int rc = check_shape(M, N, K);
  if (rc) return rc;
  CUtensorMap am, bm, hm, gm;
  if ((rc = make_plane_map(&am, x, 1, 1, M, K))) return rc;
  if ((rc = make_plane_map(&bm, w, 1, 1, N, K))) return rc;
  if ((rc = make_plane_map(&hm, h, 1, 1, M, N))) return rc;
  if ((rc = make_plane_map(&gm, a, 1, 1, M, N))) return rc;
  Params P{};
  P.bias = bias; P.M = M; P.N = N; P.K = K;
  auto kern = mlp_gemm_kernel<false>;
  SLAK_SET_MAX_SMEM(kern, kSmem);
  kern<<<grid_for(M, N), kThreads, kSmem, st>>>(am, bm, hm, gm, P);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

// dH[M,N] = (G[M,K] Wt[N,K]^T) * gelu'(H[M,N]); colpart[mlp_parts(M,N)][N] = per-CTA column sums of dH
int mlp_fc2_dgelu_bwd(const void* g, const void* wt, const void* h, void* dh, float* colpart, int M, int N, int K,
                      cudaStream_t st) {
  using namespace mlp;
  int rc = check_shape(M, N, K);
  if (rc) return rc;
  CUtensorMap am, bm, om;
  if ((rc = make_plane_map(&am, g, 1, 1, M, K))) return rc;
  if ((rc = make_plane_map(&bm, wt, 1, 1, N, K))) return rc;
  if ((rc = make_plane_map(&om, dh, 1, 1, M, N))) return rc;
  Params P{};
  P.h = (const __nv_bfloat16*)h; P.colpart = colpart; P.M = M; P.N = N; P.K = K;
  auto kern = mlp_gemm_kernel<true>;
  SLAK_SET_MAX_SMEM(kern, kSmem);
  kern<<<grid_for(M, N), kThreads, kSmem, st>>>(am, bm, om, om, P);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

}  // namespace tc
}  // namespace slak
