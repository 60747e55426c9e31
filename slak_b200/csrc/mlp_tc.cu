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

template <int BN, int EPI> struct Cfg {
  // DGELU brings the saved pre-activation H in by TMA into a third staging slab per group (it is overwritten in place by
  // dH), and pays for that with one operand stage less: its main loop (K = C <= 768) is short next to its epilogue
  static constexpr bool kDgelu = (EPI == EPI_DGELU);
  static constexpr int kStages = (BN == 256 ? 3 : 4) - (kDgelu ? 1 : 0);
  static constexpr int kATile = 2 * kBox;                      // 128 rows x 64 k
  static constexpr int kBTile = (BN / 64) * kBox;
  static constexpr int kStage = kATile + kBTile;
  static constexpr int kSlabs = kDgelu ? 3 : 2;                // staging slabs (128 rows x 128 B) per epilogue group
  static constexpr int kOffStg = kStages * kStage;
  static constexpr int kOffCol = kOffStg + 2 * kSlabs * 2 * kBox;   // DGELU: per-CTA column accumulators [N <= 3072] fp32 + scratch
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

// ---- packed fp32 pairs (FFMA2 / FMUL2 / FADD2: two fp32 lanes per instruction on sm_100a) -------------------------
// The GELU epilogues are bound by instruction issue (each element costs ~25 scalar instructions in the A&S 7.1.26
// form, and its exp / rcp and the fp32 -> bf16 -> fp32 round trips go through the quarter-rate XU pipe); in packed
// form with a polynomial Phi they cost ~10.
// Phi(x) = (1 + erf(x / sqrt 2)) / 2 for a pair: Phi = sat(0.5 + x Q(x^2)), Q the degree-8 near-minimax polynomial of
// (Phi(x) - 0.5) / x on [-4.25, 4.25] (odd Chebyshev fit).  Beyond the interval the leading term (positive) drives
// 0.5 + x Q to +-infinity and the saturation gives exactly 1 / 0.  |error| <= 1.1e-5 over the whole real line in fp32
// Horner form (tools/gelu_poly.py), i.e. <= 1e-5 |x| in gelu(x) = x Phi(x): far below the bf16 rounding of the result.
__device__ __forceinline__ f2 phi2(f2 x, f2 u) {
  f2 q = splat(5.342467094e-11f);
  q = fma2(q, u, splat(-5.157194671e-09f));
  q = fma2(q, u, splat(2.201571192e-07f));
  q = fma2(q, u, splat(-5.536209756e-06f));
  q = fma2(q, u, splat(9.255817713e-05f));
  q = fma2(q, u, splat(-1.103906194e-03f));
  q = fma2(q, u, splat(9.802624583e-03f));
  q = fma2(q, u, splat(-6.632731855e-02f));
  q = fma2(q, u, splat(3.988958895e-01f));
  float x0, x1, q0, q1;
  un2(x, x0, x1);
  un2(q, q0, q1);
  return mk2(fma_sat(x0, q0, 0.5f), fma_sat(x1, q1, 0.5f));
}
// gelu(x) = x Phi(x)
__device__ __forceinline__ f2 gelu2(f2 x) { return mul2(x, phi2(x, mul2(x, x))); }
// gelu'(x) = Phi(x) + x phi(x), phi(x) = exp(-x^2 / 2) / sqrt(2 pi) through ex2.approx (one MUFU per element)
__device__ __forceinline__ f2 dgelu2(f2 x) {
  const f2 u = mul2(x, x);
  const f2 P = phi2(x, u);
  float e0, e1;
  un2(mul2(u, splat(-0.72134752044448170f)), e0, e1);           // -x^2 / 2 * log2(e)
  asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(e0));
  asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(e1));
  return fma2(mul2(x, splat(0.39894228040143268f)), mk2(e0, e1), P);
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
mlp_gemm_nt_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                   const __grid_constant__ CUtensorMap o0map, const __grid_constant__ CUtensorMap o1map, Params P) {
  using C = Cfg<BN, EPI>;
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

  constexpr int B_FULL = 0, B_EMPTY = kStages, B_ACC_FULL = 2 * kStages, B_ACC_EMPTY = B_ACC_FULL + 2, B_H_FULL = B_ACC_EMPTY + 2;
  const uint32_t bar0 = base + C::kOffBar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::kOffBar + 192);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(BAR(B_FULL + s), 1); mbar_init(BAR(B_EMPTY + s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(BAR(B_ACC_FULL + a), 1); mbar_init(BAR(B_ACC_EMPTY + a), 8); }
    for (int hb = 0; hb < 6; ++hb) mbar_init(BAR(B_H_FULL + hb), 1);     // DGELU: H slab landed (group g, slab b: index 3 g + b)
    mbar_fence_init();
    tma_prefetch_desc(&amap); tma_prefetch_desc(&bmap); tma_prefetch_desc(&o0map);
    if (EPI == EPI_FC1 || EPI == EPI_DGELU) tma_prefetch_desc(&o1map);
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
    uint8_t* stg = sm + C::kOffStg + g * C::kSlabs * 2 * kBox;   // kSlabs slabs [128 rows][128 B] swizzled (each = two 64-row boxes)
    const uint32_t stg_s = base + C::kOffStg + g * C::kSlabs * 2 * kBox;
    float* col = reinterpret_cast<float*>(sm + C::kOffCol);
    float* scratch = col + 3072;                          // [8 warps][64]
    float* bias_s = col + g * 64;                         // FC1 / BIAS: the group's rounded bias slab (col[] is DGELU's)
    int it = 0, slab_ctr = 0;
    // DGELU: the saved pre-activation H of a slab arrives by TMA (o1map is H's map) in the staging slab where dH is then
    // formed in place; the elected thread runs a prefetch cursor two slabs ahead of the compute (three slabs rotate)
    int pf_t = blockIdx.x, pf_sl = 0, pf_count = 0;
    auto pf_issue = [&]() {
      while (pf_t < tiles && (pf_t % n_tiles) * BN + 64 * (g * SPG + pf_sl) >= P.N)
        if (++pf_sl == SPG) { pf_sl = 0; pf_t += gridDim.x; }
      if (pf_t >= tiles) return;
      const int b = pf_count % 3;
      const int pm0 = (pf_t / n_tiles) * BM, pn0 = (pf_t % n_tiles) * BN + 64 * (g * SPG + pf_sl);
      const uint32_t bar = BAR(B_H_FULL + 3 * g + b), dst = stg_s + b * 2 * kBox;
      mbar_expect_tx(bar, 2 * kBox);
      tma_load_3d(dst, &o1map, bar, pn0, pm0, 0);
      tma_load_3d(dst + kBox, &o1map, bar, pn0, pm0 + 64, 0);
      ++pf_count;
      if (++pf_sl == SPG) { pf_sl = 0; pf_t += gridDim.x; }
    };
    if constexpr (EPI == EPI_DGELU) {
      if (e == 0 && lane == 0) { pf_issue(); pf_issue(); }
    }
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int m0 = (t / n_tiles) * BM, nt0 = (t % n_tiles) * BN;
      const int ab = it & 1, aph = (it >> 1) & 1;
      const int m = m0 + L;
#pragma unroll 1
      for (int sl = 0; sl < SPG; ++sl) {
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
        // staging: the single-output epilogues alternate between the group's two slabs; FC1 writes H to slab 0 and A to
        // slab 1 and stores them as separate bulk groups.  Either way a slab is rewritten one slab period after its
        // store was issued, and only the store issued before the most recent one has to have finished reading.
        const int buf = (EPI == EPI_FC1) ? 0 : (EPI == EPI_DGELU ? slab_ctr % 3 : (slab_ctr & 1));
        if constexpr (EPI == EPI_DGELU) {
          mbar_wait(BAR(B_H_FULL + 3 * g + buf), (slab_ctr / 3) & 1);       // H of this slab landed in slab `buf`
        } else {
          if (e == 0 && lane == 0) bulk_wait_group_read<1>();
        }
        if constexpr (EPI == EPI_FC1 || EPI == EPI_BIAS) {
          // the slab's 64 bias values, rounded to bf16 once (b.to(bf16) of the module path), for broadcast reads below;
          // the previous slab's readers are past the barrier that closed it
          if (e == 1 || e == 2) {
            const int c = (e - 1) * 32 + lane;
            bias_s[c] = (n0 + c < P.N) ? __bfloat162float(__float2bfloat16_rn(__ldg(P.bias + n0 + c))) : 0.f;
          }
        }
        if constexpr (EPI != EPI_DGELU) named_bar_sync(nb, 128);
        uint8_t* s0 = stg + buf * 2 * kBox;
        uint8_t* s1 = stg + 2 * kBox;
        if constexpr (EPI == EPI_FC1) {
          // pass 1: H = acc + bias, rounded once; its store drains while pass 2 computes the GELU
          uint32_t hb[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + 8 * j);          // bf16-rounded bias, broadcast reads
            const float4 b1 = *reinterpret_cast<const float4*>(bias_s + 8 * j + 4);
            const f2 bb[4] = {mk2(b0.x, b0.y), mk2(b0.z, b0.w), mk2(b1.x, b1.y), mk2(b1.z, b1.w)};
#pragma unroll
            for (int k = 0; k < 4; ++k) hb[4 * j + k] = pack2(add2(mk2u(v[8 * j + 2 * k], v[8 * j + 2 * k + 1]), bb[k]));
            if (P.write_h)
              *reinterpret_cast<uint4*>(s0 + (uint32_t)L * 128 + ((j ^ (L & 7)) << 4)) =
                  make_uint4(hb[4 * j], hb[4 * j + 1], hb[4 * j + 2], hb[4 * j + 3]);
          }
          if (P.write_h) {
            fence_proxy_async();
            named_bar_sync(nb, 128);
            if (e == 0 && lane == 0) {
              tma_store_3d(&o0map, stg_s, n0, m0, 0);
              tma_store_3d(&o0map, stg_s + kBox, n0, m0 + 64, 0);
              bulk_commit_group();
              bulk_wait_group_read<1>();                 // the previous slab's A store (older than the H store just issued)
            }
          } else if (e == 0 && lane == 0) {
            bulk_wait_group_read<0>();
          }
          // pass 2: A = gelu(H) of the value that is stored
          uint32_t ab[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) ab[i] = pack2(gelu2(unpack2(hb[i])));
          named_bar_sync(nb, 128);                       // slab 1 is free (the elected thread's wait above)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(s1 + (uint32_t)L * 128 + ((j ^ (L & 7)) << 4)) =
                make_uint4(ab[4 * j], ab[4 * j + 1], ab[4 * j + 2], ab[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint32_t ob[4];
            if constexpr (EPI == EPI_BIAS) {
              const float4 b0 = *reinterpret_cast<const float4*>(bias_s + 8 * j);
              const float4 b1 = *reinterpret_cast<const float4*>(bias_s + 8 * j + 4);
              const f2 bb[4] = {mk2(b0.x, b0.y), mk2(b0.z, b0.w), mk2(b1.x, b1.y), mk2(b1.z, b1.w)};
#pragma unroll
              for (int k = 0; k < 4; ++k) ob[k] = pack2(add2(mk2u(v[8 * j + 2 * k], v[8 * j + 2 * k + 1]), bb[k]));
            } else if constexpr (EPI == EPI_DGELU) {
              const uint4 hq = *reinterpret_cast<const uint4*>(s0 + (uint32_t)L * 128 + ((j ^ (L & 7)) << 4));   // H, replaced by dH below
              const uint32_t hb[4] = {hq.x, hq.y, hq.z, hq.w};
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ob[k] = pack2(mul2(mk2u(v[8 * j + 2 * k], v[8 * j + 2 * k + 1]), dgelu2(unpack2(hb[k]))));
              if (m >= P.M) { ob[0] = ob[1] = ob[2] = ob[3] = 0u; }          // rows beyond M must not reach the column sums
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) ob[k] = pack_bf16(__uint_as_float(v[8 * j + 2 * k]), __uint_as_float(v[8 * j + 2 * k + 1]));
            }
            *reinterpret_cast<uint4*>(s0 + (uint32_t)L * 128 + ((j ^ (L & 7)) << 4)) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
          }
        }
        if constexpr (EPI == EPI_DGELU) {
          // column sums of the STORED (rounded) dH: warp e wrote rows 32 e .. 32 e + 31 of the slab itself, lane l now adds
          // up columns 2l, 2l + 1 over them (one conflict-free 128-byte row per load instruction)
          __syncwarp();
          f2 acc2 = splat(0.f);
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            const uint32_t row = (uint32_t)(e * 32 + r);
            const uint32_t w = *reinterpret_cast<const uint32_t*>(s0 + row * 128 + ((((uint32_t)lane >> 2) ^ (row & 7)) << 4) + (lane & 3) * 4);
            acc2 = add2(acc2, unpack2(w));
          }
          *reinterpret_cast<f2*>(scratch + (warp - 4) * 64 + 2 * lane) = acc2;
        }
        fence_proxy_async();
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
          if constexpr (EPI == EPI_FC1) {
            tma_store_3d(&o1map, stg_s + 2 * kBox, n0, m0, 0);
            tma_store_3d(&o1map, stg_s + 2 * kBox + kBox, n0, m0 + 64, 0);
          } else {
            const uint32_t a0 = stg_s + buf * 2 * kBox;
            tma_store_3d(&o0map, a0, n0, m0, 0);
            tma_store_3d(&o0map, a0 + kBox, n0, m0 + 64, 0);
          }
          bulk_commit_group();
          if constexpr (EPI == EPI_DGELU) {
            bulk_wait_group_read<1>();      // the store before this one has let go of its slab: H of slab + 2 goes there
            pf_issue();
          }
        }
        ++slab_ctr;                         // counts the slabs actually processed (skipped ones take no staging slab)
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
  SLAK_SET_MAX_SMEM(kern, (Cfg<BN, EPI>::kSmem));
  kern<<<grid, kThreads, Cfg<BN, EPI>::kSmem, st>>>(am, bm, o0, o1, P);
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
  if ((rc = make_plane_map(&o1, epi == EPI_DGELU ? aux_h : (out1 ? out1 : out0), 1, 1, M, N))) return rc;   // DGELU: H's map
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
  int s = sm_count() / tiles;                             // at most one wave of CTAs (a second, partial wave costs a full pass)
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
