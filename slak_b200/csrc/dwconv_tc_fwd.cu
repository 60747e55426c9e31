// Fused three-branch large-kernel depthwise forward on the 5th-gen tensor cores (bf16 in,
// fp32 accumulate in TMEM, bf16 out):
//     y1 = dwconv_{KL x 5}(x)   y2 = dwconv_{5 x KL}(x)   y3 = dwconv_{5 x 5}(x)
// (the three convolutions of ReparamLargeKernelConv.forward, models/SLaK.py:89-100), x read ONCE.
//
// Formulation (banded-Toeplitz GEMMs, per channel):
//   long axis of a branch  -> contraction (K) against a Toeplitz matrix of the taps, built in
//                             shared memory once per CTA (B operand, K-major, SWIZZLE_128B)
//   short axis (5 taps)    -> five accumulating MMAs whose A operand (the image plane) starts
//                             (t-2) ROWS later: a row shift is a +128-byte descriptor offset
//   y2|y3 [(plane,p), q]  += X [(plane,p+r-2), w]   * [T2_r ; T3_r][q, w]    M=128 N=128 K=64
//   y1^T  [(plane,q), p]  += X^T[(plane,q+s-2), h]  * T1_s[p, h]             M=128 N=64  K=64
//   M stacks two planes of the same channel; each 64x64 plane tile is zero beyond 56 rows, so
//   the zero rows double as the "same" padding between the two stacked planes.
//   X comes from one TMA load per plane (OOB zero fill = padding); X^T is made in shared memory
//   with ldmatrix.trans/stmatrix by two warps.
//
// Warp roles (256 threads): w0 TMA producer | w1 MMA issuer | w2-3 transposers (w2 owns TMEM
// alloc) | w4-7 epilogue (TMEM -> registers -> bf16 -> global; y1 is transposed back through smem).
#include "common.cuh"
#include "tc_common.cuh"

// How a SWIZZLE_128B K-major operand that starts on a 128-byte row which is not 1024-byte aligned
// must be described (measured with tools/umma_probe.cu): base_offset field value.
#ifndef SLAK_TC_BASE_OFF
#define SLAK_TC_BASE_OFF(addr) 0u
#endif

namespace slak {
namespace tc {

constexpr int kStages = 3;                       // X (natural) slots in flight
constexpr int kTStages = 1;                      // X^T slots (refilled in the shadow of the b2|b3 MMAs)
constexpr int kAccBufs = 2;
constexpr int kPlaneBytes = 64 * 128;            // one 64x64 bf16 tile
constexpr int kUnitBytes = 2 * kPlaneBytes;      // two stacked planes
constexpr int kPad = 1024;                       // zero rows before/after a unit tile
constexpr int kToep1Bytes = 5 * 64 * 128;        // 40 KB
constexpr int kToep23Bytes = 5 * 128 * 128;      // 80 KB
constexpr int kXSlot = kPad + kUnitBytes + kPad; // 18 KB
constexpr int kOffToep1 = 0;
constexpr int kOffToep23 = kOffToep1 + kToep1Bytes;
constexpr int kOffXN = kOffToep23 + kToep23Bytes;
constexpr int kOffXT = kOffXN + kStages * kXSlot;
constexpr int kOffY1 = kOffXT + kTStages * kXSlot;         // 16 KB staging for the y1 transpose
constexpr int kOffBar = kOffY1 + kUnitBytes;
constexpr int kSmemBytes = kOffBar + 256 + 1024;           // + alignment slack
constexpr int kTmemCols = 512;
constexpr int kAccCols = 192;                               // D1T: 64 cols, D23: 128 cols
constexpr int kNumTransposerWarps = 2;

struct FwdParams {
  const float* w1; const float* w2; const float* w3;       // fp32 taps [C,KL,5] [C,5,KL] [C,5,5]
  __nv_bfloat16* y1; __nv_bfloat16* y2; __nv_bfloat16* y3;
  int N, C, H, W, KL;
  int splits;            // CTAs per channel
  int pairs_per_c;       // ceil(N/2)
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(256, 1)
lk3_fwd_tc_kernel(const __grid_constant__ CUtensorMap xmap, FwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte aligned base (SWIZZLE_128B atoms)
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x / P.splits;
  const int split = blockIdx.x % P.splits;
  const int u_begin = (int)(((long long)P.pairs_per_c * split) / P.splits);
  const int u_end = (int)(((long long)P.pairs_per_c * (split + 1)) / P.splits);
  const int KL = P.KL, pad = KL / 2, H = P.H, W = P.W;

  // barriers
  constexpr int B_XN_FULL = 0, B_XN_EMPTY = kStages, B_XT_FULL = 2 * kStages, B_XT_EMPTY = B_XT_FULL + kTStages,
                B_ACC_FULL = B_XT_EMPTY + kTStages, B_ACC_EMPTY = B_ACC_FULL + kAccBufs;
  const uint32_t bar0 = base + kOffBar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kOffBar + 128);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(BAR(B_XN_FULL + s), 1);                          // TMA expect_tx arrive
      mbar_init(BAR(B_XN_EMPTY + s), 1 + kNumTransposerWarps);   // MMA commit + transposers done reading
    }
    for (int s = 0; s < kTStages; ++s) {
      mbar_init(BAR(B_XT_FULL + s), kNumTransposerWarps);        // transposers wrote X^T
      mbar_init(BAR(B_XT_EMPTY + s), 1);                         // MMA commit
    }
    for (int a = 0; a < kAccBufs; ++a) {
      mbar_init(BAR(B_ACC_FULL + a), 1);                         // MMA commit
      mbar_init(BAR(B_ACC_EMPTY + a), 4);                        // one arrival per epilogue warp
    }
    mbar_fence_init();
    tma_prefetch_desc(&xmap);
  }

  // ---- zero the pads of the X / X^T slots and build the Toeplitz operands (all threads) -------
  {
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int s = 0; s < kStages + kTStages; ++s) {
      uint8_t* slot = sm + kOffXN + s * kXSlot;
      for (int i = tid; i < kPad / 16; i += 256) {
        reinterpret_cast<uint4*>(slot)[i] = z;
        reinterpret_cast<uint4*>(slot + kPad + kUnitBytes)[i] = z;
      }
    }
    // taps of channel c, rounded to bf16, staged as fp32 in the y1 staging area
    float* wst = reinterpret_cast<float*>(sm + kOffY1);
    float* w1s = wst;                 // [KL][5]
    float* w2s = wst + KL * 5;        // [5][KL]
    float* w3s = wst + 2 * KL * 5;    // [5][5]
    for (int i = tid; i < KL * 5; i += 256) {
      w1s[i] = P.w1[(size_t)c * KL * 5 + i];
      w2s[i] = P.w2[(size_t)c * KL * 5 + i];
    }
    if (tid < 25) w3s[tid] = P.w3[(size_t)c * 25 + tid];
    __syncthreads();
    // T1_s[p][h] = w1[h-p+pad][s]
    for (int ch = tid; ch < 5 * 64 * 8; ch += 256) {
      const int s = ch / 512, rem = ch - s * 512, p = rem >> 3, k8 = rem & 7;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int t = (k8 * 8 + j) - p + pad;
        v[j] = (t >= 0 && t < KL) ? w1s[t * 5 + s] : 0.f;
      }
      uint4 q = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      *reinterpret_cast<uint4*>(sm + kOffToep1 + s * (64 * 128) + p * 128 + ((k8 ^ (p & 7)) << 4)) = q;
    }
    // T23_r rows 0..63: T2_r[q][w] = w2[r][w-q+pad] ; rows 64..127: T3_r[q][w] = w3[r][w-q+2]
    for (int ch = tid; ch < 5 * 128 * 8; ch += 256) {
      const int r = ch / 1024, rem = ch - r * 1024, row = rem >> 3, k8 = rem & 7;
      float v[8];
      if (row < 64) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int t = (k8 * 8 + j) - row + pad;
          v[j] = (t >= 0 && t < KL) ? w2s[r * KL + t] : 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int t = (k8 * 8 + j) - (row - 64) + 2;
          v[j] = (t >= 0 && t < 5) ? w3s[r * 5 + t] : 0.f;
        }
      }
      uint4 q = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      *reinterpret_cast<uint4*>(sm + kOffToep23 + r * (128 * 128) + row * 128 + ((k8 ^ (row & 7)) << 4)) = q;
    }
  }
  fence_proxy_async();      // generic-proxy writes above are read by the tensor core (async proxy)
  if (warp == 2) tmem_alloc<kTmemCols>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int n_units = u_end - u_begin;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      for (int i = 0; i < n_units; ++i) {
        const int st = i % kStages, ph = (i / kStages) & 1;
        mbar_wait(BAR(B_XN_EMPTY + st), ph ^ 1);
        const int n0 = 2 * (u_begin + i);
        const uint32_t dst = base + kOffXN + st * kXSlot + kPad;
        mbar_expect_tx(BAR(B_XN_FULL + st), kUnitBytes);
        const int na = n0, nb = min(n0 + 1, P.N - 1);
        tma_load_3d(dst, &xmap, BAR(B_XN_FULL + st), 0, 0, na * P.C + c);
        tma_load_3d(dst + kPlaneBytes, &xmap, BAR(B_XN_FULL + st), 0, 0, nb * P.C + c);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one()) {
      constexpr uint32_t idesc23 = umma_idesc_bf16(128, 128);
      constexpr uint32_t idesc1 = umma_idesc_bf16(128, 64);
      for (int i = 0; i < n_units; ++i) {
        const int st = i % kStages, ph = (i / kStages) & 1;
        const int ab = i % kAccBufs, aph = (i / kAccBufs) & 1;
        const int ts = i % kTStages, tph = (i / kTStages) & 1;
        mbar_wait(BAR(B_ACC_EMPTY + ab), aph ^ 1);   // epilogue drained this accumulator buffer
        mbar_wait(BAR(B_XN_FULL + st), ph);          // X landed
        tc_fence_after();
        const uint32_t xn = base + kOffXN + st * kXSlot + kPad;
        const uint32_t xt = base + kOffXT + ts * kXSlot + kPad;
        const uint32_t d1 = tmem + ab * kAccCols;
        const uint32_t d23 = d1 + 64;
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t a = xn + (r - 2) * 128 + k * 32;
            const uint32_t b = base + kOffToep23 + r * (128 * 128) + k * 32;
            umma_bf16(d23, umma_desc_k_sw128(a, SLAK_TC_BASE_OFF(a)), umma_desc_k_sw128(b, 0), idesc23, (r | k) != 0);
          }
        umma_commit(BAR(B_XN_EMPTY + st));           // X slot free (with the transposers' arrivals)
        mbar_wait(BAR(B_XT_FULL + ts), tph);         // X^T written
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < 5; ++s)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t a = xt + (s - 2) * 128 + k * 32;
            const uint32_t b = base + kOffToep1 + s * (64 * 128) + k * 32;
            umma_bf16(d1, umma_desc_k_sw128(a, SLAK_TC_BASE_OFF(a)), umma_desc_k_sw128(b, 0), idesc1, (s | k) != 0);
          }
        umma_commit(BAR(B_XT_EMPTY + ts));           // X^T slot free
        umma_commit(BAR(B_ACC_FULL + ab));           // accumulators ready
      }
    }
  } else if (warp < 4) {
    // ================= transposers: X (natural) -> X^T, 8x8 blocks =================
    const int tw = warp - 2;
    const int m = lane >> 3, kk = lane & 7;     // matrix id within the x4, row within the 8x8 block
    for (int i = 0; i < n_units; ++i) {
      const int st = i % kStages, ph = (i / kStages) & 1;
      const int ts = i % kTStages, tph = (i / kTStages) & 1;
      mbar_wait(BAR(B_XN_FULL + st), ph);       // X landed
      mbar_wait(BAR(B_XT_EMPTY + ts), tph ^ 1); // previous X^T of this slot consumed
      const uint32_t xn = base + kOffXN + st * kXSlot + kPad;
      const uint32_t xt = base + kOffXT + ts * kXSlot + kPad;
      for (int it = tw; it < 32; it += kNumTransposerWarps) {
        const int h = it >> 4, bi = (it >> 1) & 7, g = it & 1;
        const int bj = 4 * g + m;
        const uint32_t src = xn + (64 * h + 8 * bi + kk) * 128 + ((bj ^ kk) << 4);
        const uint32_t dst = xt + (64 * h + 8 * bj + kk) * 128 + ((bi ^ kk) << 4);
        uint32_t r0, r1, r2, r3;
        ldmatrix_x4_trans(src, r0, r1, r2, r3);
        stmatrix_x4(dst, r0, r1, r2, r3);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(BAR(B_XT_FULL + ts));       // X^T ready
        mbar_arrive(BAR(B_XN_EMPTY + st));      // done reading X
      }
    }
  } else {
    // ================= epilogue =================
    const int e = warp - 4;
    const int L = e * 32 + lane;                // TMEM lane = (plane half, row)
    const int half = L >> 6, row = L & 63;
    const size_t plane_elems = (size_t)H * W;
    uint8_t* y1s = sm + kOffY1;
    const int wchunks = W >> 3;
    for (int i = 0; i < n_units; ++i) {
      const int ab = i % kAccBufs, aph = (i / kAccBufs) & 1;
      const int n = 2 * (u_begin + i) + half;
      const bool plane_ok = n < P.N;
      const size_t pbase = ((size_t)(plane_ok ? n : 0) * P.C + c) * plane_elems;
      mbar_wait(BAR(B_ACC_FULL + ab), aph);
      tc_fence_after();
      const uint32_t t0 = tmem + ((uint32_t)(e * 32) << 16) + ab * kAccCols;
      uint32_t v[64];
      // ---- y2 (cols 64..127 of the buffer) and y3 (cols 128..191): natural orientation ----
#pragma unroll
      for (int br = 0; br < 2; ++br) {
        tmem_ld32(t0 + 64 + br * 64, v);
        tmem_ld32(t0 + 64 + br * 64 + 32, v + 32);
        tmem_ld_wait();
        if (plane_ok && row < H) {
          __nv_bfloat16* yo = (br == 0 ? P.y2 : P.y3) + pbase + (size_t)row * W;
#pragma unroll
          for (int ck = 0; ck < 8; ++ck) {          // static register indices: no local-memory spill
            if (ck < wchunks) {
              uint4 o;
              o.x = pack_bf16(__uint_as_float(v[8 * ck + 0]), __uint_as_float(v[8 * ck + 1]));
              o.y = pack_bf16(__uint_as_float(v[8 * ck + 2]), __uint_as_float(v[8 * ck + 3]));
              o.z = pack_bf16(__uint_as_float(v[8 * ck + 4]), __uint_as_float(v[8 * ck + 5]));
              o.w = pack_bf16(__uint_as_float(v[8 * ck + 6]), __uint_as_float(v[8 * ck + 7]));
              *reinterpret_cast<uint4*>(yo + 8 * ck) = o;
            }
          }
        }
      }
      // ---- y1^T (cols 0..63): this thread holds column `row`(=q) for p = 0..63 -> staging[p][q] ----
      tmem_ld32(t0, v);
      tmem_ld32(t0 + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + ab));  // accumulators drained
#pragma unroll
      for (int p = 0; p < 64; ++p) {
        const uint32_t off = (uint32_t)(half * 64 + p) * 128 + ((((uint32_t)row >> 3) ^ (p & 7)) << 4) + (row & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(y1s + off) = __float2bfloat16_rn(__uint_as_float(v[p]));
      }
      named_bar_sync(1, 128);
      if (plane_ok && row < H) {
        __nv_bfloat16* yo = P.y1 + pbase + (size_t)row * W;
        for (int ck = 0; ck < wchunks; ++ck) {
          const uint32_t off = (uint32_t)(half * 64 + row) * 128 + ((ck ^ (row & 7)) << 4);
          *reinterpret_cast<uint4*>(yo + 8 * ck) = *reinterpret_cast<const uint4*>(y1s + off);
        }
      }
      named_bar_sync(1, 128);                     // staging free for the next unit
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<kTmemCols>(tmem);
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

bool lk3_tc_supported(int N, int C, int H, int W, int KL) {
  (void)N; (void)C;
  return H >= 8 && W >= 8 && H <= 62 && W <= 62 && (W % 8) == 0 && (KL & 1) && KL >= 5 && KL * 5 * 2 + 25 <= 4000;
}

int lk3_fwd_tc(const void* x, const float* w1, const float* w2, const float* w3, void* y1, void* y2, void* y3,
               int N, int C, int H, int W, int KL, cudaStream_t st) {
  SLAK_REQUIRE(lk3_tc_supported(N, C, H, W, KL), SLAK_ERR_UNSUPPORTED,
               "tensor-core path needs 8 <= H,W <= 62, W %% 8 == 0 (got %dx%d)", H, W);
  SLAK_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, SLAK_ERR_BAD_ARG, "x must be 16-byte aligned");
  CUtensorMap map;
  int rc = make_plane_map(&map, x, N, C, H, W);
  if (rc) return rc;
  FwdParams P;
  P.w1 = w1; P.w2 = w2; P.w3 = w3;
  P.y1 = (__nv_bfloat16*)y1; P.y2 = (__nv_bfloat16*)y2; P.y3 = (__nv_bfloat16*)y3;
  P.N = N; P.C = C; P.H = H; P.W = W; P.KL = KL;
  P.pairs_per_c = (N + 1) / 2;
  // CTAs per channel: fill the SMs with whole waves, at least ~8 plane pairs per CTA
  const int sms = sm_count();
  int best = 1; double best_eff = 0.0;
  const int max_s = P.pairs_per_c >= 8 ? P.pairs_per_c / 8 : 1;
  for (int s = 1; s <= max_s && s <= 64; ++s) {
    const long long ctas = (long long)C * s;
    const long long waves = (ctas + sms - 1) / sms;
    const int per = (P.pairs_per_c + s - 1) / s;
    const double eff = (double)C * P.pairs_per_c / ((double)waves * sms * per) * (per / (per + 1.5));
    if (eff > best_eff) { best_eff = eff; best = s; }
  }
  P.splits = best;
  SLAK_CUDA_TRY(cudaFuncSetAttribute(lk3_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  lk3_fwd_tc_kernel<<<C * P.splits, 256, kSmemBytes, st>>>(map, P);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

}  // namespace tc
}  // namespace slak
