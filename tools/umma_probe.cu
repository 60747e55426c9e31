// Development probe (not part of the product): establishes how tcgen05.mma treats a
// SWIZZLE_128B K-major operand whose start address is shifted by whole 128-byte rows
// (not 1024-byte aligned), with base_offset = 0 or = (row & 7).  Prints max |err| per variant.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(1) << 16;                          // LBO (ignored for swizzled K-major)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;  // SBO
  d |= (uint64_t)1 << 46;                            // version = 1 (Blackwell)
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}

// A: rows [0, AROWS) x 64 k, SW128 K-major at smem A0 (1024-aligned); B: 64 rows (n) x 64 k.
constexpr int AROWS = 160;
__global__ void probe(const __nv_bfloat16* __restrict__ Ag, const __nv_bfloat16* __restrict__ Bg,
                      float* __restrict__ Dg, int shift, int use_base_off, int N) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* A0 = smem;                       // AROWS*128
  uint8_t* B0 = smem + 24 * 1024;           // 128*128 max
  uint64_t* bar = (uint64_t*)(smem + 48 * 1024);
  uint32_t* tmem_slot = (uint32_t*)(smem + 48 * 1024 + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < AROWS * 64; i += blockDim.x) {
    int r = i / 64, k = i % 64;
    int off = r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2;
    *(__nv_bfloat16*)(A0 + off) = Ag[i];
  }
  for (int i = tid; i < N * 64; i += blockDim.x) {
    int r = i / 64, k = i % 64;
    int off = r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2;
    *(__nv_bfloat16*)(B0 + off) = Bg[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    // idesc: c=f32, a=b=bf16, K-major, N, M=128
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t a_addr = smem_u32(A0) + shift * 128;
    uint32_t b_addr = smem_u32(B0);
    for (int k = 0; k < 4; ++k) {
      uint64_t da = make_desc(a_addr + k * 32, 1024, use_base_off ? (shift & 7) : 0);
      uint64_t db = make_desc(b_addr + k * 32, 1024, 0);
      uint32_t acc = k > 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)));
  }
  // everyone waits for the commit
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"(0));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // 4 warps read their 32 lanes, N columns, 32 at a time
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 32; ++j) Dg[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128));
}

// ---- MN-major probe: D[m][n] = sum_k A[k][m] * B[k][n]; A/B tiles are [k rows][64 elems] SW128.
// A has M=128 as two 64-wide atoms LBO bytes apart (possibly OVERLAPPING: same tile, shifted rows).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// smem: T0 tile rows [-8, 72) (zero outside [0,64)), T1 likewise, B tile rows [-8,72)
__global__ void probe_mn(const __nv_bfloat16* __restrict__ T0g, const __nv_bfloat16* __restrict__ T1g,
                         const __nv_bfloat16* __restrict__ Bg, float* __restrict__ Dg,
                         int shiftA, int shiftB, int lbo_mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* T0 = smem + 1024;            // row 0 of tile 0 (8 zero rows before at smem+0)
  uint8_t* T1 = smem + 1024 + 10240;    // 80 rows per tile region
  uint8_t* B0 = smem + 1024 + 20480;
  uint64_t* bar = (uint64_t*)(smem + 40 * 1024);
  uint32_t* tmem_slot = (uint32_t*)(smem + 40 * 1024 + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 40 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  __syncthreads();
  for (int i = tid; i < 64 * 64; i += blockDim.x) {
    int r = i / 64, k = i % 64;
    int off = r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2;
    *(__nv_bfloat16*)(T0 + off) = T0g[i];
    *(__nv_bfloat16*)(T1 + off) = T1g[i];
    *(__nv_bfloat16*)(B0 + off) = Bg[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    // a_major = b_major = MN (bits 15,16), N=64, M=128
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
    // lbo_mode 0: atoms = T0, T1 (LBO = 10240); lbo_mode 1: atoms = T0 shifted by shiftA and shiftA+1 rows (LBO = 128)
    uint32_t a_addr = smem_u32(T0) + shiftA * 128;
    uint32_t lbo = lbo_mode == 0 ? 10240 : 128;
    uint32_t b_addr = smem_u32(B0) + shiftB * 128;
    for (int k = 0; k < 4; ++k) {
      uint64_t da = make_desc_mn(a_addr + k * 2048, lbo, 1024);
      uint64_t db = make_desc_mn(b_addr + k * 2048, 0, 1024);
      uint32_t acc = k > 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)));
  }
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"(0));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 32; ++j) Dg[(size_t)tid * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128));
}

static void run_mn() {
  std::vector<__nv_bfloat16> T0(64 * 64), T1(64 * 64), B(64 * 64);
  std::vector<float> T0f(64 * 64), T1f(64 * 64), Bf(64 * 64);
  srand(2);
  auto fill = [](std::vector<__nv_bfloat16>& a, std::vector<float>& f) {
    for (size_t i = 0; i < a.size(); ++i) { float v = (float)((rand() % 17) - 8) / 8.f; a[i] = __float2bfloat16(v); f[i] = __bfloat162float(a[i]); }
  };
  fill(T0, T0f); fill(T1, T1f); fill(B, Bf);
  __nv_bfloat16 *d0, *d1, *dB; float* dD;
  CK(cudaMalloc(&d0, 8192)); CK(cudaMalloc(&d1, 8192)); CK(cudaMalloc(&dB, 8192)); CK(cudaMalloc(&dD, 128 * 64 * 4));
  CK(cudaMemcpy(d0, T0.data(), 8192, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d1, T1.data(), 8192, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), 8192, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, 41 * 1024));
  std::vector<float> D(128 * 64);
  auto at = [](const std::vector<float>& t, int r, int c) { return (r >= 0 && r < 64) ? t[r * 64 + c] : 0.f; };
  for (int mode = 0; mode < 2; ++mode)
    for (int sa : {0, -2, 1, 3})
      for (int sb : {0, -1, 2}) {
        CK(cudaMemset(dD, 0, 128 * 64 * 4));
        probe_mn<<<1, 128, 41 * 1024>>>(d0, d1, dB, dD, sa, sb, mode);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k) {
              float a;
              if (mode == 0) a = (m < 64) ? at(T0f, k + sa, m) : at(T1f, k + sa, m - 64);
              else a = (m < 64) ? at(T0f, k + sa, m) : at(T0f, k + sa + 1, m - 64);
              ref += (double)a * at(Bf, k + sb, n);
            }
            maxerr = fmax(maxerr, fabs(ref - D[m * 64 + n]));
          }
        printf("MN mode=%d shiftA=%d shiftB=%d maxerr=%g\n", mode, sa, sb, maxerr);
      }
}

int main(int argc, char** argv) {
  if (argc > 1 && argv[1][0] == 'm') { run_mn(); return 0; }
  std::vector<__nv_bfloat16> A(AROWS * 64), B(128 * 64);
  std::vector<float> Af(AROWS * 64), Bf(128 * 64);
  srand(1);
  for (size_t i = 0; i < A.size(); ++i) { float v = (float)((rand() % 17) - 8) / 8.f; A[i] = __float2bfloat16(v); Af[i] = __bfloat162float(A[i]); }
  for (size_t i = 0; i < B.size(); ++i) { float v = (float)((rand() % 17) - 8) / 8.f; B[i] = __float2bfloat16(v); Bf[i] = __bfloat162float(B[i]); }
  __nv_bfloat16 *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, 128 * 128 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024));
  std::vector<float> D(128 * 128);
  for (int N : {64, 128})
    for (int bo = 0; bo < 2; ++bo)
      for (int shift : {0, 1, 2, 3, 5, 8, 9, 30}) {
        CK(cudaMemset(dD, 0, 128 * 128 * 4));
        probe<<<1, 128, 50 * 1024>>>(dA, dB, dD, shift, bo, N);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k) ref += (double)Af[(m + shift) * 64 + k] * Bf[n * 64 + k];
            maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
          }
        printf("N=%d base_off=%d shift=%d maxerr=%g\n", N, bo, shift, maxerr);
      }
  return 0;
}
