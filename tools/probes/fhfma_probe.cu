// Issue-rate probe (sm_100a): FFMA vs FFMA2 (fma.rn.f32x2) vs FHFMA.BF16 (fma.rn.f32.bf16) vs FHADD.BF16 vs shift-unpack.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/fhfma_probe tools/probes/fhfma_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, unsigned seed, int iters) {
  float a[8]; unsigned x = seed + threadIdx.x, y = seed * 3 + threadIdx.x;
  unsigned long long p[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = (float)(threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2 * i]) << 32) | __float_as_uint(a[2 * i + 1]);
  unsigned short xl, xh, yl, yh;
  asm("mov.b32 {%0,%1}, %2;" : "=h"(xl), "=h"(xh) : "r"(x));
  asm("mov.b32 {%0,%1}, %2;" : "=h"(yl), "=h"(yh) : "r"(y));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) a[i] = fmaf(a[i], 1.0001f, 0.5f);
        if (MODE == 1 && i < 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(p[(i + 1) & 3]), "l"(p[(i + 2) & 3]));
        if (MODE == 2) asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(a[i]) : "h"((i & 1) ? xh : xl), "h"((i & 1) ? yh : yl));
        if (MODE == 3) asm volatile("add.rn.f32.bf16 %0, %1, %0;" : "+f"(a[i]) : "h"((i & 1) ? xh : xl));
        if (MODE == 4) { unsigned u = __float_as_uint(a[i]); asm volatile("shl.b32 %0, %0, 16;" : "+r"(u)); a[i] = __uint_as_float(u | 0x3f800000u); }
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int per_iter_thread_ops) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096;
  k<MODE><<<148 * 8, 256>>>(out, 1, 16);
  cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(out, 1, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_instr = (double)148 * 8 * 8 * iters * per_iter_thread_ops;   // CTAs x warps x iterations x instructions per thread
  printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM (at 1.965 GHz)\n", name, ms, warp_instr / 148 / (ms * 1e-3 * 1.965e9));
  cudaFree(out);
}
int main() {
  run<0>("FFMA", 64); run<1>("FFMA2 (f32x2)", 32); run<2>("FHFMA.BF16", 64); run<3>("FHADD.BF16", 64); run<4>("SHL+LOP", 128);
  return 0;
}
