// Does TMA (tile mode) accept inner-dimension coordinates that are not multiples of 16 bytes?  (loads AND stores)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probes/tma_unaligned_probe tools/probes/tma_unaligned_probe.cu -lcuda
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int INNER>
__global__ void k(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map, int c_in, int c_out, int rows) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t dst = (smem_u32(smem) + 1023u) & ~1023u, b = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(INNER * 2 * rows) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(&in_map), "r"(b), "r"(c_in), "r"(0) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(b), "r"(0) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&out_map), "r"(dst), "r"(c_out), "r"(0) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int INNER>
int run(EncodeFn enc, int width, int rows, int c_in, int c_out, CUtensorMapSwizzle sw, const char* name) {
  const size_t n = (size_t)width * rows;
  std::vector<uint16_t> h(n), o(n);
  for (size_t i = 0; i < n; ++i) h[i] = (uint16_t)(i * 7 + 3);
  uint16_t *din, *dout;
  cudaMalloc(&din, n * 2); cudaMalloc(&dout, n * 2);
  cudaMemcpy(din, h.data(), n * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xAB, n * 2);
  CUtensorMap mi, mo;
  cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)width * 2};
  cuuint32_t box[2] = {INNER, (cuuint32_t)rows}, es[2] = {1, 1};
  CUresult r1 = enc(&mi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, din, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&mo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dout, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r1 || r2) { printf("%s: encode failed %d %d\n", name, (int)r1, (int)r2); return 1; }
  k<INNER><<<1, 32, 64 * 1024>>>(mi, mo, c_in, c_out, rows);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: kernel error %s\n", name, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(o.data(), dout, n * 2, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r = 0; r < rows; ++r)
    for (int x = 0; x < width; ++x) {
      const uint16_t got = o[(size_t)r * width + x];
      uint16_t want = 0xABAB;
      const int j = x - c_out;
      if (j >= 0 && j < INNER) want = (c_in + j < width) ? h[(size_t)r * width + c_in + j] : 0;   // OOB source reads as zero
      if (got != want) { if (bad < 5) printf("  %s mismatch r=%d x=%d got %04x want %04x\n", name, r, x, got, want); ++bad; }
    }
  printf("%s: load at %d, store at %d, inner %d: %s\n", name, c_in, c_out, INNER, bad ? "MISMATCH" : "ok");
  cudaFree(din); cudaFree(dout);
  return bad != 0;
}

int main() {
  cudaSetDevice(0);
  cudaFuncSetAttribute(k<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no encode entry point\n"); return 2; }
  EncodeFn enc = (EncodeFn)p;
  int rc = 0;
  rc |= run<64>(enc, 392, 64, 0, 0, CU_TENSOR_MAP_SWIZZLE_128B, "aligned sw128");
  rc |= run<64>(enc, 392, 64, 196, 196, CU_TENSOR_MAP_SWIZZLE_128B, "8-byte offset sw128");
  rc |= run<64>(enc, 392, 64, 196 + 64, 196 + 64, CU_TENSOR_MAP_SWIZZLE_128B, "8-byte offset +64 sw128");
  rc |= run<64>(enc, 392, 64, 196 + 192, 0, CU_TENSOR_MAP_SWIZZLE_128B, "tail block (partly OOB) sw128");
  rc |= run<64>(enc, 392, 64, 49, 147, CU_TENSOR_MAP_SWIZZLE_128B, "2-byte offset sw128");
  rc |= run<48>(enc, 392, 64, 49, 49, CU_TENSOR_MAP_SWIZZLE_NONE, "2-byte offset, inner 48, no swizzle");
  rc |= run<48>(enc, 392, 64, 343, 343, CU_TENSOR_MAP_SWIZZLE_NONE, "2-byte offset, last plane, inner 48");
  printf(rc ? "RESULT: unaligned TMA coordinates NOT usable\n" : "RESULT: unaligned TMA coordinates work\n");
  return rc;
}
