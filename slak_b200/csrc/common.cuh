// Shared helpers for the slak_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/slak_b200.h"

namespace slak {

constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  return dev;
}

// ---- thread-local error message ------------------------------------------------
void set_error(const char* fmt, ...);

#define SLAK_CUDA_TRY(expr)                                                        \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      slak::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                      __FILE__, __LINE__);                                         \
      return SLAK_ERR_CUDA;                                                        \
    }                                                                              \
  } while (0)

// raise a kernel's dynamic shared-memory limit once per call site AND DEVICE (cudaFuncSetAttribute is per device;
// not on every launch: the attribute call must not happen while a CUDA graph is being captured more often than needed)
#define SLAK_SET_MAX_SMEM(kern, bytes)                                                             \
  do {                                                                                             \
    static int _slak_cur_smem[slak::kMaxDevices] = {0};                                            \
    const int _slak_dev = slak::current_device();                                                  \
    if ((int)(bytes) > _slak_cur_smem[_slak_dev]) {                                                \
      SLAK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      _slak_cur_smem[_slak_dev] = (int)(bytes);                                                    \
    }                                                                                              \
  } while (0)

#define SLAK_REQUIRE(cond, code, ...)                                              \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      slak::set_error(__VA_ARGS__);                                                \
      return (code);                                                               \
    }                                                                              \
  } while (0)

// ---- element conversion --------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// round an fp32 weight to the activation type and back (autocast semantics)
template <typename T> __device__ __forceinline__ float round_to(float v) { return to_f32<T>(from_f32<T>(v)); }

static inline int dtype_size(int dtype) { return dtype == SLAK_F32 ? 4 : 2; }

inline int sm_count() {
  static int n[kMaxDevices] = {0};
  const int dev = current_device();
  if (n[dev] == 0) {
    if (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n[dev] = 148;
  }
  return n[dev];
}

}  // namespace slak
