// Sparse-mask engine kernels (sparse_core.py:316-333, funcs.py:107-114).
//  * multi-tensor mask apply: one launch multiplies every masked parameter (and, for SGD,
//    its momentum buffer) by its 0/1 mask -- the reference issues one elementwise kernel
//    per tensor per step.
//  * magnitude prune: radix select of the k-th smallest key (|w| bits, flat index) over
//    eight 8-bit digits, entirely on device (no .item() host syncs, no full sort).
#include "common.cuh"

namespace slak {

__global__ void mask_apply_kernel(float* const* __restrict__ w_ptrs,
                                  const float* const* __restrict__ m_ptrs,
                                  float* const* __restrict__ e_ptrs,
                                  const int64_t* __restrict__ numels) {
  const int t = blockIdx.y;
  float* w = w_ptrs[t];
  const float* m = m_ptrs[t];
  float* e = e_ptrs ? e_ptrs[t] : nullptr;
  const int64_t n = numels[t];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(e)) & 15) == 0;
  if (vec) {
    const int64_t n4 = n >> 2;
    float4* w4 = reinterpret_cast<float4*>(w);
    const float4* m4 = reinterpret_cast<const float4*>(m);
    float4* e4 = reinterpret_cast<float4*>(e);
    for (int64_t j = i; j < n4; j += stride) {
      float4 a = w4[j], b = __ldg(m4 + j);
      a.x *= b.x; a.y *= b.y; a.z *= b.z; a.w *= b.w;
      w4[j] = a;
      if (e) {
        float4 c = e4[j];
        c.x *= b.x; c.y *= b.y; c.z *= b.z; c.w *= b.w;
        e4[j] = c;
      }
    }
    for (int64_t j = (n4 << 2) + i; j < n; j += stride) {
      w[j] *= m[j];
      if (e) e[j] *= m[j];
    }
  } else {
    for (int64_t j = i; j < n; j += stride) {
      w[j] *= m[j];
      if (e) e[j] *= m[j];
    }
  }
}

int mask_apply(float* const* w_ptrs, const float* const* m_ptrs, float* const* e_ptrs,
               const int64_t* numels, int count, int64_t max_numel, cudaStream_t st) {
  if (count == 0) return SLAK_OK;
  int bx = (int)((max_numel / 4 + 255) / 256);
  if (bx < 1) bx = 1;
  const int cap = (4 * sm_count() + count - 1) / count + 1;
  if (bx > cap) bx = cap;
  dim3 grid(bx, count);
  mask_apply_kernel<<<grid, 256, 0, st>>>(w_ptrs, m_ptrs, e_ptrs, numels);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

// ---- magnitude prune -------------------------------------------------------------
struct PruneState {
  unsigned long long prefix;      // high digits of the k-th smallest key found so far
  long long k_rem;                // 1-based rank still to resolve inside the prefix bucket
  unsigned int hist[256];
};

// Selection key of element i (ascending order = selection order, low word = index: ties go to the lower index, like a
// stable sort):
//   mode 0  magnitude prune (funcs.py:107-114): |w| ascending                        -> the k smallest lose their mask
//   mode 1  growth by score (gradient_growth / momentum_growth, funcs.py:196-299):
//           |score| * (mask == 0) DESCENDING (inverted bits)                          -> the k largest get a mask
//   mode 2  plain k-th largest |x| (SNIP's global threshold, sparse_core.py:36-38)
__device__ __forceinline__ unsigned long long select_key(int mode, const float* __restrict__ w, const float* __restrict__ mask,
                                                         int64_t idx) {
  unsigned int mag = __float_as_uint(w[idx]) & 0x7fffffffu;  // |w| bits: monotone for non-negatives, NaN last
  if (mode == 1 && mask[idx] != 0.f) mag = 0u;
  if (mode != 0) mag = ~mag;
  return ((unsigned long long)mag << 32) | (unsigned long long)(unsigned int)idx;
}

__global__ void prune_init_kernel(PruneState* st, long long k) {
  if (threadIdx.x == 0) { st->prefix = 0ull; st->k_rem = k; }
  st->hist[threadIdx.x] = 0u;
}

__global__ void prune_hist_kernel(const float* __restrict__ w, const float* __restrict__ mask, int64_t n, PruneState* st,
                                  int pass, int mode) {
  __shared__ unsigned int sh[256];
  sh[threadIdx.x] = 0u;
  __syncthreads();
  const int shift = 56 - 8 * pass;
  const unsigned long long prefix = st->prefix;
  const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    unsigned long long key = select_key(mode, w, mask, i);
    if ((key & himask) == (prefix & himask)) atomicAdd(&sh[(key >> shift) & 0xff], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], sh[threadIdx.x]);
}

__global__ void prune_pick_kernel(PruneState* st, int pass) {
  // single block of 256 threads; integer counts -> exact and order independent
  __shared__ unsigned int h[256];
  h[threadIdx.x] = st->hist[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    long long k = st->k_rem;
    long long cum = 0;
    int b = 0;
    for (; b < 256; ++b) {
      if (cum + (long long)h[b] >= k) break;
      cum += h[b];
    }
    if (b > 255) b = 255;
    const int shift = 56 - 8 * pass;
    st->prefix |= ((unsigned long long)b << shift);
    st->k_rem = k - cum;
  }
  __syncthreads();
  st->hist[threadIdx.x] = 0u;
}

__global__ void prune_write_kernel(const float* __restrict__ w, float* __restrict__ mask, int64_t n,
                                   const PruneState* st, int mode) {
  const unsigned long long kth = st->prefix;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (select_key(mode, w, mask, i) <= kth) mask[i] = mode == 0 ? 0.f : 1.f;
}
// mode 2: |x| of the k-th largest element
__global__ void select_value_kernel(const PruneState* st, float* out) {
  *out = __uint_as_float(~(unsigned int)(st->prefix >> 32) & 0x7fffffffu);
}

// 0/1 fp32 mask <-> bit mask (32 elements per word, element i = bit i % 32 of word i / 32): packed-mask checkpoints and
// the mask broadcast are 32x smaller than the fp32 masks the reference moves (sparse_core.py:404-407)
__global__ void mask_pack_kernel(const float* __restrict__ mask, uint32_t* __restrict__ words, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n32 = (n + 31) / 32 * 32;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += stride) {
    const unsigned int b = __ballot_sync(0xffffffffu, i < n && mask[i] != 0.f);
    if ((threadIdx.x & 31) == 0) words[i >> 5] = b;
  }
}
__global__ void mask_unpack_kernel(const uint32_t* __restrict__ words, float* __restrict__ mask, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    mask[i] = ((words[i >> 5] >> (i & 31)) & 1u) ? 1.f : 0.f;
}

__global__ void fill_kernel(float* p, int64_t n, float v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

size_t mask_prune_workspace(int64_t) { return sizeof(PruneState); }

static int blocks_for(int64_t n) {
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  return blocks < 1 ? 1 : blocks;
}
static int radix_select(int mode, const float* w, const float* mask, int64_t n, int64_t k, PruneState* ps, cudaStream_t st) {
  const int blocks = blocks_for(n);
  prune_init_kernel<<<1, 256, 0, st>>>(ps, (long long)k);
  for (int pass = 0; pass < 8; ++pass) {
    prune_hist_kernel<<<blocks, 256, 0, st>>>(w, mask, n, ps, pass, mode);
    prune_pick_kernel<<<1, 256, 0, st>>>(ps, pass);
  }
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

int mask_prune_magnitude(const float* w, float* mask, int64_t n, int64_t k, void* workspace,
                         cudaStream_t st) {
  if (k <= 0 || n <= 0) return SLAK_OK;
  const int blocks = blocks_for(n);
  if (k >= n) {
    fill_kernel<<<blocks, 256, 0, st>>>(mask, n, 0.f);
    SLAK_CUDA_TRY(cudaGetLastError());
    return SLAK_OK;
  }
  PruneState* ps = (PruneState*)workspace;
  int rc = radix_select(0, w, mask, n, k, ps, st);
  if (rc) return rc;
  prune_write_kernel<<<blocks, 256, 0, st>>>(w, mask, n, ps, 0);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

// mask[i] = 1 at the k positions of largest |score| among the positions with mask == 0 (active positions score 0)
int mask_grow_topk(const float* score, float* mask, int64_t n, int64_t k, void* workspace, cudaStream_t st) {
  if (k <= 0 || n <= 0) return SLAK_OK;
  const int blocks = blocks_for(n);
  if (k >= n) {
    fill_kernel<<<blocks, 256, 0, st>>>(mask, n, 1.f);
    SLAK_CUDA_TRY(cudaGetLastError());
    return SLAK_OK;
  }
  PruneState* ps = (PruneState*)workspace;
  int rc = radix_select(1, score, mask, n, k, ps, st);
  if (rc) return rc;
  prune_write_kernel<<<blocks, 256, 0, st>>>(score, mask, n, ps, 1);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

// out[0] = |x| of the k-th largest magnitude (1 <= k <= n)
int select_kth_largest_abs(const float* x, int64_t n, int64_t k, void* workspace, float* out, cudaStream_t st) {
  PruneState* ps = (PruneState*)workspace;
  int rc = radix_select(2, x, nullptr, n, k, ps, st);
  if (rc) return rc;
  select_value_kernel<<<1, 1, 0, st>>>(ps, out);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

int mask_pack_bits(const float* mask, uint32_t* words, int64_t n, cudaStream_t st) {
  if (n <= 0) return SLAK_OK;
  mask_pack_kernel<<<blocks_for(n), 256, 0, st>>>(mask, words, n);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}
int mask_unpack_bits(const uint32_t* words, float* mask, int64_t n, cudaStream_t st) {
  if (n <= 0) return SLAK_OK;
  mask_unpack_kernel<<<blocks_for(n), 256, 0, st>>>(words, mask, n);
  SLAK_CUDA_TRY(cudaGetLastError());
  return SLAK_OK;
}

}  // namespace slak
