// sm_100a primitives used by the tensor-core kernels: mbarrier, TMA, tcgen05 (UMMA/TMEM),
// ldmatrix/stmatrix.  Thin inline-PTX wrappers, no library dependency.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>

namespace slak {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

// ---- proxies / fences ---------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- TMA ----------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// tile store smem -> global (out-of-bounds part of the box is clipped); completion is tracked by bulk async-groups
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(m), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the bulk stores of all but the N most recent groups have finished READING shared memory
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA ---------------------------------------------------------------------------------
// K-major SWIZZLE_128B shared-memory operand descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), base_offset [49,52), layout [61,64)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: bf16 x bf16 -> fp32, both operands K-major (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// all previously issued tcgen05.mma of this thread done -> one arrival on the mbarrier
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- ldmatrix / stmatrix (8x8 b16 blocks) ---------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void stmatrix_x4(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};"
               ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

}  // namespace tc
}  // namespace slak

// ---- additions: cp.async pieces, 16-column TMEM load, tile-class descriptor -------------------
namespace slak {
namespace tc {

template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const void* src) {
  static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async size");
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
}

// plane tile class chosen for an H x W plane: tile edge (64/32/16, 0 = unsupported), global piece
// size in bytes for loads/stores, and whether the TMA tiled path applies
struct TcShape { int tile; int cb; bool tma; };
TcShape tc_shape(int H, int W);
int tc_pick_splits(int C, int units);
// work partition of the channel-walking kernels over (channel, unit) items in channel-major order
struct TcPlan { int grid; int per_cta; int units_per_c; int splits; };
TcPlan tc_plan(int N, int C, int tile, int planes_per_unit);

// ---- epilogue helpers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// E consecutive fp32 values (raw bits) -> bf16 -> one store of 2*E bytes
template <int E>
__device__ __forceinline__ void store_bf16_piece(__nv_bfloat16* dst, const uint32_t* v) {
  if constexpr (E == 8) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(__uint_as_float(v[0]), __uint_as_float(v[1])),
                                                pack_bf16(__uint_as_float(v[2]), __uint_as_float(v[3])),
                                                pack_bf16(__uint_as_float(v[4]), __uint_as_float(v[5])),
                                                pack_bf16(__uint_as_float(v[6]), __uint_as_float(v[7])));
  } else if constexpr (E == 4) {
    *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(__uint_as_float(v[0]), __uint_as_float(v[1])),
                                                pack_bf16(__uint_as_float(v[2]), __uint_as_float(v[3])));
  } else if constexpr (E == 2) {
    *reinterpret_cast<uint32_t*>(dst) = pack_bf16(__uint_as_float(v[0]), __uint_as_float(v[1]));
  } else {
    *dst = __float2bfloat16_rn(__uint_as_float(v[0]));
  }
}
// E bf16 values read from global (2*E bytes) added to fp32 values held as raw bits
template <int E>
__device__ __forceinline__ void add_bf16_piece(uint32_t* v, const __nv_bfloat16* src) {
  if constexpr (E >= 2) {
    uint32_t raw[E / 2];
    if constexpr (E == 8) { const uint4 t = *reinterpret_cast<const uint4*>(src); raw[0] = t.x; raw[1] = t.y; raw[2] = t.z; raw[3] = t.w; }
    else if constexpr (E == 4) { const uint2 t = *reinterpret_cast<const uint2*>(src); raw[0] = t.x; raw[1] = t.y; }
    else { raw[0] = *reinterpret_cast<const uint32_t*>(src); }
#pragma unroll
    for (int j = 0; j < E / 2; ++j) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[j]));
      v[2 * j] = __float_as_uint(__uint_as_float(v[2 * j]) + f.x);
      v[2 * j + 1] = __float_as_uint(__uint_as_float(v[2 * j + 1]) + f.y);
    }
  } else {
    v[0] = __float_as_uint(__uint_as_float(v[0]) + __bfloat162float(*src));
  }
}
// split form of the two adds above: request the piece early (raw registers), add it later
template <int E>
__device__ __forceinline__ void ld_bf16_piece_raw(uint32_t* raw, const __nv_bfloat16* src) {   // E/2 registers
  static_assert(E >= 2, "2-byte pieces are not prefetched");
  if constexpr (E == 8) { const uint4 t = __ldg(reinterpret_cast<const uint4*>(src)); raw[0] = t.x; raw[1] = t.y; raw[2] = t.z; raw[3] = t.w; }
  else if constexpr (E == 4) { const uint2 t = __ldg(reinterpret_cast<const uint2*>(src)); raw[0] = t.x; raw[1] = t.y; }
  else { raw[0] = __ldg(reinterpret_cast<const uint32_t*>(src)); }
}
template <int E>
__device__ __forceinline__ void add_bf16_raw(uint32_t* v, const uint32_t* raw) {
#pragma unroll
  for (int j = 0; j < E / 2; ++j) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[j]));
    v[2 * j] = __float_as_uint(__uint_as_float(v[2 * j]) + f.x);
    v[2 * j + 1] = __float_as_uint(__uint_as_float(v[2 * j + 1]) + f.y);
  }
}
template <int E>
__device__ __forceinline__ void ld_f32_piece_raw(uint32_t* raw, const float* src) {            // E registers
  static_assert(E >= 2, "single elements are not prefetched");
  if constexpr (E >= 4) {
#pragma unroll
    for (int k = 0; k < E; k += 4) {
      const uint4 t = __ldg(reinterpret_cast<const uint4*>(src + k));
      raw[k] = t.x; raw[k + 1] = t.y; raw[k + 2] = t.z; raw[k + 3] = t.w;
    }
  } else {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(src));
    raw[0] = t.x; raw[1] = t.y;
  }
}
template <int E>
__device__ __forceinline__ void add_f32_raw(uint32_t* v, const uint32_t* raw) {
#pragma unroll
  for (int k = 0; k < E; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) + __uint_as_float(raw[k]));
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// fp32 variants: E floats added from / stored to global
template <int E>
__device__ __forceinline__ void add_f32_piece(uint32_t* v, const float* src) {
  if constexpr (E >= 4) {
#pragma unroll
    for (int k = 0; k < E; k += 4) {
      const float4 t = *reinterpret_cast<const float4*>(src + k);
      v[k] = __float_as_uint(__uint_as_float(v[k]) + t.x); v[k + 1] = __float_as_uint(__uint_as_float(v[k + 1]) + t.y);
      v[k + 2] = __float_as_uint(__uint_as_float(v[k + 2]) + t.z); v[k + 3] = __float_as_uint(__uint_as_float(v[k + 3]) + t.w);
    }
  } else if constexpr (E == 2) {
    const float2 t = *reinterpret_cast<const float2*>(src);
    v[0] = __float_as_uint(__uint_as_float(v[0]) + t.x); v[1] = __float_as_uint(__uint_as_float(v[1]) + t.y);
  } else {
    v[0] = __float_as_uint(__uint_as_float(v[0]) + src[0]);
  }
}
template <int E>
__device__ __forceinline__ void store_f32_piece(float* dst, const uint32_t* v) {
  if constexpr (E >= 4) {
#pragma unroll
    for (int k = 0; k < E; k += 4)
      *reinterpret_cast<float4*>(dst + k) = make_float4(__uint_as_float(v[k]), __uint_as_float(v[k + 1]),
                                                        __uint_as_float(v[k + 2]), __uint_as_float(v[k + 3]));
  } else if constexpr (E == 2) {
    *reinterpret_cast<float2*>(dst) = make_float2(__uint_as_float(v[0]), __uint_as_float(v[1]));
  } else {
    dst[0] = __uint_as_float(v[0]);
  }
}
template <int E>
__device__ __forceinline__ void copy_piece(__nv_bfloat16* dst, const uint8_t* src) {
  if constexpr (E == 8) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
  else if constexpr (E == 4) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(src);
  else if constexpr (E == 2) *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(src);
  else *dst = *reinterpret_cast<const __nv_bfloat16*>(src);
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- plane loader for the small tile classes ---------------------------------------------------
// A plane (H x W bf16, contiguous in HBM) is copied in CB-byte pieces into a T x T block of a
// SWIZZLE_128B tile: block rows [row0, row0+H), 16-byte chunk columns [cb0, cb0 + T/8).  The piece ->
// (source offset, row, byte) map is the same for every plane, so each lane precomputes its pieces once.
template <int CB>
struct PieceMap {
  static constexpr int kMax = 8;
  int count;                 // pieces of this lane (<= kMax) or -1: too many, compute on the fly
  int per_plane, PR, W;
  uint32_t soff[kMax];       // source byte offset inside the plane
  uint32_t pb[kMax];         // (row << 8) | byte offset inside the row
  __device__ __forceinline__ void init(int H, int W_, int lane) {
    W = W_;
    PR = (W * 2) / CB;
    per_plane = H * PR;
    count = 0;
    if (per_plane > kMax * 32) { count = -1; return; }
#pragma unroll
    for (int k = 0; k < kMax; ++k) {
      const int e = lane + 32 * k;
      soff[k] = 0; pb[k] = 0;
      if (e < per_plane) {
        const int p = e / PR, j = e - p * PR;
        soff[k] = (uint32_t)(p * W * 2 + j * CB);
        pb[k] = ((uint32_t)p << 8) | (uint32_t)(j * CB);
        count = k + 1;
      }
    }
  }
};

template <int CB>
__device__ __forceinline__ void copy_piece_g2s(uint32_t dst, const uint8_t* src) {
  if constexpr (CB >= 4) {
    cp_async<CB>(dst, src);
  } else {
    const uint16_t val = *reinterpret_cast<const uint16_t*>(src);
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(dst), "h"(val) : "memory");
  }
}

// copy one plane into the tile block at (row0, chunk column cb0)
template <int CB>
__device__ __forceinline__ void load_plane_block(const PieceMap<CB>& pm, const uint8_t* __restrict__ src, uint32_t tile,
                                                 int row0, int cb0, int lane) {
  if (pm.count >= 0) {
#pragma unroll
    for (int k = 0; k < PieceMap<CB>::kMax; ++k) {
      if (k < pm.count && (lane + 32 * k) < pm.per_plane) {
        const uint32_t p = pm.pb[k] >> 8, b = pm.pb[k] & 0xff;
        const uint32_t row = (uint32_t)row0 + p;
        const uint32_t dst = tile + row * 128 + ((((uint32_t)cb0 + (b >> 4)) ^ (row & 7)) << 4) + (b & 15);
        copy_piece_g2s<CB>(dst, src + pm.soff[k]);
      }
    }
  } else {
    for (int e = lane; e < pm.per_plane; e += 32) {
      const int p = e / pm.PR, j = e - p * pm.PR;
      const uint32_t b = (uint32_t)(j * CB);
      const uint32_t row = (uint32_t)(row0 + p);
      const uint32_t dst = tile + row * 128 + ((((uint32_t)cb0 + (b >> 4)) ^ (row & 7)) << 4) + (b & 15);
      copy_piece_g2s<CB>(dst, src + (size_t)p * pm.W * 2 + b);
    }
  }
}

// 2-byte pieces cannot use cp.async: issue the loads of a batch of planes first, then the stores, so that the
// global latency is paid once per batch instead of once per piece (7x7 planes: 49 pieces, 2 per lane)
template <int NB>
__device__ __forceinline__ void load_plane_blocks_cb2(const PieceMap<2>& pm, const uint8_t* const* src, uint32_t tile,
                                                      const int* row0, const int* cb0, int nplanes, int lane) {
  uint16_t val[NB][2];
#pragma unroll
  for (int q = 0; q < NB; ++q)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      val[q][k] = 0;
      if (q < nplanes && k < pm.count && (lane + 32 * k) < pm.per_plane)
        val[q][k] = *reinterpret_cast<const uint16_t*>(src[q] + pm.soff[k]);
    }
#pragma unroll
  for (int q = 0; q < NB; ++q)
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (q < nplanes && k < pm.count && (lane + 32 * k) < pm.per_plane) {
        const uint32_t p = pm.pb[k] >> 8, b = pm.pb[k] & 0xff;
        const uint32_t row = (uint32_t)row0[q] + p;
        const uint32_t dst = tile + row * 128 + ((((uint32_t)cb0[q] + (b >> 4)) ^ (row & 7)) << 4) + (b & 15);
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(dst), "h"(val[q][k]) : "memory");
      }
}

// ---- packed fp32 pairs: FFMA2 / FMUL2 / FADD2, two fp32 lanes per instruction on sm_100a -----------------------------
typedef unsigned long long f2;
__device__ __forceinline__ f2 mk2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 mk2u(uint32_t lo, uint32_t hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void un2(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 splat(float c) { return mk2(c, c); }
// (lo, hi) -> bf16x2 bits (round to nearest even), and back (exact)
__device__ __forceinline__ uint32_t pack2(f2 a) { float lo, hi; un2(a, lo, hi); return pack_bf16(lo, hi); }
__device__ __forceinline__ f2 unpack2(uint32_t b) { return mk2u(b << 16, b & 0xffff0000u); }
__device__ __forceinline__ float fma_sat(float a, float b, float c) { float r; asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }


template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* v) {
  if constexpr (NC == 64) { tmem_ld32(taddr, v); tmem_ld32(taddr + 32, v + 32); }
  else if constexpr (NC == 32) { tmem_ld32(taddr, v); }
  else { tmem_ld16(taddr, v); }
}

}  // namespace tc
}  // namespace slak
