// Fused multi-tensor AdamW + mask apply + mask-aware EMA:
//   optimizer.step()                       torch.optim.AdamW as created by optim_factory.py:149-150
//   p.data = p.data * mask                 sparse_core.py:322-333 (Masking.apply_mask, every step)
//   ModelEma.update(model, mask)           model_sema.py:67-91
// are three passes over the same 30 - 96 M parameters in the reference (the EMA one walks the state_dict in Python);
// here every element of every tensor is read and written once, in ONE launch over a flat list of
// (tensor, offset) chunks, with the step counter on the device (CUDA-graph replayable).
//
// Per element, in fp32, in the operation order of torch's single-tensor AdamW (torch/optim/adam.py
// _single_tensor_adam with decoupled_weight_decay):
//   p   = p * (1 - lr * wd)
//   m   = lerp(m, g, 1 - beta1)                      weight < 0.5 form: m + w * (g - m)
//   v   = v * beta2 + (1 - beta2) * (g * g)
//   den = sqrt(v) * float(1 / sqrt(1 - beta2^t)) + eps   (torch divides by a Python scalar through its double reciprocal)
//   p   = p + (-(lr / (1 - beta1^t))) * (m / den)
//   p   = p * mask                                    IEEE multiply: a pruned negative weight becomes -0.0
//   ema = (ema * d + p * (1 - d)) * mask + [ema == 0 and mask != 0] * d * p      (masked tensors)
//   ema =  ema * d + (1 - d) * p                                                  (the others)
#include "common.cuh"

namespace slak {

struct AdamTables {
  float* const* p; const float* const* g; float* const* m; float* const* v;
  const float* const* mask; float* const* ema;       // entries may be NULL; the tables themselves may be NULL
  const int64_t* numel;
  const double* lr; const double* wd;                 // per tensor
  const int32_t* chunk_tensor; const int64_t* chunk_off;
};

__global__ void __launch_bounds__(256)
adamw_mask_ema_kernel(AdamTables T, int chunk_elems, double beta1, double beta2, double eps, double ema_decay,
                      const int64_t* __restrict__ step_dev, int do_adam) {
  __shared__ float sc[6];
  const int t = T.chunk_tensor[blockIdx.x];
  if (threadIdx.x == 0) {
    const double step = (double)(*step_dev + 1);
    const double lr = T.lr ? T.lr[t] : 0.0, wd = T.wd ? T.wd[t] : 0.0;
    const double bc1 = 1.0 - pow(beta1, step), bc2 = 1.0 - pow(beta2, step);
    sc[0] = (float)(1.0 - lr * wd);
    sc[1] = (float)(1.0 - beta1);
    sc[2] = (float)beta2;
    sc[3] = (float)(1.0 - beta2);
    sc[4] = (float)(1.0 / sqrt(bc2));               // torch divides a tensor by a Python scalar through the reciprocal taken in double
    sc[5] = (float)(-(lr / bc1));
  }
  __syncthreads();
  const float decay_w = sc[0], w1 = sc[1], b2 = sc[2], w2 = sc[3], inv_bc2s = sc[4], neg_step = sc[5];
  const float epsf = (float)eps, d = (float)ema_decay, omd = (float)(1.0 - ema_decay);
  float* __restrict__ p = T.p[t];
  const float* __restrict__ g = do_adam ? T.g[t] : nullptr;
  float* __restrict__ m = do_adam ? T.m[t] : nullptr;
  float* __restrict__ v = do_adam ? T.v[t] : nullptr;
  const float* __restrict__ mask = T.mask ? T.mask[t] : nullptr;
  float* __restrict__ ema = T.ema ? T.ema[t] : nullptr;
  const int64_t n = T.numel[t], off = T.chunk_off[blockIdx.x];
  const int64_t end = off + chunk_elems < n ? off + chunk_elems : n;
  for (int64_t i = off + threadIdx.x; i < end; i += blockDim.x) {
    float pv = p[i];
    if (do_adam) {
      const float gv = g[i];
      float mv = m[i], vv = v[i];
      pv = pv * decay_w;
      mv = __fmaf_rn(w1, gv - mv, mv);
      vv = __fmaf_rn(w2, __fmul_rn(gv, gv), __fmul_rn(vv, b2));      // addcmul: self + value * (t1 * t2)
      const float den = __fadd_rn(__fmul_rn(__fsqrt_rn(vv), inv_bc2s), epsf);
      pv = __fmaf_rn(neg_step, __fdiv_rn(mv, den), pv);
      m[i] = mv;
      v[i] = vv;
    }
    float mk = 1.f;
    if (mask) { mk = mask[i]; pv = pv * mk; }
    if (do_adam || mask) p[i] = pv;
    if (ema) {
      const float ev = ema[i];
      float ne = __fadd_rn(__fmul_rn(ev, d), __fmul_rn(pv, omd));
      if (mask) {
        ne = ne * mk;
        if (ev == 0.f && mk != 0.f) ne = __fadd_rn(ne, __fmul_rn(d, pv));   // newly grown weight: EMA restarts at the weight
      }
      ema[i] = ne;
    }
  }
}

__global__ void step_increment_kernel(int64_t* step) { *step += 1; }

int adamw_mask_ema(float* const* p, const float* const* g, float* const* m, float* const* v, const float* const* mask,
                   float* const* ema, const int64_t* numel, const double* lr, const double* wd, const int32_t* chunk_tensor,
                   const int64_t* chunk_off, int nchunks, int chunk_elems, double beta1, double beta2, double eps,
                   double ema_decay, int64_t* step_dev, int do_adam, cudaStream_t st) {
  if (nchunks <= 0) return SLAK_OK;
  AdamTables T{p, g, m, v, mask, ema, numel, lr, wd, chunk_tensor, chunk_off};
  adamw_mask_ema_kernel<<<nchunks, 256, 0, st>>>(T, chunk_elems, beta1, beta2, eps, ema_decay, step_dev, do_adam);
  SLAK_CUDA_TRY(cudaGetLastError());
  if (do_adam) {
    step_increment_kernel<<<1, 1, 0, st>>>(step_dev);
    SLAK_CUDA_TRY(cudaGetLastError());
  }
  return SLAK_OK;
}

}  // namespace slak
