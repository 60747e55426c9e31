"""Fused SLaK Block (models/SLaK.py:153-166 with the Decom large-kernel branch of :89-100) as ONE
autograd node over the C ABI:

  forward   cast -> [tcgen05] three depthwise branches + BN statistics -> BN finalize (+ SyncBN
            all-reduce of the per-channel sums) -> BN apply + sum + NCHW->NHWC + LayerNorm (1 pass) ->
            Linear / GELU / Linear (cuBLAS through torch.mm, bf16) -> gamma * . + residual (1 pass)
  backward  the mirror image: residual/gamma -> MLP backward -> LayerNorm backward + BatchNorm
            reductions (1 pass) -> BN backward apply for the 3 branches (1 pass) -> [tcgen05] fused
            dgrad + wgrad

Numerics: bf16 activations between the stages, fp32 statistics / parameters / residual stream --
the dataflow of the reference under autocast with the depthwise branch made autocast-eligible.
Used by slak_b200.slak.Block when the layout allows (CUDA, fp32 residual stream, autocast bf16,
Decom with the small branch, BN on, tensor-core plane sizes); everything else takes the
module-by-module path.
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.nn.functional as F

from . import _lib, ops, syncbn


def _p(t):
    return None if t is None else t.data_ptr()


def _ptr3(a, b, c):
    return (ctypes.c_void_p * 3)(_p(a), _p(b), _p(c))


def _ck(rc, what):
    _lib.check(rc, what)


# the pointwise MLP runs on this library's tcgen05 GEMMs (csrc/mlp_tc.cu); SLAK_FUSED_MLP=0 routes it through
# torch.mm (cuBLAS) + the separate GELU passes instead (also taken for widths the kernels do not cover)
FUSED_MLP = os.environ.get("SLAK_FUSED_MLP", "1") == "1"
EPI_FC1, EPI_BIAS, EPI_DGELU, EPI_PLAIN = 0, 1, 2, 3


def _fused_mlp_ok(M, C):
    """Shapes the tcgen05 MLP kernels take: 16-byte rows for the tensor maps, hidden width within the DGELU epilogue's
    shared-memory column accumulators."""
    return M > 0 and C % 8 == 0 and 4 * C <= 3072


def _gemm_nt(lib, st, epi, a, b, bias, aux_h, out0, out1, colpart, M, N, K):
    _ck(lib.slak_mlp_gemm_nt(epi, _p(a), _p(b), _p(bias), _p(aux_h), _p(out0), _p(out1), _p(colpart), M, N, K, st),
        "slak_mlp_gemm_nt")
    ops._count(1)


def _wgrad(lib, st, p, q, M, Ma, Nb):
    """dW[Ma, Nb] = p[M, Ma]^T q[M, Nb] (fp32): split-K tcgen05 GEMM over the tokens + fixed-order fold."""
    splits = lib.slak_mlp_wgrad_splits(M, Ma, Nb)
    part = torch.empty((splits, Ma * Nb), dtype=torch.float32, device=p.device)
    _ck(lib.slak_mlp_gemm_tn_splitk(_p(p), _p(q), _p(part), M, Ma, Nb, st), "slak_mlp_gemm_tn_splitk")
    ops._count(1)
    if splits == 1:
        return part.view(Ma, Nb)
    return _colsum(lib, part, st).view(Ma, Nb)


def _colsum(lib, part2d, st, out=None):
    """Fixed-order sum over the per-CTA partial rows [rows, cols] -> [cols] (one small launch of this library)."""
    rows, cols = part2d.shape
    if out is None:
        out = torch.empty(cols, dtype=torch.float32, device=part2d.device)
    _ck(lib.slak_colsum_f32(_p(part2d), rows, cols, _p(out), st), "slak_colsum_f32")
    ops._count(1)
    return out


def _dist_world(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_world_size(group)
    return None, 1


class FusedBlockFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, w2, w3, bw1, bw2, bw3, bb1, bb2, bb3, lnw, lnb, W1, b1, W2, b2, gamma, dp, cfg):
        # the second output (bf16 copy for the next Block's conv) is not differentiable: without this, autograd would
        # materialise a zero gradient of its full size (a 77 MB fill per stage-1 Block) just to hand it to backward
        ctx.set_materialize_grads(False)
        with torch.cuda.device(x.device):
            return FusedBlockFunction._forward(ctx, x, w1, w2, w3, bw1, bw2, bw3, bb1, bb2, bb3, lnw, lnb, W1, b1, W2, b2,
                                               gamma, dp, cfg)

    @staticmethod
    def _forward(ctx, x, w1, w2, w3, bw1, bw2, bw3, bb1, bb2, bb3, lnw, lnb, W1, b1, W2, b2, gamma, dp, cfg):
        lib = _lib.load()
        st = _lib.current_stream_ptr()
        N, C, H, W = x.shape
        HW = H * W
        KL = w1.size(2)
        dev = x.device
        bf16 = torch.bfloat16
        # the bf16 copy of the residual stream: handed over by the previous Block's residual kernel when there is one
        # (its `out_bf16` output), else one cast pass
        xb = cfg.get("xb")
        if xb is None or xb.shape != x.shape or xb.dtype != bf16 or not xb.is_contiguous():
            xb = x.to(bf16)
        y1, y2, y3 = torch.empty_like(xb), torch.empty_like(xb), torch.empty_like(xb)
        scale = torch.empty((3, C), dtype=torch.float32, device=dev)
        shift = torch.empty((C,), dtype=torch.float32, device=dev)
        training = cfg["training"]
        sync = cfg["sync_bn"]
        if training:
            count = float(N * HW)
            count_dev = None
            dist, world = _dist_world(cfg["process_group"]) if sync else (None, 1)
            ex = syncbn.get(cfg["process_group"], dev) if world > 1 else None
            # [C][6] sums, then this rank's element count per channel (SyncBN sums both over the ranks: ranks may hold
            # different batch sizes, torch/nn/modules/_functions.py:33-60), then the global count (peer-memory path)
            if ex is not None:
                sums_buf = ex.slot(cfg["sites"][0], C * 6 + 2, torch.float64)        # lives in the symmetric buffer
            else:
                sums_buf = torch.empty((C * 6 + 2,), dtype=torch.float64, device=dev)
            sums = sums_buf[:C * 6]
            need = lib.slak_block_conv_fwd_workspace(N, C, H, W)
            ws = ops._workspace(need, dev)
            with ops.timed("dw_fwd", (N, C, H, W, KL)):
                _ck(lib.slak_block_conv_fwd(_p(xb), _p(w1), _p(w2), _p(w3), _p(y1), _p(y2), _p(y3), _p(sums), _p(ws),
                                            ws.numel(), N, C, H, W, KL, st), "slak_block_conv_fwd")
            mean = torch.empty((3, C), dtype=torch.float32, device=dev)
            istd = torch.empty((3, C), dtype=torch.float32, device=dev)
            rm, rv = cfg["running_mean"], cfg["running_var"]
            if ex is not None:
                # SyncBatchNorm, exchange fused into the finalize kernel: one-shot all-reduce over NVLink peer memory
                site = cfg["sites"][0]
                sums_buf[C * 6:C * 6 + 1].fill_(count)
                _ck(lib.slak_bn3_finalize_fwd_sync(ex.ptrs, ex.slot_off(site), ex.flag_off(site), ex.rank, ex.world,
                                                   ex.epoch_ptr(site), _ptr3(bw1, bw2, bw3), _ptr3(bb1, bb2, bb3), _ptr3(*rm),
                                                   _ptr3(*rv), cfg["bn_eps"], cfg["bn_momentum"], C, _p(scale), _p(shift),
                                                   _p(mean), _p(istd), st), "slak_bn3_finalize_fwd_sync")
                count_dev = sums_buf[C * 6 + 1:]          # the global count, written by the kernel
                ops._count(1)
            else:
                if world > 1:                 # SyncBatchNorm over NCCL: statistics and counts of the global batch
                    sums_buf[C * 6:C * 6 + 1].fill_(count)
                    dist.all_reduce(sums_buf[:C * 6 + 1], group=cfg["process_group"])
                    count_dev = sums_buf[C * 6:C * 6 + 1]
                _ck(lib.slak_bn3_finalize_fwd(_p(sums), count, _p(count_dev), _ptr3(bw1, bw2, bw3), _ptr3(bb1, bb2, bb3),
                                              _ptr3(*rm), _ptr3(*rv), cfg["bn_eps"], cfg["bn_momentum"], C, _p(scale),
                                              _p(shift), _p(mean), _p(istd), st), "slak_bn3_finalize_fwd")
            nbts = [nbt for nbt in cfg["num_batches_tracked"] if nbt is not None]
            if nbts:
                torch._foreach_add_(nbts, 1)                  # one launch for the three counters
            ops._count(3)
        elif cfg.get("merged"):
            # re-parameterised layer: one kernel, x read once and the sum written once; w3 carries the merged bias.  The
            # LayerNorm kernel takes it as "three branches": u, u, u with scale (1, 0, 0) and no shift
            y1 = y2 = y3 = ops.lk_merged_forward(xb, w1, w2, w3)
            scale.zero_()
            scale[0].fill_(1.0)
            shift.zero_()
            mean = istd = None
            count = float(N * HW)
            count_dev = None
        else:
            ys = ops.lk_branches_forward(xb, w1, w2, w3)
            y1, y2, y3 = ys
            rm, rv = cfg["running_mean"], cfg["running_var"]
            _ck(lib.slak_bn3_eval_affine(_ptr3(bw1, bw2, bw3), _ptr3(bb1, bb2, bb3), _ptr3(*rm), _ptr3(*rv),
                                         cfg["bn_eps"], C, _p(scale), _p(shift), st), "slak_bn3_eval_affine")
            mean = istd = None
            count = float(N * HW)
            count_dev = None
            ops._count(1)
        xn = torch.empty((N, H, W, C), dtype=bf16, device=dev)
        mu = torch.empty((N * HW,), dtype=torch.float32, device=dev)
        rstd = torch.empty((N * HW,), dtype=torch.float32, device=dev)
        with ops.timed("glue_ln_fwd", (N, C, HW)):
            _ck(lib.slak_bn3_sum_ln_fwd(_p(y1), _p(y2), _p(y3), _p(scale), _p(shift), _p(lnw), _p(lnb), cfg["ln_eps"],
                                        _p(xn), _p(mu), _p(rstd), N, C, HW, st), "slak_bn3_sum_ln_fwd")
        # pointwise MLP: bf16 operands, fp32 accumulate
        M = N * HW
        xf = xn.view(M, C)
        fused_mlp = FUSED_MLP and _fused_mlp_ok(M, C)
        W1t = W2t = None
        if fused_mlp and training and W1.dtype == torch.float32 and W1.is_contiguous() and W2.is_contiguous():
            # autocast's bf16 copies of the two weights and, in the same pass, their transposes (the K-major operands of
            # the data-gradient GEMMs in backward)
            W1b, W1t = torch.empty((4 * C, C), dtype=bf16, device=dev), torch.empty((C, 4 * C), dtype=bf16, device=dev)
            W2b, W2t = torch.empty((C, 4 * C), dtype=bf16, device=dev), torch.empty((4 * C, C), dtype=bf16, device=dev)
            _ck(lib.slak_cast_transpose_bf16(_p(W1), _p(W1b), _p(W1t), 4 * C, C, st), "slak_cast_transpose_bf16")
            _ck(lib.slak_cast_transpose_bf16(_p(W2), _p(W2b), _p(W2t), C, 4 * C, st), "slak_cast_transpose_bf16")
            ops._count(2)
        else:
            W1b, W2b = W1.to(bf16), W2.to(bf16)
        with ops.timed("mlp_fwd", (M, C)):
            if fused_mlp:
                # tcgen05 GEMMs (csrc/mlp_tc.cu): pwconv1 + bias + GELU in one kernel (H and A written once, H only
                # when a backward follows), pwconv2 + bias in the second
                h = torch.empty((M, 4 * C), dtype=bf16, device=dev) if training else None
                a = torch.empty((M, 4 * C), dtype=bf16, device=dev)
                _gemm_nt(lib, st, EPI_FC1, xf, W1b, b1.float().contiguous(), None, h, a, None, M, 4 * C, C)
                h2 = torch.empty((M, C), dtype=bf16, device=dev)
                _gemm_nt(lib, st, EPI_BIAS, a, W2b, b2.float().contiguous(), None, h2, None, None, M, C, 4 * C)
            else:
                h = torch.addmm(b1.to(bf16), xf, W1b.t())
                a = F.gelu(h)
                h2 = torch.addmm(b2.to(bf16), a, W2b.t())
        out = torch.empty_like(x)
        out_b = torch.empty_like(xb) if cfg.get("emit_bf16") else None      # the next Block's conv input, written in this pass
        with ops.timed("glue_res_fwd", (N, C, HW)):
            _ck(lib.slak_block_residual_fwd(_p(x), _p(h2), _p(gamma), _p(dp), _p(out), _p(out_b), N, C, HW, st),
                "slak_block_residual_fwd")
        ops._count(2)
        ctx.cfg = cfg
        ctx.fused_mlp = fused_mlp
        ctx.count = count
        ctx.count_dev = count_dev
        ctx.dims = (N, C, H, W, KL)
        ctx.save_for_backward(xb, w1, w2, w3, bw1, bw2, bw3, y1, y2, y3, scale, shift, mean, istd, lnw, mu, rstd,
                              xn, h, a, h2, W1b, W2b, gamma, dp, W1t, W2t)
        if out_b is not None:
            ctx.mark_non_differentiable(out_b)
            return out, out_b
        return out

    @staticmethod
    def backward(ctx, dout, *unused):
        if dout is None:                           # (gradients are not materialised: the output did not reach the loss)
            return (None,) * 19
        with torch.cuda.device(dout.device):
            return FusedBlockFunction._backward(ctx, dout)

    @staticmethod
    def _backward(ctx, dout):
        (xb, w1, w2, w3, bw1, bw2, bw3, y1, y2, y3, scale, shift, mean, istd, lnw, mu, rstd, xn, h, a, h2, W1b, W2b,
         gamma, dp, W1t, W2t) = ctx.saved_tensors
        cfg = ctx.cfg
        if not cfg["training"]:
            raise RuntimeError("FusedBlockFunction.backward is only defined for training-mode BatchNorm")
        lib = _lib.load()
        st = _lib.current_stream_ptr()
        N, C, H, W, KL = ctx.dims
        HW = H * W
        dev = dout.device
        bf16 = torch.bfloat16
        dout = dout.contiguous()
        if dout.dtype != torch.float32:
            dout = dout.float()
        # ---- residual / gamma ---------------------------------------------------------------------
        parts = lib.slak_block_residual_bwd_parts(N, C, HW)
        dh2 = torch.empty((N * HW, C), dtype=bf16, device=dev)
        dgp = torch.empty((parts, 2, C), dtype=torch.float32, device=dev)
        with ops.timed("glue_res_bwd", (N, C, HW)):
            _ck(lib.slak_block_residual_bwd(_p(dout), _p(h2), _p(gamma), _p(dp), _p(dh2), _p(dgp), N, C, HW, st),
                "slak_block_residual_bwd")
        dg2 = _colsum(lib, dgp.view(parts, 2 * C), st).view(2, C)
        dgamma, db2 = dg2[0], dg2[1]
        # ---- MLP backward ---------------------------------------------------------------------------------
        K = h.shape[1]
        M = N * HW
        xf = xn.view(M, C)
        with ops.timed("mlp_bwd", (M, C)):
            if ctx.fused_mlp:
                # tcgen05 GEMMs (csrc/mlp_tc.cu): dH = (dH2 W2) * gelu'(H) with the bias-gradient partials in the
                # epilogue (dA never reaches HBM), dXn = dH W1, and the two weight gradients as split-K GEMMs over
                # the tokens with a fixed-order fold
                if W2t is None:
                    W2t = W2b.t().contiguous()             # [4C, C]: K-major B operand of dH2 W2
                    W1t = W1b.t().contiguous()             # [C, 4C]: K-major B operand of dH W1
                parts = lib.slak_mlp_parts(M, K)
                hp = torch.empty((parts, K), dtype=torch.float32, device=dev)
                dh = torch.empty_like(h)
                _gemm_nt(lib, st, EPI_DGELU, dh2, W2t, None, h, dh, None, hp, M, K, C)
                db1 = _colsum(lib, hp, st)
                dW2 = _wgrad(lib, st, dh2, a, M, C, K)
                dW1 = _wgrad(lib, st, dh, xf, M, K, C)
                dxn = torch.empty((M, C), dtype=bf16, device=dev)
                _gemm_nt(lib, st, EPI_PLAIN, dh, W1t, None, None, dxn, None, None, M, C, K)
            else:
                dW2 = torch.mm(dh2.t(), a).float()
                da = torch.mm(dh2, W2b)
                parts = lib.slak_gelu_bwd_bias_parts(M, K)
                hp = torch.empty((parts, K), dtype=torch.float32, device=dev)
                _ck(lib.slak_gelu_bwd_bias(_p(da), _p(h), _p(da), _p(hp), M, K, st), "slak_gelu_bwd_bias")   # in place
                dh = da
                del da
                db1 = _colsum(lib, hp, st)
                dW1 = torch.mm(dh.t(), xf).float()
                dxn = torch.mm(dh, W1b)
        del dh
        # ---- LayerNorm backward + BatchNorm reductions ------------------------------------------------
        parts = lib.slak_bn3_sum_ln_bwd_parts(N, C, HW)
        du = torch.empty_like(xb)
        part = torch.empty((parts, 6, C), dtype=torch.float32, device=dev)
        with ops.timed("glue_ln_bwd", (N, C, HW)):
            _ck(lib.slak_bn3_sum_ln_bwd(_p(dxn), _p(y1), _p(y2), _p(y3), _p(scale), _p(shift), _p(lnw), _p(mu), _p(rstd),
                                        _p(du), _p(part), N, C, HW, st), "slak_bn3_sum_ln_bwd")
        coef = torch.empty((9, C), dtype=torch.float32, device=dev)
        dbnw = torch.empty((3, C), dtype=torch.float32, device=dev)
        dbnb = torch.empty((3, C), dtype=torch.float32, device=dev)
        dist, world = _dist_world(cfg["process_group"]) if cfg["sync_bn"] else (None, 1)
        ex = syncbn.get(cfg["process_group"], dev) if world > 1 else None
        if ex is not None:
            # the fold writes [dlnw, dlnb, S(4 x C)] straight into this site's slot of the symmetric buffer; the finalize
            # kernel exchanges S with the peers (global sums for dy, this rank's own sums for the BN parameter gradients,
            # as torch's SyncBatchNorm: its all-reduce comes after grad_weight / grad_bias)
            site = cfg["sites"][1]
            red = _colsum(lib, part.view(parts, 6 * C), st, out=ex.slot(site, 6 * C, torch.float32)).view(6, C)
            dlnw, dlnb = red[0].clone(), red[1].clone()
            _ck(lib.slak_bn3_finalize_bwd_sync(ex.ptrs, ex.slot_off(site) + 2 * C * 4, ex.flag_off(site), ex.rank, ex.world,
                                               ex.epoch_ptr(site), _p(ctx.count_dev), _ptr3(bw1, bw2, bw3), _p(mean), _p(istd),
                                               C, _p(coef), _p(dbnw), _p(dbnb), st), "slak_bn3_finalize_bwd_sync")
        else:
            red = _colsum(lib, part.view(parts, 6 * C), st).view(6, C)
            dlnw, dlnb = red[0], red[1]
            S = red[2:6].contiguous()
            S_local = None
            if world > 1:
                # dy needs the GLOBAL sums; the BN weight / bias gradients are taken from this rank's own sums, as
                # torch's SyncBatchNorm does (its all-reduce comes after grad_weight / grad_bias), so that the
                # data-parallel gradient average gives global_sum / world and not the global sum itself
                S_local = S.clone()
                dist.all_reduce(S, group=cfg["process_group"])
            _ck(lib.slak_bn3_finalize_bwd(_p(S), _p(S_local), ctx.count, _p(ctx.count_dev), _ptr3(bw1, bw2, bw3),
                                          _p(mean), _p(istd), C, _p(coef), _p(dbnw), _p(dbnb), st), "slak_bn3_finalize_bwd")
        dy1, dy2, dy3 = torch.empty_like(xb), torch.empty_like(xb), torch.empty_like(xb)
        with ops.timed("glue_bwd_apply", (N, C, HW)):
            _ck(lib.slak_bn3_bwd_apply(_p(du), _p(y1), _p(y2), _p(y3), _p(coef), _p(dy1), _p(dy2), _p(dy3), N, C, HW, st),
                "slak_bn3_bwd_apply")
        del du
        ops._count(4)
        # ---- depthwise branches: fused tensor-core dgrad / wgrad ------------------------------------------
        dx = torch.empty_like(dout)               # shortcut + branch gradient, fp32, written by the dgrad epilogue
        tmp = torch.empty_like(dy3)
        with ops.timed("dw_dgrad", (N, C, H, W, KL)):
            _ck(lib.slak_lk_branches_bwd_data_f32(_p(dy1), _p(dy2), _p(dy3), _p(w1), _p(w2), _p(w3), _p(dout), _p(dx),
                                                  _p(tmp), N, C, H, W, KL, 5, st), "slak_lk_branches_bwd_data_f32")
        ops._count(2)
        with ops.timed("dw_wgrad", (N, C, H, W, KL)):
            dw1, dw2, dw3 = ops.lk_branches_backward_filter(xb, dy1, dy2, dy3, KL, 5)
        return (dx, dw1, dw2, dw3, dbnw[0], dbnw[1], dbnw[2], dbnb[0], dbnb[1], dbnb[2], dlnw, dlnb,
                dW1, db1, dW2, db2, dgamma, None, None)


def fused_block_supported(block, x) -> bool:
    """Can `block` (slak_b200.slak.Block) run through FusedBlockFunction on input x?"""
    lk = block.large_kernel
    if not (x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and x.is_contiguous()):
        return False
    if not (torch.is_autocast_enabled() and torch.get_autocast_dtype('cuda') == torch.bfloat16):
        return False
    if hasattr(lk, "lkb_reparam_v"):           # re-parameterised Decom layer: inference only
        N, C, H, W = x.shape
        return (not torch.is_grad_enabled() and block.gamma is not None and lk.small_kernel_or_5() == 5 and C <= 1024 and
                block.norm.data_format == "channels_last" and
                ops.lk_branches_bwd_uses_tc(_Shape(N, C, H, W), lk.kernel_size, 5))
    if not (getattr(lk, "Decom", False) and hasattr(lk, "small_conv") and hasattr(lk, "LoRA1") and block.gamma is not None):
        return False
    if lk.small_kernel != 5 or not all(hasattr(b, "bn") for b in (lk.LoRA1, lk.LoRA2, lk.small_conv)):
        return False
    if not block.training and torch.is_grad_enabled():
        return False                           # eval-mode BN backward is not implemented in the fused node
    N, C, H, W = x.shape
    if C > 1024 or block.norm.data_format != "channels_last":
        return False
    bns = (lk.LoRA1.bn, lk.LoRA2.bn, lk.small_conv.bn)
    if not all(bn.affine and bn.weight.dtype == torch.float32 for bn in bns):
        return False
    # the fused node implements exactly one BatchNorm configuration for the three branches; anything a user may have
    # changed on an individual BN (frozen bn.eval() inside a training Block, cumulative averaging with momentum=None,
    # mixed eps/momentum, SyncBatchNorm on some branches only) goes through the module-by-module path
    if not all(bn.training == block.training and bn.momentum is not None for bn in bns):
        return False
    if len({(float(bn.eps), float(bn.momentum), bool(bn.track_running_stats), type(bn)) for bn in bns}) != 1:
        return False
    if isinstance(bns[0], torch.nn.SyncBatchNorm) and len({id(bn.process_group) for bn in bns}) != 1:
        return False
    return ops.lk_branches_uses_tc(_Shape(N, C, H, W), lk.kernel_size, 5)


class _Shape:
    """Minimal stand-in carrying .shape/.dtype for ops.lk_branches_uses_tc."""

    def __init__(self, N, C, H, W):
        self.shape = (N, C, H, W)
        self.dtype = torch.bfloat16


def _sync_sites(block, group, device):
    """(forward site, backward site) of this Block in the peer-memory exchange: assigned at first use, in execution order,
    which is the same on every rank."""
    _, world = _dist_world(group)
    if world <= 1:
        return None
    ex = syncbn.get(group, device)
    if ex is None:
        return None
    if getattr(block, "_slak_sync_sites", None) is None:
        block._slak_sync_sites = (ex.new_site(), ex.new_site())
    return block._slak_sync_sites


def fused_block_forward(block, x):
    lk = block.large_kernel
    if hasattr(lk, "lkb_reparam_v"):
        v, h = lk.lkb_reparam_v, lk.lkb_reparam_h
        cfg = {"training": False, "merged": True, "sync_bn": False, "process_group": None, "bn_eps": 0.0, "bn_momentum": 0.0,
               "ln_eps": float(block.norm.eps), "running_mean": (None,) * 3, "running_var": (None,) * 3,
               "num_batches_tracked": (None,) * 3}
        return FusedBlockFunction.apply(
            x, v.weight.detach(), h.weight.detach(), v.bias.detach(), None, None, None, None, None, None,
            block.norm.weight, block.norm.bias, block.pwconv1.weight, block.pwconv1.bias, block.pwconv2.weight,
            block.pwconv2.bias, block.gamma, None, cfg)
    bns = (lk.LoRA1.bn, lk.LoRA2.bn, lk.small_conv.bn)
    sync = isinstance(bns[0], torch.nn.SyncBatchNorm)
    track = all(bn.track_running_stats and bn.running_mean is not None for bn in bns)
    training = block.training or not track
    if sync and not hasattr(block, "_slak_sync_sites"):
        block._slak_sync_sites = None
    cfg = {
        "training": training, "sync_bn": sync, "process_group": getattr(bns[0], "process_group", None) if sync else None,
        "sites": _sync_sites(block, getattr(bns[0], "process_group", None), x.device) if (sync and training) else None,
        "bn_eps": float(bns[0].eps), "bn_momentum": float(bns[0].momentum),
        "ln_eps": float(block.norm.eps),
        "running_mean": tuple(bn.running_mean if (track and block.training) or not training else None for bn in bns),
        "running_var": tuple(bn.running_var if (track and block.training) or not training else None for bn in bns),
        "num_batches_tracked": tuple(bn.num_batches_tracked if (track and block.training) else None for bn in bns),
    }
    dp = None
    p = getattr(block.drop_path, "drop_prob", 0.0)
    if block.training and p > 0.0:
        keep = 1.0 - p
        dp = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device).bernoulli_(keep).div_(keep)
    cfg["xb"] = getattr(x, "_slak_bf16", None)                 # left there by the previous fused Block
    cfg["emit_bf16"] = bool(getattr(block, "_slak_emit_bf16", False))
    res = FusedBlockFunction.apply(
        x, lk.LoRA1.conv.weight, lk.LoRA2.conv.weight, lk.small_conv.conv.weight,
        bns[0].weight, bns[1].weight, bns[2].weight, bns[0].bias, bns[1].bias, bns[2].bias,
        block.norm.weight, block.norm.bias, block.pwconv1.weight, block.pwconv1.bias, block.pwconv2.weight,
        block.pwconv2.bias, block.gamma, dp, cfg)
    if isinstance(res, tuple):
        out, out_b = res
        out._slak_bf16 = out_b
        return out
    return res
