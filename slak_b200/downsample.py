"""Downsampling layer between two stages (reference models/SLaK.py:194-199: LayerNorm(channels_first) followed by
Conv2d(kernel_size=2, stride=2)) as ONE autograd node on this library's kernels.

A 2 x 2 stride-2 convolution reads every input pixel exactly once: it is a GEMM over non-overlapping patches,
    out[(n, ho, wo), co] = sum_k A[(n, ho, wo), k] * Wp[co, k] + b[co],      k = ((h & 1) * 2 + (w & 1)) * Cin + ci.
The LayerNorm kernel writes its (bf16) output directly in that patch-row layout, so the convolution, its data gradient
and its weight gradient are the tcgen05 GEMMs of csrc/mlp_tc.cu and nothing is converted between NCHW and NHWC by a
library (cuDNN spends more time in nchwToNhwc / nhwcToNchw than in the convolution itself on these shapes).

forward : ln2d_patch_fwd (x fp32 NCHW -> A bf16) ; gemm_nt + bias (-> Y bf16 token-major) ; nhwc_to_nchw (-> fp32 NCHW
          residual stream of the next stage + the bf16 copy its first Block's depthwise kernels read)
backward: nchw_to_nhwc (dOut -> dY bf16, bias gradient as column sums) ; gemm_nt (dA = dY Wp) ; split-K gemm_tn
          (dWp = dY^T A) ; ln2d_patch_bwd (dA, x -> dx fp32 NCHW, dlnw, dlnb)
"""
from __future__ import annotations

import torch

from . import _lib, ops
from .block import EPI_BIAS, EPI_PLAIN, _ck, _colsum, _gemm_nt, _p, _wgrad


def fused_downsample_supported(ln, conv, x) -> bool:
    if not (x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and x.is_contiguous()):
        return False
    if not (torch.is_autocast_enabled() and torch.get_autocast_dtype('cuda') == torch.bfloat16):
        return False
    N, C, H, W = x.shape
    if tuple(conv.kernel_size) != (2, 2) or tuple(conv.stride) != (2, 2) or tuple(conv.padding) != (0, 0):
        return False
    if tuple(conv.dilation) != (1, 1) or conv.groups != 1 or conv.bias is None or conv.in_channels != C:
        return False
    Co = conv.out_channels
    if C % 8 or C > 768 or Co % 8 or Co > 768 or H % 2 or W % 2 or W >= 1024 or H * W >= (1 << 20):
        return False
    return getattr(ln, "data_format", None) == "channels_first" and ln.weight.dtype == torch.float32 and conv.weight.dtype == torch.float32


class DownsampleFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lnw, lnb, weight, bias, eps):
        ctx.set_materialize_grads(False)          # no zero-filled gradient for the (non-differentiable) bf16 copy
        lib = _lib.load()
        N, C, H, W = x.shape
        Co = weight.shape[0]
        Ho, Wo = H // 2, W // 2
        M, K = N * Ho * Wo, 4 * C
        dev, bf16 = x.device, torch.bfloat16
        with torch.cuda.device(dev):
            st = _lib.current_stream_ptr()
            A = torch.empty((M, K), dtype=bf16, device=dev)
            mean = torch.empty(N * H * W, dtype=torch.float32, device=dev)
            rstd = torch.empty_like(mean)
            with ops.timed("down_ln_fwd", (N, C, H * W)):
                _ck(lib.slak_ln2d_patch_fwd(_p(x), _p(lnw), _p(lnb), float(eps), _p(A), _p(mean), _p(rstd), N, C, H, W, st),
                    "slak_ln2d_patch_fwd")
            ops._count(1)
            # [Co, Ci, 2, 2] -> [Co, (kh, kw, ci)] bf16: the K order of the patch rows
            Wp = weight.detach().permute(0, 2, 3, 1).reshape(Co, K).to(bf16).contiguous()
            Y = torch.empty((M, Co), dtype=bf16, device=dev)
            with ops.timed("down_gemm_fwd", (M, Co, K)):
                _gemm_nt(lib, st, EPI_BIAS, A, Wp, bias.detach().float().contiguous(), None, Y, None, None, M, Co, K)
            out = torch.empty((N, Co, Ho, Wo), dtype=torch.float32, device=dev)
            out_b = torch.empty((N, Co, Ho, Wo), dtype=bf16, device=dev)
            with ops.timed("down_out_fwd", (N, Co, Ho * Wo)):
                _ck(lib.slak_nhwc_to_nchw(_p(Y), _p(out), _p(out_b), N, Co, Ho * Wo, st), "slak_nhwc_to_nchw")
            ops._count(1)
        ctx.save_for_backward(x, lnw, mean, rstd, A, Wp)
        ctx.dims = (N, C, H, W, Co)
        ctx.param_dtypes = (lnw.dtype, lnb.dtype, weight.dtype, bias.dtype)
        ctx.mark_non_differentiable(out_b)
        return out, out_b

    @staticmethod
    def backward(ctx, dout, _unused):
        if dout is None:
            return (None,) * 6
        lib = _lib.load()
        x, lnw, mean, rstd, A, Wp = ctx.saved_tensors
        N, C, H, W, Co = ctx.dims
        Ho, Wo = H // 2, W // 2
        M, K = N * Ho * Wo, 4 * C
        dev, bf16 = x.device, torch.bfloat16
        dout = dout.contiguous()
        if dout.dtype != torch.float32:
            dout = dout.float()
        with torch.cuda.device(dev):
            st = _lib.current_stream_ptr()
            parts = lib.slak_nchw_to_nhwc_parts(N, Co, Ho * Wo)
            dY = torch.empty((M, Co), dtype=bf16, device=dev)
            cp = torch.empty((parts, 2, Co), dtype=torch.float32, device=dev)
            with ops.timed("down_out_bwd", (N, Co, Ho * Wo)):
                _ck(lib.slak_nchw_to_nhwc(_p(dout), _p(dY), _p(cp), N, Co, Ho * Wo, st), "slak_nchw_to_nhwc")
            ops._count(1)
            db = _colsum(lib, cp.view(parts, 2 * Co), st).view(2, Co)[1]
            dA = torch.empty((M, K), dtype=bf16, device=dev)
            WpT = Wp.t().contiguous()                                  # [K, Co]: K-major operand of dA = dY Wp
            with ops.timed("down_gemm_bwd", (M, Co, K)):
                _gemm_nt(lib, st, EPI_PLAIN, dY, WpT, None, None, dA, None, None, M, K, Co)
                dWp = _wgrad(lib, st, dY, A, M, Co, K)                 # [Co, K] fp32
            dW = dWp.view(Co, 2, 2, C).permute(0, 3, 1, 2).contiguous()
            dx = torch.empty_like(x)
            lparts = lib.slak_ln2d_patch_bwd_parts(N, C, H, W)
            lp = torch.empty((lparts, 2, C), dtype=torch.float32, device=dev)
            with ops.timed("down_ln_bwd", (N, C, H * W)):
                _ck(lib.slak_ln2d_patch_bwd(_p(dA), _p(x), _p(lnw), _p(mean), _p(rstd), _p(dx), _p(lp), N, C, H, W, st),
                    "slak_ln2d_patch_bwd")
            ops._count(1)
            dl = _colsum(lib, lp.view(lparts, 2 * C), st).view(2, C)
        pd = ctx.param_dtypes
        return dx, dl[0].to(pd[0]), dl[1].to(pd[1]), dW.to(pd[2]), db.to(pd[3]), None


def fused_downsample(ln, conv, x):
    """LayerNorm(channels_first) + Conv2d(2, stride 2) of a downsampling layer; returns the fp32 NCHW output with its bf16
    copy attached (`_slak_bf16`, read by the first fused Block of the next stage)."""
    out, out_b = DownsampleFunction.apply(x, ln.weight, ln.bias, conv.weight, conv.bias, ln.eps)
    out._slak_bf16 = out_b
    return out


# ------------------------------------------------------------------------------------------------------------------------
# Stem: Conv2d(Cin, C, kernel_size=4, stride=4) -> LayerNorm(channels_first)   (reference models/SLaK.py:189-193)
# ------------------------------------------------------------------------------------------------------------------------
STEM_K = 64        # patch row length: 16 * Cin (<= 64) values, zero padded (one 128-byte swizzle atom of the GEMM's K loop)


def fused_stem_supported(conv, ln, x) -> bool:
    if not (x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and x.is_contiguous()) or x.requires_grad:
        return False
    if not (torch.is_autocast_enabled() and torch.get_autocast_dtype('cuda') == torch.bfloat16):
        return False
    N, Cin, H, W = x.shape
    if tuple(conv.kernel_size) != (4, 4) or tuple(conv.stride) != (4, 4) or tuple(conv.padding) != (0, 0):
        return False
    if tuple(conv.dilation) != (1, 1) or conv.groups != 1 or conv.bias is None or conv.in_channels != Cin or Cin > 4:
        return False
    C = conv.out_channels
    if C % 8 or C > 768 or H % 4 or W % 4:
        return False
    return getattr(ln, "data_format", None) == "channels_first" and ln.weight.dtype == torch.float32 and conv.weight.dtype == torch.float32


class StemFunction(torch.autograd.Function):
    """patch rows of the image -> tcgen05 GEMM + bias -> LayerNorm over token rows -> NCHW fp32 (+ bf16 copy).  The image gets no
    gradient (fused_stem_supported refuses inputs that require one)."""

    @staticmethod
    def forward(ctx, x, weight, bias, lnw, lnb, eps):
        ctx.set_materialize_grads(False)
        lib = _lib.load()
        N, Cin, H, W = x.shape
        C = weight.shape[0]
        Ho, Wo = H // 4, W // 4
        M = N * Ho * Wo
        dev, bf16 = x.device, torch.bfloat16
        with torch.cuda.device(dev):
            st = _lib.current_stream_ptr()
            A = torch.empty((M, STEM_K), dtype=bf16, device=dev)
            with ops.timed("stem_patch", (N, Cin, H * W)):
                _ck(lib.slak_patchify4(_p(x), _p(A), N, Cin, H, W, st), "slak_patchify4")
            ops._count(1)
            Wp = torch.zeros((C, STEM_K), dtype=bf16, device=dev)
            Wp[:, :Cin * 16] = weight.detach().reshape(C, Cin * 16)
            Y = torch.empty((M, C), dtype=bf16, device=dev)
            with ops.timed("stem_gemm_fwd", (M, C, STEM_K)):
                _gemm_nt(lib, st, EPI_BIAS, A, Wp, bias.detach().float().contiguous(), None, Y, None, None, M, C, STEM_K)
            out = torch.empty((N, C, Ho, Wo), dtype=torch.float32, device=dev)
            out_b = torch.empty((N, C, Ho, Wo), dtype=bf16, device=dev)
            mean = torch.empty(M, dtype=torch.float32, device=dev)
            rstd = torch.empty_like(mean)
            with ops.timed("stem_ln_fwd", (N, C, Ho * Wo)):
                _ck(lib.slak_ln_rows_fwd(_p(Y), _p(lnw), _p(lnb), float(eps), _p(out), _p(out_b), _p(mean), _p(rstd), N, C, Ho * Wo, st),
                    "slak_ln_rows_fwd")
            ops._count(1)
        ctx.save_for_backward(A, Y, lnw, mean, rstd)
        ctx.dims = (N, Cin, H, W, C)
        ctx.param_dtypes = (weight.dtype, bias.dtype, lnw.dtype, lnb.dtype)
        ctx.mark_non_differentiable(out_b)
        return out, out_b

    @staticmethod
    def backward(ctx, dout, _unused):
        if dout is None:
            return (None,) * 6
        lib = _lib.load()
        A, Y, lnw, mean, rstd = ctx.saved_tensors
        N, Cin, H, W, C = ctx.dims
        Ho, Wo = H // 4, W // 4
        M = N * Ho * Wo
        dev, bf16 = A.device, torch.bfloat16
        dout = dout.contiguous()
        if dout.dtype != torch.float32:
            dout = dout.float()
        with torch.cuda.device(dev):
            st = _lib.current_stream_ptr()
            parts = lib.slak_ln_rows_bwd_parts(N, C, Ho * Wo)
            dY = torch.empty((M, C), dtype=bf16, device=dev)
            part = torch.empty((parts, 3, C), dtype=torch.float32, device=dev)
            with ops.timed("stem_ln_bwd", (N, C, Ho * Wo)):
                _ck(lib.slak_ln_rows_bwd(_p(dout), _p(Y), _p(lnw), _p(mean), _p(rstd), _p(dY), _p(part), N, C, Ho * Wo, st),
                    "slak_ln_rows_bwd")
            ops._count(1)
            red = _colsum(lib, part.view(parts, 3 * C), st).view(3, C)
            with ops.timed("stem_gemm_bwd", (M, C, STEM_K)):
                dWp = _wgrad(lib, st, dY, A, M, C, STEM_K)             # [C, 64] fp32
            dW = dWp[:, :Cin * 16].reshape(C, Cin, 4, 4)
        pd = ctx.param_dtypes
        return None, dW.to(pd[0]), red[2].to(pd[1]), red[0].to(pd[2]), red[1].to(pd[3]), None


def fused_stem(conv, ln, x):
    out, out_b = StemFunction.apply(x, conv.weight, conv.bias, ln.weight, ln.bias, ln.eps)
    out._slak_bf16 = out_b
    return out
