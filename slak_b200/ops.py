"""Tensor-level wrappers over the C ABI: torch supplies device memory and the stream,
the kernels do the work.  Everything here requires CUDA tensors; nothing falls back.

Mirrors the reference operator boundary:
  cutlass/examples/19_large_depthwise_conv2d_torch_extension/frontend.h:3-10   (the 6 functions)
  depthwise_conv2d_implicit_gemm.py:14-49                                      (autograd glue)
"""
from __future__ import annotations

import torch

from . import _lib


def _check_input(t: torch.Tensor, name: str) -> None:
    # same checks as CHECK_INPUT in forward_fp32.cu:194-196
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def _conv_dims(x: torch.Tensor, w: torch.Tensor):
    if x.dim() != 4 or w.dim() != 4 or w.size(1) != 1 or w.size(0) != x.size(1):
        raise RuntimeError(f"expected x [N,C,H,W] and weight [C,1,kh,kw], got {tuple(x.shape)} and {tuple(w.shape)}")
    N, C, H, W = x.shape
    return N, C, H, W, w.size(2), w.size(3)


_ws_cache: dict = {}

# ---- instrumentation used by bench.py (launch accounting + CUDA-event timing of one kernel) ----
_launches = 0
_prof = {"match": None, "events": [], "tagged": []}


def launch_count() -> int:
    """Number of slak_b200 CUDA kernels launched so far by this process."""
    return _launches


def _count(n: int) -> None:
    global _launches
    _launches += n


def profile_reset(match) -> None:
    """match = dict(N,C,H,W,kh,kw) of the forward launches to bracket with CUDA events, or None.
    match["all"] = True additionally brackets every depthwise / pointwise-MLP kernel group of the fused Block
    (see `timed`), collected in _prof["tagged"] as (tag, key, start_event, end_event)."""
    _prof["match"] = match
    _prof["events"] = []
    _prof["tagged"] = []


class timed:
    """`with ops.timed("dw_fwd", (N, C, H, W, KL)):` -- CUDA events around a kernel group when bench.py asked for the
    per-kernel roofline table (profile_reset({"all": True, ...})); free otherwise.  The events may sit inside a
    CUDA-graph capture (external events)."""

    __slots__ = ("tag", "key", "ev")

    def __init__(self, tag, key):
        self.tag, self.key, self.ev = tag, key, None

    def __enter__(self):
        m = _prof["match"]
        if m is not None and m.get("all"):
            self.ev = _new_event()
            self.ev.record()
        return self

    def __exit__(self, *exc):
        if self.ev is not None:
            e1 = _new_event()
            e1.record()
            _prof["tagged"].append((self.tag, self.key, self.ev, e1))
        return False


def _new_event():
    m = _prof["match"]
    if m is not None and m.get("external_events"):
        return torch.cuda.Event(enable_timing=True, external=True)    # usable inside CUDA-graph capture
    return torch.cuda.Event(enable_timing=True)


def profile_collect():
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in _prof["events"])
    return {"count": len(_prof["events"]), "ms_total": ms}


def _profiled(N, C, H, W, kh, kw, dtype):
    m = _prof["match"]
    return (m is not None and "N" in m and dtype == torch.bfloat16 and
            (N, C, H, W, kh, kw) == (m["N"], m["C"], m["H"], m["W"], m["kh"], m["kw"]))


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Per-(device, stream) scratch buffer, grown on demand."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def dwconv2d_forward(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """y = depthwise cross-correlation of x with w, stride 1, padding (kh//2, kw//2)."""
    _check_input(x, "input")
    _check_input(w, "weight")
    N, C, H, W, kh, kw = _conv_dims(x, w)
    y = torch.empty_like(x)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        rc = lib.slak_dwconv2d_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), N, C, H, W, kh, kw,
                                   _lib.dtype_code(x.dtype), _lib.dtype_code(w.dtype),
                                   _lib.current_stream_ptr())
    _lib.check(rc, "slak_dwconv2d_fwd")
    _count(1)
    return y


def dwconv2d_backward_data(dy: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    _check_input(dy, "grad")
    _check_input(w, "weight")
    N, C, H, W, kh, kw = _conv_dims(dy, w)
    dx = torch.empty_like(dy)
    lib = _lib.load()
    with torch.cuda.device(dy.device):
        rc = lib.slak_dwconv2d_bwd_data(dy.data_ptr(), w.data_ptr(), dx.data_ptr(), N, C, H, W, kh, kw,
                                        _lib.dtype_code(dy.dtype), _lib.dtype_code(w.dtype),
                                        _lib.current_stream_ptr())
    _lib.check(rc, "slak_dwconv2d_bwd_data")
    _count(1)
    return dx


def dwconv2d_backward_filter(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """dw in fp32 whatever the activation dtype (backward_filter_fp16.cu:18,187); `w` only gives the shape."""
    _check_input(dy, "grad")
    _check_input(x, "input")
    if dy.dtype != x.dtype or dy.shape != x.shape:
        raise RuntimeError("grad and input must have the same dtype and shape")
    N, C, H, W, kh, kw = _conv_dims(x, w)
    dw = torch.empty((C, 1, kh, kw), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    code = _lib.dtype_code(x.dtype)
    need = lib.slak_dwconv2d_bwd_filter_workspace(N, C, H, W, kh, kw, code)
    with torch.cuda.device(x.device):
        ws = _workspace(need, x.device)
        rc = lib.slak_dwconv2d_bwd_filter(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), N, C, H, W, kh, kw,
                                          code, ws.data_ptr(), ws.numel(), _lib.current_stream_ptr())
    _lib.check(rc, "slak_dwconv2d_bwd_filter")
    _count(2)   # partial-sum kernel + fixed-order reduce
    return dw


class DepthwiseConv2dFunction(torch.autograd.Function):
    """One autograd node for all dtypes.  dtype of y = dtype of x; the weight stays the fp32
    Parameter and is rounded to x's dtype inside the kernel (what
    custom_fwd(cast_inputs=torch.float16) does to the reference's FP16 path,
    depthwise_conv2d_implicit_gemm.py:33-38); dw comes back fp32."""

    @staticmethod
    def forward(ctx, x, w):
        x = x.contiguous()
        w = w.contiguous()
        ctx.save_for_backward(x, w)
        return dwconv2d_forward(x, w)

    @staticmethod
    def backward(ctx, grad):
        x, w = ctx.saved_tensors
        grad = grad.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = dwconv2d_backward_data(grad, w)
        if ctx.needs_input_grad[1]:
            dw = dwconv2d_backward_filter(grad, x, w).to(w.dtype)
        return dx, dw


def depthwise_conv2d(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        raise TypeError("Only support fp32, fp16 and bf16, get {}".format(x.dtype))
    if w.dtype != torch.float32 and w.dtype != x.dtype:
        w = w.to(x.dtype)
    return DepthwiseConv2dFunction.apply(x, w)


# ---- fused three-branch forward (models/SLaK.py:89-100) ------------------------------------
def lk_branches_uses_tc(x: torch.Tensor, KL: int, KS: int) -> bool:
    N, C, H, W = x.shape
    return bool(_lib.load().slak_lk_branches_uses_tc(N, C, H, W, KL, KS, _lib.dtype_code(x.dtype)))


def lk_branches_bwd_uses_tc(x: torch.Tensor, KL: int, KS: int) -> bool:
    N, C, H, W = x.shape
    return bool(_lib.load().slak_lk_branches_bwd_uses_tc(N, C, H, W, KL, KS, _lib.dtype_code(x.dtype)))


def lk_branches_forward(x, w1, w2, w3=None):
    """(y1, y2, y3) = (dwconv_{KLxKS}, dwconv_{KSxKL}, dwconv_{KSxKS})(x) with the fp32 Parameters
    w1 [C,1,KL,KS], w2 [C,1,KS,KL], w3 [C,1,KS,KS] (or None).  One tcgen05 kernel when the shape
    allows it (see slak_b200.h), otherwise the CUDA-core kernels."""
    _check_input(x, "input")
    for t, nm in ((w1, "w1"), (w2, "w2")) + (((w3, "w3"),) if w3 is not None else ()):
        _check_input(t, nm)
        if t.dtype != torch.float32:
            raise RuntimeError(f"{nm} must be the fp32 parameter")
    N, C, H, W = x.shape
    KL, KS = w1.size(2), w1.size(3)
    if tuple(w2.shape) != (C, 1, KS, KL) or (w3 is not None and tuple(w3.shape) != (C, 1, KS, KS)):
        raise RuntimeError("branch weight shapes do not match [C,1,KL,KS] / [C,1,KS,KL] / [C,1,KS,KS]")
    y1, y2 = torch.empty_like(x), torch.empty_like(x)
    y3 = torch.empty_like(x) if w3 is not None else None
    lib = _lib.load()
    code = _lib.dtype_code(x.dtype)
    tc = w3 is not None and lib.slak_lk_branches_uses_tc(N, C, H, W, KL, KS, code)
    timed = tc and _profiled(N, C, H, W, KL, KS, x.dtype)
    with torch.cuda.device(x.device):
        if timed:
            ev = (_new_event(), _new_event())
            ev[0].record()
        rc = lib.slak_lk_branches_fwd(x.data_ptr(), w1.data_ptr(), w2.data_ptr(),
                                      w3.data_ptr() if w3 is not None else None,
                                      y1.data_ptr(), y2.data_ptr(), y3.data_ptr() if y3 is not None else None,
                                      N, C, H, W, KL, KS, code, _lib.current_stream_ptr())
        if timed:
            ev[1].record()
            _prof["events"].append(ev)
    _lib.check(rc, "slak_lk_branches_fwd")
    _count(1 if tc else (3 if w3 is not None else 2))
    return y1, y2, y3


def lk_merged_forward(x, wv, wh, bias=None):
    """Inference form of the re-parameterised Decom layer: y = dwconv_{KLx5}(x, wv) + dwconv_{5xKL}(x, wh) + bias, one
    tcgen05 kernel (x read once, y written once).  bf16 tensor-core shapes only (lk_branches_bwd_uses_tc)."""
    for t, nm in ((x, "input"), (wv, "wv"), (wh, "wh")):
        _check_input(t, nm)
    N, C, H, W = x.shape
    KL = wv.size(2)
    if tuple(wv.shape) != (C, 1, KL, 5) or tuple(wh.shape) != (C, 1, 5, KL) or wv.dtype != torch.float32 or wh.dtype != torch.float32:
        raise RuntimeError("expected fp32 kernels [C,1,KL,5] and [C,1,5,KL]")
    b = None if bias is None else bias.detach().float().contiguous()
    y = torch.empty_like(x)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        rc = lib.slak_lk_merged_fwd(x.data_ptr(), wv.data_ptr(), wh.data_ptr(), None if b is None else b.data_ptr(), y.data_ptr(),
                                    N, C, H, W, KL, _lib.dtype_code(x.dtype), _lib.current_stream_ptr())
    _lib.check(rc, "slak_lk_merged_fwd")
    _count(1)
    return y


def lk_branches_backward_data(dy1, dy2, dy3, w1, w2, w3):
    """dx = dgrad(dy1,w1) + dgrad(dy2,w2) + dgrad(dy3,w3); tensor-core shapes only."""
    for t, nm in ((dy1, "dy1"), (dy2, "dy2"), (dy3, "dy3"), (w1, "w1"), (w2, "w2"), (w3, "w3")):
        _check_input(t, nm)
    N, C, H, W = dy1.shape
    KL, KS = w1.size(2), w1.size(3)
    dx = torch.empty_like(dy1)
    tmp = torch.empty_like(dy1)
    lib = _lib.load()
    with torch.cuda.device(dy1.device):
        rc = lib.slak_lk_branches_bwd_data(dy1.data_ptr(), dy2.data_ptr(), dy3.data_ptr(), w1.data_ptr(), w2.data_ptr(),
                                           w3.data_ptr(), dx.data_ptr(), tmp.data_ptr(), N, C, H, W, KL, KS,
                                           _lib.dtype_code(dy1.dtype), _lib.current_stream_ptr())
    _lib.check(rc, "slak_lk_branches_bwd_data")
    _count(2)
    return dx


def lk_branches_backward_filter(x, dy1, dy2, dy3, KL, KS):
    """(dw1, dw2, dw3) in fp32; tensor-core shapes only."""
    for t, nm in ((x, "x"), (dy1, "dy1"), (dy2, "dy2"), (dy3, "dy3")):
        _check_input(t, nm)
    N, C, H, W = x.shape
    dev = x.device
    dw1 = torch.empty((C, 1, KL, KS), dtype=torch.float32, device=dev)
    dw2 = torch.empty((C, 1, KS, KL), dtype=torch.float32, device=dev)
    dw3 = torch.empty((C, 1, KS, KS), dtype=torch.float32, device=dev)
    lib = _lib.load()
    need = lib.slak_lk_branches_bwd_filter_workspace(N, C, H, W, KL, KS)
    with torch.cuda.device(dev):
        ws = _workspace(need, dev)
        rc = lib.slak_lk_branches_bwd_filter(x.data_ptr(), dy1.data_ptr(), dy2.data_ptr(), dy3.data_ptr(), dw1.data_ptr(),
                                             dw2.data_ptr(), dw3.data_ptr(), N, C, H, W, KL, KS, _lib.dtype_code(x.dtype),
                                             ws.data_ptr(), ws.numel(), _lib.current_stream_ptr())
    _lib.check(rc, "slak_lk_branches_bwd_filter")
    _count(4)
    return dw1, dw2, dw3


class LKBranchesFunction(torch.autograd.Function):
    """(y1, y2, y3) = the three depthwise branches of ReparamLargeKernelConv (models/SLaK.py:89-100) as one
    autograd node: x is read once in forward; backward is the fused tensor-core dgrad/wgrad where the shape
    allows, else the per-branch CUDA-core kernels."""

    @staticmethod
    def forward(ctx, x, w1, w2, w3):
        x = x.contiguous()
        w1, w2, w3 = w1.contiguous(), w2.contiguous(), w3.contiguous()
        ctx.save_for_backward(x, w1, w2, w3)
        return lk_branches_forward(x, w1, w2, w3)

    @staticmethod
    def backward(ctx, g1, g2, g3):
        x, w1, w2, w3 = ctx.saved_tensors
        g1, g2, g3 = g1.contiguous(), g2.contiguous(), g3.contiguous()
        KL, KS = w1.size(2), w1.size(3)
        need_dx = ctx.needs_input_grad[0]
        need_dw = any(ctx.needs_input_grad[1:])
        dx = dw1 = dw2 = dw3 = None
        if lk_branches_bwd_uses_tc(x, KL, KS):
            if need_dx:
                dx = lk_branches_backward_data(g1, g2, g3, w1, w2, w3)
            if need_dw:
                dw1, dw2, dw3 = lk_branches_backward_filter(x, g1, g2, g3, KL, KS)
        else:
            if need_dx:
                dx = dwconv2d_backward_data(g1, w1)
                dx += dwconv2d_backward_data(g2, w2)
                dx += dwconv2d_backward_data(g3, w3)
            if need_dw:
                dw1 = dwconv2d_backward_filter(g1, x, w1)
                dw2 = dwconv2d_backward_filter(g2, x, w2)
                dw3 = dwconv2d_backward_filter(g3, x, w3)
        return dx, dw1, dw2, dw3


def lk_branches(x, w1, w2, w3):
    if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        raise TypeError("Only support fp32, fp16 and bf16, get {}".format(x.dtype))
    return LKBranchesFunction.apply(x, w1, w2, w3)


# ---------------------------------------------------------------------------------------
# LayerNorm over the channels of an NCHW tensor (models/SLaK.py:256-261, "channels_first")
# ---------------------------------------------------------------------------------------
def layernorm2d_forward(x, weight, bias, eps, out_dtype=None, need_stats=True):
    """y[n,c,h,w] = weight[c] * (x - mean_c) / sqrt(var_c + eps) + bias[c]; returns (y, mean, rstd)."""
    _check_input(x, "input")
    if x.dim() != 4 or x.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("layernorm2d expects a 4-d fp32 or bf16 NCHW tensor")
    out_dtype = out_dtype or x.dtype
    N, C, H, W = x.shape
    w32 = weight.detach().float().contiguous()
    b32 = bias.detach().float().contiguous()
    y = torch.empty((N, C, H, W), dtype=out_dtype, device=x.device)
    mean = torch.empty(N * H * W, dtype=torch.float32, device=x.device) if need_stats else None
    rstd = torch.empty_like(mean) if need_stats else None
    lib = _lib.load()
    with torch.cuda.device(x.device):
        rc = lib.slak_layernorm2d_fwd(x.data_ptr(), _lib.dtype_code(x.dtype), w32.data_ptr(), b32.data_ptr(), float(eps),
                                      y.data_ptr(), _lib.dtype_code(out_dtype),
                                      mean.data_ptr() if need_stats else None, rstd.data_ptr() if need_stats else None,
                                      N, C, H * W, _lib.current_stream_ptr())
    _lib.check(rc, "slak_layernorm2d_fwd")
    _count(1)
    return y, mean, rstd


def layernorm2d_backward(g, x, weight, mean, rstd):
    """(dx, dweight, dbias) of layernorm2d_forward; dx has x's dtype, the parameter gradients are fp32."""
    _check_input(g, "grad")
    _check_input(x, "input")
    N, C, H, W = x.shape
    w32 = weight.detach().float().contiguous()
    lib = _lib.load()
    parts = lib.slak_layernorm2d_bwd_parts(N, H * W)
    part = torch.empty(parts * 2 * C, dtype=torch.float32, device=x.device)
    dx = torch.empty_like(x)
    dw = torch.empty(C, dtype=torch.float32, device=x.device)
    db = torch.empty(C, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.slak_layernorm2d_bwd(g.data_ptr(), _lib.dtype_code(g.dtype), x.data_ptr(), _lib.dtype_code(x.dtype),
                                      w32.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dx.data_ptr(), part.data_ptr(),
                                      dw.data_ptr(), db.data_ptr(), N, C, H * W, _lib.current_stream_ptr())
    _lib.check(rc, "slak_layernorm2d_bwd")
    _count(2)
    return dx, dw, db


class LayerNorm2dFunction(torch.autograd.Function):
    """Autograd node of the channels_first LayerNorm; `out_dtype` lets the caller ask for the dtype its consumer
    wants (bf16 in front of an autocast conv) instead of a separate cast pass."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype):
        x = x.contiguous()
        y, mean, rstd = layernorm2d_forward(x, weight, bias, eps, out_dtype, need_stats=True)
        ctx.save_for_backward(x, weight, mean, rstd)
        ctx.param_dtypes = (weight.dtype, bias.dtype)
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight, mean, rstd = ctx.saved_tensors
        dx, dw, db = layernorm2d_backward(g.contiguous(), x, weight, mean, rstd)
        return dx, dw.to(ctx.param_dtypes[0]), db.to(ctx.param_dtypes[1]), None, None


def layernorm2d(x, weight, bias, eps=1e-6, out_dtype=None):
    if not x.is_cuda:
        raise RuntimeError("slak_b200.ops.layernorm2d needs CUDA tensors (no CPU fallback)")
    return LayerNorm2dFunction.apply(x, weight, bias, eps, out_dtype)
