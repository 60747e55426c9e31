"""Depthwise-conv oracle: ctypes front of oracle/dwconv_oracle.c (+ optional oracle/_ref)
and the torch statement the reference's own test uses.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_ARGS = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 6


def _load(path, names):
    lib = ctypes.CDLL(path)
    for n in names:
        fn = getattr(lib, n)
        fn.restype = None
        fn.argtypes = _ARGS
    return lib


_c = None
_ref = None


def c_lib():
    global _c
    if _c is None:
        p = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(p):
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle`")
        _c = _load(p, ["oracle_dwconv_fwd", "oracle_dwconv_bwd_data", "oracle_dwconv_bwd_filter"])
    return _c


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libslak_ref.so"))


def ref_lib():
    global _ref
    if _ref is None:
        _ref = _load(os.path.join(_HERE, "_ref", "libslak_ref.so"),
                     ["ref_dwconv_fwd", "ref_dwconv_bwd_data", "ref_dwconv_bwd_filter"])
    return _ref


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _call3(fn, a, b, out_shape, dims):
    a, b = _f32(a), _f32(b)
    out = np.empty(out_shape, dtype=np.float32)
    fn(a.ctypes.data, b.ctypes.data, out.ctypes.data, *dims)
    return out


def _dims(x, w):
    N, C, H, W = x.shape
    return (N, C, H, W, w.shape[2], w.shape[3])


# ---- C restatement (double accumulation) ------------------------------------------
def fwd_c(x, w):
    return _call3(c_lib().oracle_dwconv_fwd, x, w, x.shape, _dims(x, w))


def bwd_data_c(dy, w):
    return _call3(c_lib().oracle_dwconv_bwd_data, dy, w, dy.shape, _dims(dy, w))


def bwd_filter_c(dy, x, w_shape):
    dims = (*x.shape, w_shape[2], w_shape[3])
    return _call3(c_lib().oracle_dwconv_bwd_filter, dy, x, tuple(w_shape), dims)


# ---- the reference's own host code (oracle/_ref) ------------------------------------
def fwd_ref(x, w):
    return _call3(ref_lib().ref_dwconv_fwd, x, w, x.shape, _dims(x, w))


def bwd_data_ref(dy, w):
    return _call3(ref_lib().ref_dwconv_bwd_data, dy, w, dy.shape, _dims(dy, w))


def bwd_filter_ref(dy, x, w_shape):
    dims = (*x.shape, w_shape[2], w_shape[3])
    return _call3(ref_lib().ref_dwconv_bwd_filter, dy, x, tuple(w_shape), dims)


# ---- torch statement (test_correctness.py:8-9, generalised by forward_fp32.cu:140-143) --
def fwd_torch(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    return F.conv2d(x, w, None, 1, (w.size(2) // 2, w.size(3) // 2), 1, w.size(0))


def grads_torch(x: torch.Tensor, w: torch.Tensor, dy: torch.Tensor):
    """(dx, dw) of fwd_torch by autograd, on CPU, in the dtype given (use float64 for truth)."""
    x = x.detach().clone().requires_grad_(True)
    w = w.detach().clone().requires_grad_(True)
    y = fwd_torch(x, w)
    y.backward(dy)
    return x.grad, w.grad


def round_like(t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Round to `dtype` and come back to fp32/fp64 (autocast's cast_inputs on x and w)."""
    return t.to(dtype).to(t.dtype)
