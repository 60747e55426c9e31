"""ctypes binding of libslak_b200.so (the C ABI declared in include/slak_b200.h).

The library is loaded lazily; a missing library or a failing call raises -- there is no
CPU or PyTorch fallback anywhere in the product path.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# SLAK_B200_LIB: load another build of the same library (e.g. the -DSLAK_ROLE_PROFILE build of tools/role_profile.sh)
LIB_PATH = os.environ.get("SLAK_B200_LIB") or os.path.join(_HERE, "libslak_b200.so")

SLAK_F32, SLAK_F16, SLAK_BF16 = 0, 1, 2

_lib = None
_lock = threading.Lock()

_vp = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/slak_b200.h one to one
SIGNATURES = {
    "slak_version": (_i, []),
    "slak_last_error": (ctypes.c_char_p, []),
    "slak_device_ok": (_i, []),
    "slak_dwconv2d_uses_tc": (_i, [_i] * 8),
    "slak_dwconv2d_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "slak_dwconv2d_bwd_data": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "slak_dwconv2d_bwd_filter_workspace": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "slak_dwconv2d_bwd_filter": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "slak_lk_branches_uses_tc": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "slak_lk_branches_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "slak_lk_branches_bwd_uses_tc": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "slak_lk_branches_bwd_data": (_i, [_vp] * 8 + [_i] * 7 + [_vp]),
    "slak_lk_branches_bwd_data_f32": (_i, [_vp] * 9 + [_i] * 6 + [_vp]),
    "slak_lk_merged_fwd": (_i, [_vp] * 5 + [_i] * 6 + [_vp]),
    "slak_lk_branches_bwd_filter_workspace": (_sz, [_i] * 6),
    "slak_lk_branches_bwd_filter": (_i, [_vp] * 7 + [_i] * 7 + [_vp, _sz, _vp]),
    "slak_block_conv_fwd_workspace": (_sz, [_i] * 4),
    "slak_block_conv_fwd": (_i, [_vp] * 9 + [_sz] + [_i] * 5 + [_vp]),
    "slak_bn3_finalize_fwd": (_i, [_vp, ctypes.c_double, _vp, _vp, _vp, _vp, _vp, ctypes.c_float, ctypes.c_float, _i, _vp, _vp, _vp, _vp, _vp]),
    "slak_bn3_finalize_fwd_sync": (_i, [_vp, _sz, _sz, _i, _i, _vp, _vp, _vp, _vp, _vp, ctypes.c_float, ctypes.c_float, _i, _vp, _vp, _vp, _vp, _vp]),
    "slak_bn3_finalize_bwd_sync": (_i, [_vp, _sz, _sz, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "slak_bn3_eval_affine": (_i, [_vp] * 4 + [ctypes.c_float, _i, _vp, _vp, _vp]),
    "slak_bn3_sum_ln_fwd": (_i, [_vp] * 7 + [ctypes.c_float] + [_vp] * 3 + [_i] * 3 + [_vp]),
    "slak_block_residual_fwd": (_i, [_vp] * 6 + [_i] * 3 + [_vp]),
    "slak_ln2d_patch_fwd": (_i, [_vp] * 3 + [ctypes.c_float] + [_vp] * 3 + [_i] * 4 + [_vp]),
    "slak_ln2d_patch_bwd_parts": (_i, [_i] * 4),
    "slak_ln2d_patch_bwd": (_i, [_vp] * 7 + [_i] * 4 + [_vp]),
    "slak_nhwc_to_nchw": (_i, [_vp] * 3 + [_i] * 3 + [_vp]),
    "slak_nchw_to_nhwc_parts": (_i, [_i] * 3),
    "slak_nchw_to_nhwc": (_i, [_vp] * 3 + [_i] * 3 + [_vp]),
    "slak_patchify4": (_i, [_vp] * 2 + [_i] * 4 + [_vp]),
    "slak_ln_rows_fwd": (_i, [_vp] * 3 + [ctypes.c_float] + [_vp] * 4 + [_i] * 3 + [_vp]),
    "slak_ln_rows_bwd_parts": (_i, [_i] * 3),
    "slak_ln_rows_bwd": (_i, [_vp] * 7 + [_i] * 3 + [_vp]),
    "slak_block_residual_bwd_parts": (_i, [_i] * 3),
    "slak_block_residual_bwd": (_i, [_vp] * 6 + [_i] * 3 + [_vp]),
    "slak_gelu_bwd_bias_parts": (_i, [_i64, _i]),
    "slak_gelu_bwd_bias": (_i, [_vp] * 4 + [_i64, _i, _vp]),
    "slak_bn3_sum_ln_bwd_parts": (_i, [_i] * 3),
    "slak_bn3_sum_ln_bwd": (_i, [_vp] * 11 + [_i] * 3 + [_vp]),
    "slak_bn3_finalize_bwd": (_i, [_vp, _vp, ctypes.c_double, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "slak_bn3_bwd_apply": (_i, [_vp] * 8 + [_i] * 3 + [_vp]),
    "slak_cast_transpose_bf16": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "slak_mlp_parts": (_i, [_i, _i]),
    "slak_mlp_gemm_nt": (_i, [_i] + [_vp] * 7 + [_i] * 3 + [_vp]),
    "slak_mlp_wgrad_splits": (_i, [_i] * 3),
    "slak_mlp_gemm_tn_splitk": (_i, [_vp] * 3 + [_i] * 3 + [_vp]),
    "slak_mlp_fc1_gelu_fwd": (_i, [_vp] * 5 + [_i] * 3 + [_vp]),
    "slak_mlp_fc2_dgelu_bwd": (_i, [_vp] * 5 + [_i] * 3 + [_vp]),
    "slak_colsum_f32": (_i, [_vp, _i, _i, _vp, _vp]),
    "slak_layernorm2d_fwd": (_i, [_vp, _i, _vp, _vp, ctypes.c_float, _vp, _i, _vp, _vp, _i, _i, _i, _vp]),
    "slak_layernorm2d_bwd_parts": (_i, [_i, _i]),
    "slak_layernorm2d_bwd": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "slak_mask_apply": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _vp]),
    "slak_adamw_mask_ema_step": (_i, [_vp] * 11 + [_i, _i] + [ctypes.c_double] * 4 + [_vp, _i, _vp]),
    "slak_mask_prune_workspace": (_sz, [_i64]),
    "slak_mask_prune_magnitude": (_i, [_vp, _vp, _i64, _i64, _vp, _sz, _vp]),
    "slak_mask_grow_topk": (_i, [_vp, _vp, _i64, _i64, _vp, _sz, _vp]),
    "slak_select_kth_largest_abs": (_i, [_vp, _i64, _i64, _vp, _sz, _vp, _vp]),
    "slak_mask_pack_bits": (_i, [_vp, _vp, _i64, _vp]),
    "slak_mask_unpack_bits": (_i, [_vp, _vp, _i64, _vp]),
}


class SlakError(RuntimeError):
    pass


def load():
    """Load libslak_b200.so (building is `python -m slak_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise SlakError(
                f"{LIB_PATH} is missing: build it with `python -m slak_b200.build` "
                "(there is no CPU fallback for the slak_b200 kernels)")
        lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().slak_last_error().decode("utf-8", "replace")
        raise SlakError(f"{what} failed (status {rc}): {msg}")


def dtype_code(dtype) -> int:
    import torch
    if dtype == torch.float32:
        return SLAK_F32
    if dtype == torch.float16:
        return SLAK_F16
    if dtype == torch.bfloat16:
        return SLAK_BF16
    raise TypeError("Only support fp32, fp16 and bf16, get {}".format(dtype))


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
