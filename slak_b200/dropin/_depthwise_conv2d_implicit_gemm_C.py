"""Stand-in for the reference's pybind module `_depthwise_conv2d_implicit_gemm_C`
(cutlass/examples/19_large_depthwise_conv2d_torch_extension/frontend.cpp:3-16, setup.py:14):
the same six functions, same argument order, plus the bf16 triple.  Each one is a thin
call into libslak_b200.so through slak_b200.ops (CUDA tensors required, no fallback).

Put this directory on PYTHONPATH where the reference asks for the CUTLASS example
directory (models/SLaK.py:9-10) and the reference's own
`depthwise_conv2d_implicit_gemm.py` imports it unchanged.
"""
import torch

from slak_b200 import ops as _ops


def _expect(t, dtype, name):
    if t.dtype != dtype:
        raise RuntimeError(f"{name} expected {dtype}, got {t.dtype}")


def _fwd(dtype):
    def forward(input, weight):
        _expect(input, dtype, "input")
        _expect(weight, dtype, "weight")
        return _ops.dwconv2d_forward(input, weight)
    return forward


def _bwd_data(dtype):
    def backward_data(grad, weight):
        _expect(grad, dtype, "grad")
        _expect(weight, dtype, "weight")
        return _ops.dwconv2d_backward_data(grad, weight)
    return backward_data


def _bwd_filter(dtype):
    def backward_filter(grad, input, weight):
        _expect(grad, dtype, "grad")
        _expect(input, dtype, "input")
        # fp32 result for every dtype (backward_filter_fp16.cu:187)
        return _ops.dwconv2d_backward_filter(grad, input, weight)
    return backward_filter


forward_fp32 = _fwd(torch.float32)
backward_data_fp32 = _bwd_data(torch.float32)
backward_filter_fp32 = _bwd_filter(torch.float32)
forward_fp16 = _fwd(torch.float16)
backward_data_fp16 = _bwd_data(torch.float16)
backward_filter_fp16 = _bwd_filter(torch.float16)
forward_bf16 = _fwd(torch.bfloat16)
backward_data_bf16 = _bwd_data(torch.bfloat16)
backward_filter_bf16 = _bwd_filter(torch.bfloat16)
