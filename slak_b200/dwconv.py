"""`DepthWiseConv2dImplicitGEMM`: the reference's operator module surface
(depthwise_conv2d_implicit_gemm.py:52-66) over the sm_100a kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops

__all__ = ["DepthWiseConv2dImplicitGEMM"]


class DepthWiseConv2dImplicitGEMM(nn.Conv2d):
    """Depthwise conv, stride 1, 'same' padding applied inside the kernel.

    Like the reference, `padding` is NOT passed to nn.Conv2d (the module reports
    padding=(0, 0)); `kernel` is an int or a (kh, kw) tuple, both odd.
    fp32 and fp16 as the reference, bf16 added; other dtypes raise TypeError.
    """

    def __init__(self, channels, kernel, bias=False):
        super().__init__(channels, channels, kernel, groups=channels, bias=bias)
        kh, kw = self.kernel_size
        if kh % 2 == 0 or kw % 2 == 0:
            raise ValueError(f"kernel sides must be odd for a same-size output, got {self.kernel_size}")

    def forward(self, x):
        if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            raise TypeError("Only support fp32, fp16 and bf16, get {}".format(x.dtype))
        x = ops.depthwise_conv2d(x, self.weight)
        if self.bias is not None:
            x = x + self.bias.to(x).view(1, -1, 1, 1)
        return x
