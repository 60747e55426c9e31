"""Functional CPU restatement of the reference network's hot path, over a plain
state_dict -- TEST INFRASTRUCTURE (checker for slak_b200.slak and the CPU baseline).

Every function cites the reference lines it restates (models/SLaK.py):
  depthwise op          depthwise_conv2d_implicit_gemm.py:57-66 + forward_fp32.cu:140-143
  conv_bn               :38-47 (conv, then BatchNorm when bn=True)
  reparam_large_kernel  :89-100
  layer_norm_cf         :256-261 (channels_first)
  block                 :153-166
  forward               :226-235
Pinned by tests/golden/ref_slak_tiny_*.npz, produced by oracle/gen_golden.py from the
reference's own classes.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def dwconv(x, w):
    return F.conv2d(x, w, None, 1, (w.size(2) // 2, w.size(3) // 2), 1, w.size(0))


def batch_norm(x, sd, prefix, training, eps=1e-5, momentum=0.1, update_stats=False):
    """nn.(Sync)BatchNorm forward (get_bn :24-28): batch statistics in training mode (biased
    variance for normalisation), running statistics in eval mode."""
    w, b = sd[prefix + "weight"], sd[prefix + "bias"]
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if update_stats:
            n = x.numel() / x.size(1)
            sd[prefix + "running_mean"].mul_(1 - momentum).add_(momentum * mean.detach())
            sd[prefix + "running_var"].mul_(1 - momentum).add_(momentum * var.detach() * n / max(n - 1, 1))
    else:
        mean, var = sd[prefix + "running_mean"], sd[prefix + "running_var"]
    inv = torch.rsqrt(var + eps)
    return (x - mean[None, :, None, None]) * (inv * w)[None, :, None, None] + b[None, :, None, None]


def conv_bn(x, sd, prefix, training):
    y = dwconv(x, sd[prefix + "conv.weight"])
    if (prefix + "bn.weight") in sd:
        y = batch_norm(y, sd, prefix + "bn.", training)
    return y


def reparam_large_kernel(x, sd, prefix, training):
    if (prefix + "lkb_reparam.weight") in sd:
        y = dwconv(x, sd[prefix + "lkb_reparam.weight"])
        return y + sd[prefix + "lkb_reparam.bias"].view(1, -1, 1, 1)
    if (prefix + "LoRA1.conv.weight") in sd:
        out = conv_bn(x, sd, prefix + "LoRA1.", training) + conv_bn(x, sd, prefix + "LoRA2.", training)
    else:
        out = conv_bn(x, sd, prefix + "lkb_origin.", training)
    if (prefix + "small_conv.conv.weight") in sd:
        out = out + conv_bn(x, sd, prefix + "small_conv.", training)
    return out


def layer_norm_cf(x, w, b, eps=1e-6):
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[:, None, None] * x + b[:, None, None]


def block(x, sd, prefix, training):
    """Block.forward with drop_path = identity (eval, or drop_path_rate 0)."""
    inp = x
    x = reparam_large_kernel(x, sd, prefix + "large_kernel.", training)
    x = x.permute(0, 2, 3, 1)
    x = F.layer_norm(x, (x.size(-1),), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], 1e-6)
    x = F.linear(x, sd[prefix + "pwconv1.weight"], sd[prefix + "pwconv1.bias"])
    x = F.gelu(x)
    x = F.linear(x, sd[prefix + "pwconv2.weight"], sd[prefix + "pwconv2.bias"])
    if (prefix + "gamma") in sd:
        x = sd[prefix + "gamma"] * x
    x = x.permute(0, 3, 1, 2)
    return inp + x


def forward(x, sd, depths, training=False):
    for i in range(4):
        p = f"downsample_layers.{i}."
        if i == 0:
            x = F.conv2d(x, sd[p + "0.weight"], sd[p + "0.bias"], stride=4)
            x = layer_norm_cf(x, sd[p + "1.weight"], sd[p + "1.bias"])
        else:
            x = layer_norm_cf(x, sd[p + "0.weight"], sd[p + "0.bias"])
            x = F.conv2d(x, sd[p + "1.weight"], sd[p + "1.bias"], stride=2)
        for j in range(depths[i]):
            x = block(x, sd, f"stages.{i}.{j}.", training)
    x = F.layer_norm(x.mean([-2, -1]), (x.size(1),), sd["norm.weight"], sd["norm.bias"], 1e-6)
    return F.linear(x, sd["head.weight"], sd["head.bias"])
