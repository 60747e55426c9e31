"""Host-side mirror of the reference's `models/SLaK.py` for the hot path: same class names,
constructor arguments, attribute names and state_dict keys (so released checkpoints and
`sparse_core.Masking`'s name scan keep working), with the depthwise path routed through
the sm_100a kernels.

Reference map (models/SLaK.py):
  get_conv2d :21-22     conv_bn :38-47          get_bn :24-28        fuse_bn :49-58
  ReparamLargeKernelConv :60-122                Block :126-166       SLaK :168-235
  LayerNorm :237-261    SLaK_tiny/small/base/large :264-286

The Block / large-kernel path and the channels_first LayerNorm of the stem and downsampling
layers run on this library's kernels; the strided stem / downsampling convolutions and the head
stay stock PyTorch modules (out of scope).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import block as block_mod
from . import downsample as downsample_mod
from . import ops
from .dwconv import DepthWiseConv2dImplicitGEMM

use_sync_bn = True
FUSED_BLOCK = True      # route eligible Blocks through the fused node (slak_b200/block.py)
FUSED_DOWNSAMPLE = os.environ.get("SLAK_FUSED_DOWNSAMPLE", "1") == "1"   # downsampling layers through slak_b200/downsample.py


# ---- small utilities the reference takes from timm (timm is not a dependency here) --------
def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class DropPath(nn.Module):
    """Stochastic depth per sample (timm.models.layers.DropPath semantics: keep with
    probability 1-p, scale kept samples by 1/(1-p))."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.dim() - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0:
            mask.div_(keep)
        return x * mask

    def extra_repr(self):
        return f"drop_prob={self.drop_prob:.3f}"


_MODEL_REGISTRY = {}


def register_model(fn):
    _MODEL_REGISTRY[fn.__name__] = fn
    try:  # also visible to timm.create_model when timm is installed (main.py:301-312)
        from timm.models.registry import register_model as _timm_register
        return _timm_register(fn)
    except Exception:
        return fn


def create_model(name, **kwargs):
    kwargs.pop("pretrained", None)
    return _MODEL_REGISTRY[name](**kwargs)


# ---- building blocks ---------------------------------------------------------------------
def get_conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias):
    # like the reference, every conv built through here is the depthwise operator and
    # stride/padding/dilation/groups/out_channels are ignored
    return DepthWiseConv2dImplicitGEMM(in_channels, kernel_size, bias=bias)


def get_bn(channels):
    return nn.SyncBatchNorm(channels) if use_sync_bn else nn.BatchNorm2d(channels)


def conv_bn(in_channels, out_channels, kernel_size, stride, padding, groups, dilation=1, bn=True):
    if padding is None:
        padding = kernel_size // 2
    seq = nn.Sequential()
    seq.add_module("conv", get_conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, False))
    if bn:
        seq.add_module("bn", get_bn(out_channels))
    return seq


def conv_bn_relu(in_channels, out_channels, kernel_size, stride, padding, groups, dilation=1):
    seq = conv_bn(in_channels, out_channels, kernel_size, stride, padding, groups, dilation)
    seq.add_module("nonlinear", nn.ReLU())
    return seq


def fuse_bn(conv, bn):
    """Fold an eval-mode BN into the conv that feeds it: returns (kernel, bias)."""
    std = torch.sqrt(bn.running_var + bn.eps)
    # operation order of models/SLaK.py:56-58, so that merged checkpoints are bit-identical: the bias uses
    # (mean * gamma) / std, not mean * (gamma / std)
    return conv.weight * (bn.weight / std).reshape(-1, 1, 1, 1), bn.bias - bn.running_mean * bn.weight / std


class ReparamLargeKernelConv(nn.Module):
    """K x small + small x K (+ small x small) depthwise branches, each followed by BN,
    summed (Decom=True); or one K x K branch (+ small) (Decom=False); or a single merged
    conv with bias (small_kernel_merged=True)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, groups, small_kernel,
                 small_kernel_merged=False, Decom=False, bn=True):
        super().__init__()
        self.kernel_size = kernel_size
        self.small_kernel = small_kernel
        self.Decom = Decom
        pad = kernel_size // 2
        if small_kernel_merged:
            self.lkb_reparam = get_conv2d(in_channels, out_channels, kernel_size, stride, pad, 1, groups, True)
            return
        if Decom:
            self.LoRA1 = conv_bn(in_channels, out_channels, (kernel_size, small_kernel), stride, pad, groups, bn=bn)
            self.LoRA2 = conv_bn(in_channels, out_channels, (small_kernel, kernel_size), stride, pad, groups, bn=bn)
        else:
            self.lkb_origin = conv_bn(in_channels, out_channels, kernel_size, stride, pad, groups, bn=bn)
        if small_kernel is not None and small_kernel < kernel_size:
            self.small_conv = conv_bn(in_channels, out_channels, small_kernel, stride, small_kernel // 2, groups, bn=bn)

    def branches(self):
        """The conv_bn branches that are summed, in the reference's order."""
        if hasattr(self, "lkb_reparam") or hasattr(self, "lkb_reparam_v"):
            return []
        out = [self.LoRA1, self.LoRA2] if self.Decom else [self.lkb_origin]
        if hasattr(self, "small_conv"):
            out.append(self.small_conv)
        return out

    def _fused_ok(self, x):
        # the fused three-branch node needs the Decom layout with a small branch, no conv bias, fp32 taps
        return (self.Decom and hasattr(self, "small_conv") and hasattr(self, "LoRA1") and x.is_cuda and x.dim() == 4 and
                x.dtype in (torch.float32, torch.float16, torch.bfloat16) and
                self.LoRA1.conv.bias is None and self.LoRA1.conv.weight.dtype == torch.float32)

    def forward(self, inputs):
        if hasattr(self, "lkb_reparam"):
            return self.lkb_reparam(inputs)
        if hasattr(self, "lkb_reparam_v"):
            # re-parameterised Decom layout (merge_kernel below): K x 5 (+ bias) and 5 x K; one tcgen05 kernel when no
            # gradient is needed and the shape allows, else the two operator modules
            v, h = self.lkb_reparam_v, self.lkb_reparam_h
            if (inputs.is_cuda and inputs.dtype == torch.bfloat16 and inputs.dim() == 4 and not
                    (torch.is_grad_enabled() and (inputs.requires_grad or v.weight.requires_grad)) and
                    ops.lk_branches_bwd_uses_tc(inputs, self.kernel_size, 5) and self.small_kernel_or_5() == 5):
                return ops.lk_merged_forward(inputs.contiguous(), v.weight.detach(), h.weight.detach(), v.bias)
            return v(inputs) + h(inputs)
        if self._fused_ok(inputs):
            # one node for the three convolutions (x read once; tensor cores where the shape allows),
            # then the reference's per-branch BN and sum (models/SLaK.py:93-95)
            ys = ops.lk_branches(inputs, self.LoRA1.conv.weight, self.LoRA2.conv.weight, self.small_conv.conv.weight)
            outs = [b.bn(y) if hasattr(b, "bn") else y for b, y in zip((self.LoRA1, self.LoRA2, self.small_conv), ys)]
        else:
            outs = [b(inputs) for b in self.branches()]
        out = outs[0]
        for o in outs[1:]:
            out = out + o
        return out

    def small_kernel_or_5(self):
        return 5 if self.small_kernel is None else self.small_kernel

    def get_equivalent_decom(self):
        """Inference re-parameterisation of the Decom layout (the reference's
        merge_kernel, models/SLaK.py:102-122, only covers `lkb_origin`): each BatchNorm folded into its conv as fuse_bn
        does (:49-58), the small x small kernel added into the centre of the small x K one.  Returns
        (kernel K x small, kernel small x K, bias)."""
        def fold(branch):
            if hasattr(branch, "bn"):
                return fuse_bn(branch.conv, branch.bn)
            return branch.conv.weight, torch.zeros(branch.conv.weight.size(0), device=branch.conv.weight.device)
        kv, bv = fold(self.LoRA1)
        kh, bh = fold(self.LoRA2)
        bias = bv + bh
        if hasattr(self, "small_conv"):
            ks, bs = fold(self.small_conv)
            p = (self.kernel_size - self.small_kernel) // 2
            kh = kh + F.pad(ks, [p, p, 0, 0])
            bias = bias + bs
        return kv, kh, bias

    def get_equivalent_kernel_bias(self):
        eq_k, eq_b = fuse_bn(self.lkb_origin.conv, self.lkb_origin.bn)
        if hasattr(self, "small_conv"):
            small_k, small_b = fuse_bn(self.small_conv.conv, self.small_conv.bn)
            eq_b = eq_b + small_b
            p = (self.kernel_size - self.small_kernel) // 2
            eq_k = eq_k + F.pad(small_k, [p, p, p, p])
        return eq_k, eq_b

    def merge_kernel(self):
        if self.Decom and hasattr(self, "LoRA1"):
            kv, kh, bias = self.get_equivalent_decom()
            c = self.LoRA1.conv
            self.lkb_reparam_v = get_conv2d(c.in_channels, c.out_channels, tuple(kv.shape[2:]), 1, None, 1, c.groups, True)
            self.lkb_reparam_h = get_conv2d(c.in_channels, c.out_channels, tuple(kh.shape[2:]), 1, None, 1, c.groups, False)
            self.lkb_reparam_v.weight.data = kv.detach().clone()
            self.lkb_reparam_v.bias.data = bias.detach().clone()
            self.lkb_reparam_h.weight.data = kh.detach().clone()
            for name in ("LoRA1", "LoRA2", "small_conv"):
                if hasattr(self, name):
                    self.__delattr__(name)
            return
        eq_k, eq_b = self.get_equivalent_kernel_bias()
        conv = self.lkb_origin.conv
        self.lkb_reparam = get_conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride,
                                      conv.padding, conv.dilation, conv.groups, True)
        self.lkb_reparam.weight.data = eq_k
        self.lkb_reparam.bias.data = eq_b
        self.__delattr__("lkb_origin")
        if hasattr(self, "small_conv"):
            self.__delattr__("small_conv")


class LayerNorm(nn.Module):
    """LayerNorm over channels for channels_last (N,H,W,C) or channels_first (N,C,H,W)."""

    def __init__(self, normalized_shape, eps=1e-6, data_format="channels_last"):
        super().__init__()
        if data_format not in ("channels_last", "channels_first"):
            raise NotImplementedError
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps = eps
        self.data_format = data_format
        self.normalized_shape = (normalized_shape,)
        self.out_dtype_autocast = None     # set by SLaK for the downsampling layers (see forward)

    def forward(self, x):
        if self.data_format == "channels_last":
            return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        if x.is_cuda and x.dim() == 4 and x.dtype in (torch.float32, torch.bfloat16):
            # one streaming kernel over NCHW (slak_b200/csrc/layernorm2d.cu) instead of ten elementwise/reduce
            # launches.  Under autocast the reference's expression promotes to fp32 (pow/mean run in fp32), so the
            # result is fp32 unless the owner of this layer asked for its consumer's dtype (`out_dtype_autocast`:
            # the stride-2 conv behind a downsampling LayerNorm reads bf16 anyway)
            out_dtype = torch.float32
            if torch.is_autocast_enabled():
                if self.out_dtype_autocast is not None and torch.get_autocast_dtype('cuda') == self.out_dtype_autocast:
                    out_dtype = self.out_dtype_autocast
            elif x.dtype == torch.bfloat16 and self.weight.dtype == torch.bfloat16:
                out_dtype = torch.bfloat16
            return ops.layernorm2d(x, self.weight, self.bias, self.eps, out_dtype)
        u = x.mean(1, keepdim=True)
        d = x - u
        s = d.pow(2).mean(1, keepdim=True)
        x = d / torch.sqrt(s + self.eps)
        return self.weight[:, None, None] * x + self.bias[:, None, None]


class Block(nn.Module):
    """SLaK block: large-kernel depthwise -> (N,H,W,C) -> LayerNorm -> Linear(C,4C) -> GELU ->
    Linear(4C,C) -> gamma -> (N,C,H,W) -> input + DropPath."""

    def __init__(self, dim, drop_path=0.0, layer_scale_init_value=1e-6, kernel_size=(7, 7), Decom=None, bn=True):
        super().__init__()
        self.large_kernel = ReparamLargeKernelConv(dim, dim, kernel_size[0], stride=1, groups=dim,
                                                   small_kernel=kernel_size[1], small_kernel_merged=False,
                                                   Decom=Decom, bn=bn)
        self.norm = LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma = (nn.Parameter(layer_scale_init_value * torch.ones(dim), requires_grad=True)
                      if layer_scale_init_value > 0 else None)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x):
        if FUSED_BLOCK and x.is_cuda and torch.is_autocast_enabled():
            # the fused node keeps the residual stream in fp32 and NCHW-contiguous; the first Block of a stage
            # receives the (possibly bf16 / channels_last) output of the downsampling conv: bf16 -> fp32 is exact
            # and the reference's `input + gamma*x` promotes to fp32 there anyway (models/SLaK.py:161-165)
            xf = x if x.dtype == torch.float32 else x.float()
            xf = xf.contiguous()
            if block_mod.fused_block_supported(self, xf):
                return block_mod.fused_block_forward(self, xf)   # one autograd node, see slak_b200/block.py
        shortcut = x
        if x.is_cuda and torch.is_autocast_enabled():
            # the depthwise branch is autocast-eligible here (the reference pins fp32 inputs to its
            # fp32 kernel, depthwise_conv2d_implicit_gemm.py:16); the residual stream keeps x's dtype
            x = x.to(torch.get_autocast_dtype('cuda'))
        x = self.large_kernel(x)
        x = x.permute(0, 2, 3, 1)
        x = self.norm(x)
        x = self.pwconv2(self.act(self.pwconv1(x)))
        if self.gamma is not None:
            x = self.gamma * x
        x = x.permute(0, 3, 1, 2)
        return shortcut + self.drop_path(x)


class SLaK(nn.Module):
    def __init__(self, in_chans=3, num_classes=1000, depths=[3, 3, 9, 3], dims=[96, 192, 384, 768],
                 drop_path_rate=0.0, layer_scale_init_value=1e-6, head_init_scale=1.0,
                 kernel_size=[51, 49, 47, 13, 5], width_factor=1.0, Decom=None, bn=True):
        super().__init__()
        dims = [int(d * width_factor) for d in dims]
        self.kernel_size = kernel_size
        self.downsample_layers = nn.ModuleList()
        self.downsample_layers.append(nn.Sequential(
            nn.Conv2d(in_chans, dims[0], kernel_size=4, stride=4),
            LayerNorm(dims[0], eps=1e-6, data_format="channels_first")))
        for i in range(3):
            self.downsample_layers.append(nn.Sequential(
                LayerNorm(dims[i], eps=1e-6, data_format="channels_first"),
                nn.Conv2d(dims[i], dims[i + 1], kernel_size=2, stride=2)))
            # the conv behind this LayerNorm is autocast-eligible: let the LayerNorm kernel round to bf16 itself
            self.downsample_layers[-1][0].out_dtype_autocast = torch.bfloat16

        rates = [r.item() for r in torch.linspace(0, drop_path_rate, sum(depths))]
        self.stages = nn.ModuleList()
        at = 0
        for i in range(4):
            blocks = [Block(dim=dims[i], drop_path=rates[at + j], layer_scale_init_value=layer_scale_init_value,
                            kernel_size=(kernel_size[i], kernel_size[-1]), Decom=Decom, bn=bn)
                      for j in range(depths[i])]
            for b_ in blocks[:-1]:
                b_._slak_emit_bf16 = True      # its successor is a Block: the fused residual kernel also writes the bf16 copy
            self.stages.append(nn.Sequential(*blocks))
            at += depths[i]

        self.norm = nn.LayerNorm(dims[-1], eps=1e-6)
        self.head = nn.Linear(dims[-1], num_classes)
        self.apply(self._init_weights)
        self.head.weight.data.mul_(head_init_scale)
        self.head.bias.data.mul_(head_init_scale)

    def _init_weights(self, m):
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def forward_features(self, x):
        for i, (down, stage) in enumerate(zip(self.downsample_layers, self.stages)):
            if (i > 0 and FUSED_DOWNSAMPLE and x.is_cuda and torch.is_autocast_enabled()
                    and downsample_mod.fused_downsample_supported(down[0], down[1], x)):
                # LayerNorm + 2x2 stride-2 conv as LayerNorm -> patch rows -> tcgen05 GEMM (slak_b200/downsample.py)
                x = downsample_mod.fused_downsample(down[0], down[1], x)
            elif (i == 0 and FUSED_DOWNSAMPLE and x.is_cuda and torch.is_autocast_enabled()
                    and downsample_mod.fused_stem_supported(down[0], down[1], x)):
                # 4x4 stride-4 conv + LayerNorm as patch rows -> tcgen05 GEMM -> LayerNorm over token rows
                x = downsample_mod.fused_stem(down[0], down[1], x)
            else:
                x = down(x)
            x = stage(x)
        return self.norm(x.mean([-2, -1]))

    def forward(self, x):
        return self.head(self.forward_features(x))


@register_model
def SLaK_tiny(pretrained=False, **kwargs):
    return SLaK(depths=[3, 3, 9, 3], dims=[96, 192, 384, 768], **kwargs)


@register_model
def SLaK_small(pretrained=False, **kwargs):
    return SLaK(depths=[3, 3, 27, 3], dims=[96, 192, 384, 768], **kwargs)


@register_model
def SLaK_base(pretrained=False, in_22k=False, **kwargs):
    return SLaK(depths=[3, 3, 27, 3], dims=[128, 256, 512, 1024], **kwargs)


@register_model
def SLaK_large(pretrained=False, in_22k=False, **kwargs):
    return SLaK(depths=[3, 3, 27, 3], dims=[192, 384, 768, 1536], **kwargs)
