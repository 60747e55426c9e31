"""Downsampling layer (reference models/SLaK.py:194-199, LayerNorm(channels_first) -> Conv2d(k=2, s=2)) as one autograd node
on this library's kernels (slak_b200/downsample.py) against the same two modules in fp64."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("N,Ci,Co,H,W", [(2, 16, 32, 8, 8), (3, 24, 48, 6, 10), (2, 96, 192, 56, 56), (2, 384, 768, 14, 14),
                                         (1, 40, 64, 10, 6)])
def test_fused_downsample_matches_fp64(N, Ci, Co, H, W):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from slak_b200 import downsample
    from slak_b200.slak import LayerNorm
    torch.manual_seed(Ci + H)
    dev = torch.device("cuda:0")
    ln = LayerNorm(Ci, eps=1e-6, data_format="channels_first").to(dev)
    conv = nn.Conv2d(Ci, Co, kernel_size=2, stride=2).to(dev)
    with torch.no_grad():
        ln.weight.copy_(torch.randn(Ci) * 0.2 + 1); ln.bias.copy_(torch.randn(Ci) * 0.1)
        conv.weight.copy_(torch.randn_like(conv.weight) * 0.1); conv.bias.copy_(torch.randn(Co) * 0.1)
    x = (torch.randn(N, Ci, H, W, device=dev) * 1.5 + 0.7).requires_grad_(True)
    gout = torch.randn(N, Co, H // 2, W // 2, device=dev)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert downsample.fused_downsample_supported(ln, conv, x)
        out = downsample.fused_downsample(ln, conv, x)
    assert out.dtype == torch.float32 and out._slak_bf16.dtype == torch.bfloat16
    out.backward(gout)
    got = [out.detach(), x.grad, ln.weight.grad, ln.bias.grad, conv.weight.grad, conv.bias.grad]
    # fp64 reference
    xd = x.detach().double().requires_grad_(True)
    w, b = ln.weight.detach().double().requires_grad_(True), ln.bias.detach().double().requires_grad_(True)
    cw, cb = conv.weight.detach().double().requires_grad_(True), conv.bias.detach().double().requires_grad_(True)
    u = xd.mean(1, keepdim=True)
    s = (xd - u).pow(2).mean(1, keepdim=True)
    y = w[:, None, None] * ((xd - u) / torch.sqrt(s + 1e-6)) + b[:, None, None]
    ref = F.conv2d(y, cw, cb, stride=2)
    ref.backward(gout.double())
    want = [ref.detach(), xd.grad, w.grad, b.grad, cw.grad, cb.grad]
    names = ["out", "dx", "dlnw", "dlnb", "dW", "db"]
    # bf16 operands (2^-9 per rounding) with fp32 accumulation: one rounding of A, W and Y on the way forward, of dY, W and
    # dA on the way back
    bounds = [6e-3, 1.2e-2, 1.2e-2, 1.2e-2, 8e-3, 4e-3]
    for n, g, r, bd in zip(names, got, want, bounds):
        assert _rel(g, r) < bd, f"{n}: rel L2 {_rel(g, r):.3e} (bound {bd})"
    assert _rel(out._slak_bf16, ref.detach()) < 8e-3


def test_slak_model_uses_fused_downsample():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from slak_b200 import slak
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    m = slak.SLaK(depths=[1, 1, 1, 1], dims=[16, 32, 64, 128], kernel_size=[13, 11, 9, 7, 5], Decom=True, bn=True, num_classes=10).to(dev)
    x = torch.randn(2, 3, 64, 64, device=dev)
    outs = {}
    for flag in (True, False):
        slak.FUSED_DOWNSAMPLE = flag
        m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x)
        y.float().sum().backward()
        outs[flag] = (y.detach().float(), m.downsample_layers[1][1].weight.grad.clone(), m.downsample_layers[0][0].weight.grad.clone())
    slak.FUSED_DOWNSAMPLE = True
    assert _rel(outs[True][0], outs[False][0]) < 3e-2
    assert _rel(outs[True][1], outs[False][1]) < 5e-2
    assert _rel(outs[True][2], outs[False][2]) < 5e-2


@pytest.mark.parametrize("N,C,H,W", [(2, 16, 16, 16), (3, 24, 8, 24), (2, 96, 224, 224)])
def test_fused_stem_matches_fp64(N, C, H, W):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from slak_b200 import downsample
    from slak_b200.slak import LayerNorm
    torch.manual_seed(C + H)
    dev = torch.device("cuda:0")
    conv = nn.Conv2d(3, C, kernel_size=4, stride=4).to(dev)
    ln = LayerNorm(C, eps=1e-6, data_format="channels_first").to(dev)
    with torch.no_grad():
        ln.weight.copy_(torch.randn(C) * 0.2 + 1); ln.bias.copy_(torch.randn(C) * 0.1)
        conv.weight.copy_(torch.randn_like(conv.weight) * 0.2); conv.bias.copy_(torch.randn(C) * 0.1)
    x = torch.randn(N, 3, H, W, device=dev)
    gout = torch.randn(N, C, H // 4, W // 4, device=dev)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert downsample.fused_stem_supported(conv, ln, x)
        out = downsample.fused_stem(conv, ln, x)
    out.backward(gout)
    got = [out.detach(), conv.weight.grad, conv.bias.grad, ln.weight.grad, ln.bias.grad]
    cw, cb = conv.weight.detach().double().requires_grad_(True), conv.bias.detach().double().requires_grad_(True)
    w, b = ln.weight.detach().double().requires_grad_(True), ln.bias.detach().double().requires_grad_(True)
    y = F.conv2d(x.double(), cw, cb, stride=4)
    u = y.mean(1, keepdim=True)
    s = (y - u).pow(2).mean(1, keepdim=True)
    ref = w[:, None, None] * ((y - u) / torch.sqrt(s + 1e-6)) + b[:, None, None]
    ref.backward(gout.double())
    want = [ref.detach(), cw.grad, cb.grad, w.grad, b.grad]
    # the convolution output is rounded to bf16 before the LayerNorm (as under the reference's autocast): 2^-9 on y, amplified
    # by 1/std in the normalised output where a token's channels are close together
    for n, g, r, bd in zip(["out", "dW", "db", "dlnw", "dlnb"], got, want, [1.2e-2, 2e-2, 2e-2, 1.5e-2, 4e-3]):
        assert _rel(g, r) < bd, f"{n}: rel L2 {_rel(g, r):.3e} (bound {bd})"
