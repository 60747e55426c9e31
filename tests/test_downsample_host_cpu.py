"""Host logic of the stem / downsampling fast path (slak_b200/downsample.py): on CPU tensors, without autocast, or for layer
shapes the kernels do not cover, the layers take the module path (reference models/SLaK.py:226-231)."""
import torch
import torch.nn as nn

from slak_b200 import downsample, slak


def test_predicates_refuse_what_the_kernels_do_not_cover():
    ln = slak.LayerNorm(16, eps=1e-6, data_format="channels_first")
    conv = nn.Conv2d(16, 32, kernel_size=2, stride=2)
    x = torch.randn(2, 16, 8, 8)
    assert not downsample.fused_downsample_supported(ln, conv, x)              # CPU tensor, no autocast
    stem = nn.Conv2d(3, 16, kernel_size=4, stride=4)
    assert not downsample.fused_stem_supported(stem, ln, torch.randn(2, 3, 16, 16))
    # geometry checks do not depend on the device: a 3x3 conv or odd planes are never taken
    conv3 = nn.Conv2d(16, 32, kernel_size=3, stride=2, padding=1)
    assert not downsample.fused_downsample_supported(ln, conv3, x)
    assert not downsample.fused_downsample_supported(ln, conv, torch.randn(2, 16, 7, 7))


def test_no_cpu_fallback_anywhere_in_the_model():
    """The stem / downsampling layers fall back to torch modules on CPU tensors, but the depthwise operator has no CPU path:
    the model fails loudly instead of silently computing on the host (reference-side boundary)."""
    import pytest
    torch.manual_seed(0)
    m = slak.SLaK(depths=[1, 1, 1, 1], dims=[8, 16, 24, 32], kernel_size=[7, 7, 5, 5, 3], Decom=True, bn=True, num_classes=5)
    m.eval()
    x = torch.randn(2, 3, 32, 32)
    with torch.no_grad():
        stem_out = m.downsample_layers[0](x)                  # plain torch modules: runs
        assert stem_out.shape == (2, 8, 8, 8)
        with pytest.raises(RuntimeError, match="CUDA"):
            m(x)
