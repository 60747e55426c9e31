"""Inference re-parameterisation of the Decom layout: three BatchNorms folded and the
5 x 5 kernel merged into the 5 x K one -> two kernels + bias, one tcgen05 launch.  The reference's merge_kernel
(models/SLaK.py:102-122) covers only the non-Decom `lkb_origin` layout; the algebra here follows its fuse_bn (:49-58)."""
import copy

import pytest
import torch
import torch.nn.functional as F

from oracle import slak_model as omodel
from slak_b200 import slak


def _block(dim, K, seed=0):
    torch.manual_seed(seed)
    slak.use_sync_bn = False
    blk = slak.Block(dim=dim, drop_path=0.0, layer_scale_init_value=1.0, kernel_size=(K, 5), Decom=True, bn=True)
    for p in blk.parameters():
        if p.dim() > 1:
            torch.nn.init.normal_(p, std=0.05)
    for m in blk.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            torch.nn.init.uniform_(m.weight, 0.5, 1.5)
            torch.nn.init.uniform_(m.bias, -0.5, 0.5)
            m.running_mean.uniform_(-0.3, 0.3)
            m.running_var.uniform_(0.5, 1.5)
    return blk.eval()


def test_equivalent_decom_kernels_reproduce_the_eval_mode_layer_on_cpu():
    """Pure algebra, no GPU: conv(x, kv) + conv(x, kh) + bias == sum of the three eval-mode conv+BN branches (oracle)."""
    blk = _block(6, 13)
    lk = blk.large_kernel
    kv, kh, bias = lk.get_equivalent_decom()
    assert tuple(kv.shape) == (6, 1, 13, 5) and tuple(kh.shape) == (6, 1, 5, 13) and tuple(bias.shape) == (6,)
    x = torch.randn(3, 6, 20, 20, dtype=torch.float64)
    sd = {k: v.double() if v.dtype.is_floating_point else v for k, v in blk.state_dict().items()}
    ref = omodel.reparam_large_kernel(x, sd, "large_kernel.", training=False)
    got = (F.conv2d(x, kv.double(), None, 1, (6, 2), 1, 6) + F.conv2d(x, kh.double(), None, 1, (2, 6), 1, 6)
           + bias.double().view(1, -1, 1, 1))
    assert torch.allclose(got, ref, rtol=1e-6, atol=1e-7)
    lk2 = copy.deepcopy(lk)
    lk2.merge_kernel()
    assert not hasattr(lk2, "LoRA1") and not hasattr(lk2, "small_conv") and lk2.branches() == []
    assert sorted(k for k in lk2.state_dict()) == ["lkb_reparam_h.weight", "lkb_reparam_v.bias", "lkb_reparam_v.weight"]


@pytest.mark.gpu
@pytest.mark.parametrize("dim,hw,K", [(16, 56, 51), (24, 28, 49), (40, 14, 47), (64, 7, 13)])
def test_merged_block_eval_matches_unmerged_and_oracle_on_gpu(dim, hw, K):
    blk = _block(dim, K, seed=dim)
    x = torch.randn(5, dim, hw, hw)
    sd = {k: v.clone() for k, v in blk.state_dict().items()}
    ref = omodel.block(x, sd, "", training=False)                       # fp32 oracle, unmerged
    merged = copy.deepcopy(blk)
    merged.large_kernel.merge_kernel()
    blk, merged = blk.cuda(), merged.cuda()
    xg = x.cuda()
    from slak_b200 import ops
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        l0 = ops.launch_count()
        y_m = merged(xg)                                                # fused node, ONE depthwise launch
        launches_merged = ops.launch_count() - l0
        l0 = ops.launch_count()
        y_u = blk(xg)
        launches_unmerged = ops.launch_count() - l0
        slak.FUSED_BLOCK = False
        try:
            y_mod = merged(xg)                                          # module path of the merged layer
        finally:
            slak.FUSED_BLOCK = True
    rel = lambda a, b: ((a.float().cpu() - b).abs().max() / b.abs().max()).item()
    assert rel(y_m, ref) < 2e-2 and rel(y_u, ref) < 2e-2 and rel(y_mod, ref) < 3e-2, (rel(y_m, ref), rel(y_u, ref), rel(y_mod, ref))
    assert launches_merged < launches_unmerged
    # the layer alone: merged kernel against the sum of the three eval-mode branches
    lk_u, lk_m = blk.large_kernel, merged.large_kernel
    xb = xg.bfloat16()
    with torch.no_grad():
        a = lk_m(xb).float()
        b = sum(br(xb).float() for br in lk_u.branches())
    assert ((a - b).abs().max() / b.abs().max()).item() < 2e-2
