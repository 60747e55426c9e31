"""The Block / model mirror on CUDA against the oracle's functional restatement of models/SLaK.py
and against golden vectors from the reference's own classes."""
import os

import numpy as np
import pytest
import torch

from oracle import slak_model as omodel
from slak_b200 import slak

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"


def _load_block(tag, dim, ks):
    z = np.load(os.path.join(GOLD, f"ref_block_{tag}.npz"))
    slak.use_sync_bn = False
    blk = slak.Block(dim=dim, drop_path=0.0, layer_scale_init_value=0.5, kernel_size=ks, Decom=True, bn=True)
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd0.")}
    blk.load_state_dict(sd)          # the reference's state_dict keys load unchanged
    return z, blk.to(DEV)


@pytest.mark.parametrize("tag,dim,ks", [("k13", 8, (13, 5)), ("k51", 6, (51, 5))])
def test_block_matches_reference_golden_fp32(tag, dim, ks):
    z, blk = _load_block(tag, dim, ks)
    x = torch.from_numpy(z["x"]).to(DEV).requires_grad_(True)
    blk.train()
    y = blk(x)
    (y * torch.from_numpy(z["cot"]).to(DEV)).sum().backward()
    np.testing.assert_allclose(y.detach().cpu().numpy(), z["y_train"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), z["dx"], rtol=1e-3, atol=1e-4)
    for n, p in blk.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), z["grad." + n], rtol=2e-3, atol=2e-4, err_msg=n)
    for k in z.files:                       # running statistics after one training forward
        if k.startswith("sd1.") and "running" in k:
            np.testing.assert_allclose(blk.state_dict()[k[4:]].cpu().numpy(), z[k], rtol=1e-4, atol=1e-6, err_msg=k)
    blk.eval()
    with torch.no_grad():
        np.testing.assert_allclose(blk(x.detach()).cpu().numpy(), z["y_eval"], rtol=1e-4, atol=1e-5)


def test_narrow_model_matches_reference_golden():
    z = np.load(os.path.join(GOLD, "ref_slak_narrow.npz"))
    slak.use_sync_bn = False
    net = slak.SLaK(depths=[int(d) for d in z["depths"]], dims=[int(d) for d in z["dims"]], num_classes=10,
                    kernel_size=[17, 15, 13, 7, 5], Decom=True, bn=True, layer_scale_init_value=1.0)
    net.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")})
    net.to(DEV)
    x = torch.from_numpy(z["x"]).to(DEV)
    net.eval()
    with torch.no_grad():
        np.testing.assert_allclose(net(x).cpu().numpy(), z["logits_eval"], rtol=1e-3, atol=1e-4)
    net.train()
    np.testing.assert_allclose(net(x).detach().cpu().numpy(), z["logits_train"], rtol=1e-3, atol=1e-4)


def test_block_bf16_autocast_tensor_core_path_vs_oracle():
    """Stage-1 geometry (56x56, 51x5): under bf16 autocast the three branches run on the tcgen05
    kernels; compare output and every gradient with the fp32 oracle Block on the same weights."""
    torch.manual_seed(0)
    slak.use_sync_bn = False
    dim = 16
    blk = slak.Block(dim=dim, drop_path=0.0, layer_scale_init_value=1.0, kernel_size=(51, 5), Decom=True, bn=True)
    for p in blk.parameters():
        if p.dim() > 1:
            torch.nn.init.normal_(p, std=0.05)
    x = torch.randn(4, dim, 56, 56)
    cot = torch.randn(4, dim, 56, 56)
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k)
          for k, v in blk.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    yr = omodel.block(xr, sd, "", training=True)
    (yr * cot).sum().backward()
    blk = blk.to(DEV).train()
    xg = x.to(DEV).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = blk(xg)
    (y.float() * cot.to(DEV)).sum().backward()
    rel = lambda a, b: ((a.detach().cpu().double() - b.detach().double()).abs().max() / b.detach().double().abs().max()).item()
    assert rel(y, yr) < 3e-2, rel(y, yr)
    assert rel(xg.grad, xr.grad) < 5e-2, rel(xg.grad, xr.grad)
    for n, p in blk.named_parameters():
        r = rel(p.grad, sd[n].grad)
        assert r < 6e-2, (n, r)


@pytest.mark.parametrize("dim,hw,ks,dp", [(16, 56, 51, 0.0), (24, 28, 49, 0.3), (40, 14, 47, 0.0), (64, 7, 13, 0.0), (10, 20, 9, 0.0)])
def test_fused_block_equals_module_by_module_path(dim, hw, ks, dp):
    """FusedBlockFunction (fused BN/LN/residual kernels + tensor-core branches) against the same Block run
    module by module (nn.BatchNorm2d, F.layer_norm, ...) under the same bf16 autocast."""
    torch.manual_seed(1)
    slak.use_sync_bn = False
    blk = slak.Block(dim=dim, drop_path=dp, layer_scale_init_value=1.0, kernel_size=(ks, 5), Decom=True, bn=True)
    for p in blk.parameters():
        if p.dim() > 1:
            torch.nn.init.normal_(p, std=0.05)
    for m in blk.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            torch.nn.init.uniform_(m.weight, 0.5, 1.5)
            torch.nn.init.uniform_(m.bias, -0.5, 0.5)
    blk = blk.to(DEV).train()
    import copy
    ref = copy.deepcopy(blk)
    x = torch.randn(6, dim, hw, hw, device=DEV)
    cot = torch.randn(6, dim, hw, hw, device=DEV)
    outs = []
    for fused, m in ((True, blk), (False, ref)):
        slak.FUSED_BLOCK = fused
        try:
            torch.manual_seed(123)                      # same drop-path draw
            xi = x.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = m(xi)
            (y.float() * cot).sum().backward()
            outs.append((y.detach().float(), xi.grad.float(), {n: p.grad.float() for n, p in m.named_parameters()},
                         {n: b.clone().float() for n, b in m.named_buffers()}))
        finally:
            slak.FUSED_BLOCK = True
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
    (y0, dx0, g0, b0), (y1, dx1, g1, b1) = outs
    assert y0.dtype == torch.float32
    assert rel(y0, y1) < 2e-2, rel(y0, y1)
    assert rel(dx0, dx1) < 4e-2, rel(dx0, dx1)
    for n in g0:
        assert rel(g0[n], g1[n]) < 6e-2, (n, rel(g0[n], g1[n]))
    for n in b0:                                        # running statistics / num_batches_tracked
        assert rel(b0[n], b1[n]) < 2e-2, (n, rel(b0[n], b1[n]))


def test_fused_block_eval_mode_matches_module_path():
    torch.manual_seed(2)
    slak.use_sync_bn = False
    blk = slak.Block(dim=32, drop_path=0.2, layer_scale_init_value=1.0, kernel_size=(49, 5), Decom=True, bn=True).to(DEV)
    for m in blk.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.running_mean.uniform_(-0.2, 0.2)
            m.running_var.uniform_(0.5, 1.5)
    blk.eval()
    x = torch.randn(5, 32, 28, 28, device=DEV)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y0 = blk(x)
        slak.FUSED_BLOCK = False
        try:
            y1 = blk(x)
        finally:
            slak.FUSED_BLOCK = True
    assert ((y0 - y1).abs().max() / y1.abs().max()).item() < 2e-2


def test_colsum_is_fixed_order_column_sum():
    """slak_colsum_f32 folds per-CTA partial rows: compare with a float64 column sum, twice (bitwise repeatable)."""
    from slak_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(3)
    for rows, cols in [(592, 384), (1, 7), (37, 4608), (888, 12)]:
        part = torch.randn(rows, cols, device="cuda")
        outs = []
        for _ in range(2):
            out = torch.empty(cols, device="cuda")
            rc = lib.slak_colsum_f32(part.data_ptr(), rows, cols, out.data_ptr(), _lib.current_stream_ptr())
            _lib.check(rc, "slak_colsum_f32")
            outs.append(out.clone())
        assert torch.equal(outs[0], outs[1])
        ref = part.double().sum(0)
        assert torch.allclose(outs[0].double(), ref, rtol=1e-5, atol=1e-4 * rows ** 0.5)
