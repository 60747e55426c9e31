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
