"""One training step of the model bench.py times -- SLaK-T, kernel_size [51,49,47,13,5], Decom, BN, 224^2, the four
real channel counts 96/192/384/768, bf16 autocast with fp32 master weights -- against the oracle's fp32 restatement
of the reference network (oracle/slak_model.py, pinned by goldens generated from models/SLaK.py) on the same
weights and batch.

Tolerances.  The CUDA path rounds every activation that crosses a kernel boundary to bf16 (unit roundoff
u = 2^-9 per rounding, relative) where the oracle keeps fp32; accumulations are fp32 on both sides.  Through L
layers independent roundings add in quadrature, so a gradient that has passed through the whole network
(18 Blocks x ~10 rounded tensors each way) carries a relative L2 error of about u * sqrt(2 * 180) ~ 4e-2 in the
worst (earliest) layers and much less at the head.  The bounds below are those estimates with a factor ~2, per
parameter group, instead of one blanket figure; the loss is a mean over B*1000 logits and must agree to 1e-2.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import slak_model as omodel
from slak_b200 import ops, slak

pytestmark = pytest.mark.gpu
DEV = "cuda"
U = 2.0 ** -9


def _rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_slak_tiny_224_bf16_step_matches_fp32_oracle():
    torch.manual_seed(0)
    slak.use_sync_bn = False
    B = 32
    # layer scale 0.1 (not the 1e-6 default) so that every Block really contributes to the loss and its gradients
    net = slak.SLaK_tiny(kernel_size=[51, 49, 47, 13, 5], Decom=True, bn=True, drop_path_rate=0.0, num_classes=1000,
                         layer_scale_init_value=0.1)
    x = torch.randn(B, 3, 224, 224)
    y = torch.randint(0, 1000, (B,))
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k)
          for k, v in net.state_dict().items()}
    torch.set_num_threads(max(1, min(16, torch.get_num_threads())))
    out_ref = omodel.forward(x, sd, [3, 3, 9, 3], training=True)
    loss_ref = F.cross_entropy(out_ref, y)
    loss_ref.backward()

    net = net.to(DEV).train()
    l0 = ops.launch_count()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = net(x.to(DEV))
        loss = F.cross_entropy(out.float(), y.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    assert ops.launch_count() - l0 > 18 * 10, "the fused Block path did not run"

    assert abs(loss.item() - loss_ref.item()) <= 1e-2 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    assert _rel_l2(out.float().cpu(), out_ref.detach()) <= 40 * U, _rel_l2(out.float().cpu(), out_ref.detach())
    worst = {}
    for n, p in net.named_parameters():
        g, r = p.grad.detach().float().cpu(), sd[n].grad
        e = _rel_l2(g, r)
        # distance (in Blocks) of the parameter from the loss: head 0 ... stem 18
        if n.startswith(("head", "norm")):
            depth = 0
        elif n.startswith("stages"):
            i, j = int(n.split(".")[1]), int(n.split(".")[2])
            depth = sum([3, 3, 9, 3][i + 1:]) + ([3, 3, 9, 3][i] - j)
        else:
            i = int(n.split(".")[1])
            depth = sum([3, 3, 9, 3][i:])
        bound = 2.0 * U * (2 * 10 * (depth + 1)) ** 0.5 + 4 * U
        worst[n] = (e, bound)
    top = sorted(worst.items(), key=lambda kv: -kv[1][0] / kv[1][1])[:8]
    report = [(n, round(e, 4), round(b, 4)) for n, (e, b) in top]
    print("loss", loss.item(), loss_ref.item(), "worst grad rel-L2 / bound:", report)
    bad = [(n, e, b) for n, (e, b) in worst.items() if e > b]
    assert not bad, (len(bad), report)
