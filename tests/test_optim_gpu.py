"""Fused AdamW + mask apply + mask-aware EMA (slak_b200/optim.py, csrc/optim.cu) against the three reference steps it
replaces: torch.optim.AdamW (optim_factory.py:149-150), `p.data * mask` (sparse_core.py:322-333) and
ModelEma.update (model_sema.py:67-91), evaluated with stock torch ops on the same tensors."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ulps(a, b):
    ia, ib = a.view(torch.int32).long(), b.view(torch.int32).long()
    return (ia - ib).abs().max().item()


def _make(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(96, 1, 51, 5), (384, 96), (96,), (7,), (768, 3072), (13, 5, 3)]
    return [torch.nn.Parameter((torch.randn(s, generator=g) * 0.1).to(DEV)) for s in shapes]


def test_fused_adamw_matches_torch_adamw_over_steps_with_lr_schedule():
    from slak_b200.optim import FusedAdamW
    pa, pb = _make(0), _make(0)
    groups = lambda ps: [{"params": [p for p in ps if p.dim() > 1], "weight_decay": 0.05},
                         {"params": [p for p in ps if p.dim() <= 1], "weight_decay": 0.0}]
    ref = torch.optim.AdamW(groups(pa), lr=4e-3, betas=(0.9, 0.999), eps=1e-8, foreach=False, fused=False)
    ours = FusedAdamW(groups(pb), lr=4e-3, betas=(0.9, 0.999), eps=1e-8)
    worst = 0
    for step in range(6):
        g = torch.Generator().manual_seed(100 + step)
        for a, b in zip(pa, pb):
            gr = (torch.randn(a.shape, generator=g) * 0.01).to(DEV)
            a.grad, b.grad = gr.clone(), gr.clone()
        for o in (ref, ours):                        # engine.py:39-44: lr rewritten every iteration
            for grp in o.param_groups:
                grp["lr"] = 4e-3 * (1.0 - 0.1 * step)
        ref.step()
        ours.step()
        for a, b in zip(pa, pb):
            worst = max(worst, _ulps(a.data, b.data))
            assert torch.allclose(a.data, b.data, rtol=1e-6, atol=1e-9)
            assert torch.allclose(ref.state[a]["exp_avg"], ours.state[b]["exp_avg"], rtol=1e-6, atol=1e-12)
            assert torch.allclose(ref.state[a]["exp_avg_sq"], ours.state[b]["exp_avg_sq"], rtol=1e-6, atol=1e-15)
    assert int(ours.state[pb[0]]["step"].item()) == 6
    print("max difference to torch.optim.AdamW(foreach=False) after 6 steps: %d ulp" % worst)
    assert worst <= 4


def _masked_setup():
    from slak_b200 import slak
    from slak_b200.optim import FusedAdamW
    from slak_b200.sparse_core import CosineDecay, Masking
    torch.manual_seed(3)
    slak.use_sync_bn = False
    net = torch.nn.Sequential()
    net.add_module("stages", torch.nn.Sequential(slak.Block(dim=16, kernel_size=(13, 5), Decom=True, bn=True)))
    net = net.to(DEV)
    opt = FusedAdamW(net.parameters(), lr=1e-2, weight_decay=0.05)
    args = types.SimpleNamespace(device=DEV, fix=False, update_frequency=3, only_L=False, sparse_init="uniform", sparsity=0.5,
                                 distributed=False)
    mask = Masking(opt, None, CosineDecay(0.5, 100), prune_rate=0.5, prune_mode="magnitude", growth_mode="random",
                   redistribution_mode="none", args=args)
    mask.add_module(net)
    opt.attach_masking(mask)
    return net, opt, mask


def test_fused_step_applies_masks_like_optimizer_step_then_apply_mask_and_ema_like_model_sema():
    import copy
    from slak_b200.optim import ModelEma
    net, opt, mask = _masked_setup()
    ema = ModelEma(net, decay=0.9)
    opt.attach_ema(ema.ema, decay=0.9)
    ema.params_in_optimizer = True
    # reference chain on a copy: torch AdamW -> p * mask -> the model_sema.py expressions
    rnet = copy.deepcopy(net)
    ropt = torch.optim.AdamW(rnet.parameters(), lr=1e-2, weight_decay=0.05, foreach=False, fused=False)
    rema = copy.deepcopy(rnet).eval()
    rnames = dict(rnet.named_parameters())
    for step in range(7):
        g = torch.Generator().manual_seed(500 + step)
        for (n, p), (_, q) in zip(net.named_parameters(), rnet.named_parameters()):
            gr = (torch.randn(p.shape, generator=g) * 0.05).to(DEV)
            p.grad, q.grad = gr.clone(), gr.clone()
        old_masks = {n: m.clone() for n, m in mask.masks.items()}
        mask.step()                                   # fused: AdamW + masks + EMA of the parameters in one launch
        ema.update(net, mask)                         # buffers only
        ropt.step()
        with torch.no_grad():
            for n, q in rnames.items():               # Masking.step(): optimizer.step(); apply_mask() with the masks of this step
                if n in old_masks:
                    q.data = q.data * old_masks[n]
            if mask.steps % 3 == 0:                   # ... then prune + grow and apply_mask() with the new masks
                for n, q in rnames.items():
                    if n in mask.masks:
                        q.data = q.data * mask.masks[n]
            msd = rnet.state_dict()
            for k, ev in rema.state_dict().items():
                mv = msd[k].detach()
                if k in mask.masks:
                    m = mask.masks[k]
                    diff = ((ev.data != 0).byte() ^ m.data.byte()) & m.data.byte()
                    ev.data.copy_((ev.data * 0.9 + mv * (1 - 0.9)).mul_(m.data).add_(diff * 0.9 * mv))
                else:
                    ev.copy_(ev * 0.9 + (1.0 - 0.9) * mv)
        for (n, p), (_, q) in zip(net.named_parameters(), rnet.named_parameters()):
            assert torch.allclose(p.data, q.data, rtol=2e-6, atol=1e-9), n
            if n in mask.masks:
                z = mask.masks[n] == 0
                assert torch.equal(p.data[z].abs(), torch.zeros_like(p.data[z])), n       # pruned weights are (+-)0
                assert torch.equal(torch.signbit(p.data[z]), torch.signbit(q.data[z])), n  # IEEE sign of w * 0 kept
        if mask.steps % 3 == 0:
            # prune-and-grow ranks by |w|: identical decisions need identical weights; realign the copy exactly
            with torch.no_grad():
                for (n, p), (_, q) in zip(net.named_parameters(), rnet.named_parameters()):
                    q.data.copy_(p.data)
                    for k in ("exp_avg", "exp_avg_sq"):
                        ropt.state[q][k].copy_(opt.state[p][k])
        for (k, a), (_, b) in zip(ema.ema.state_dict().items(), rema.state_dict().items()):
            if a.dtype.is_floating_point:
                assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), k
            else:
                assert torch.equal(a, b), k
    assert mask.steps == 7


def test_fused_step_replays_inside_a_cuda_graph():
    from slak_b200.optim import FusedAdamW
    pa, pb = _make(1), _make(1)
    ref = torch.optim.AdamW(pa, lr=1e-3, weight_decay=0.05, foreach=False, fused=False)
    ours = FusedAdamW(pb, lr=1e-3, weight_decay=0.05)
    grads = [torch.randn_like(p) * 0.01 for p in pb]
    for a, b, gr in zip(pa, pb, grads):
        a.grad, b.grad = gr.clone(), gr                 # static gradient buffers for the graph
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ours.step()                                     # warm-up (builds the tables, uploads lr / wd)
    torch.cuda.current_stream().wait_stream(s)
    ref.step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ours.step()
    ref.step()                                          # the capture does not execute: replay = step 2
    graph.replay()
    for _ in range(3):
        ref.step()
        graph.replay()
    torch.cuda.synchronize()
    for a, b in zip(pa, pb):
        assert torch.allclose(a.data, b.data, rtol=2e-6, atol=1e-9)
    assert int(ours.state[pb[0]]["step"].item()) == 5   # the device step counter advanced with every replay
