"""Shared by the CPU and GPU masking tests: replay the golden run recorded from the REFERENCE's
sparse_core.Masking (oracle/gen_golden.py:gen_masking) through slak_b200.sparse_core.Masking."""
import os
import types

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def replay(init, only_l, device, prune_mode="magnitude", growth_mode="random"):
    from slak_b200 import slak
    from slak_b200.sparse_core import CosineDecay, Masking
    tag = f"{init}_{'onlyL' if only_l else 'all'}"
    if growth_mode != "random":
        tag += "_" + growth_mode
    z = np.load(os.path.join(GOLD, f"ref_masking_{tag}.npz"))
    slak.use_sync_bn = False
    net = torch.nn.Sequential()
    net.add_module("stages", torch.nn.Sequential(
        slak.Block(dim=8, kernel_size=(13, 5), Decom=True, bn=True, layer_scale_init_value=1.0),
        slak.Block(dim=8, kernel_size=(9, 5), Decom=True, bn=True, layer_scale_init_value=1.0)))
    with torch.no_grad():
        for n, p in net.named_parameters():
            p.copy_(torch.from_numpy(z["w_init." + n]))
    net.to(device)
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9)
    args = types.SimpleNamespace(device=str(device), fix=False, update_frequency=2, only_L=only_l,
                                 sparse_init=init, sparsity=0.4, distributed=False)
    loader = None
    if init == "snip":      # the batch SNIP() looked at in the reference run
        loader = [(torch.from_numpy(z["snip_images"]), torch.from_numpy(z["snip_labels"]))]
    mask = Masking(opt, train_loader=loader, prune_rate_decay=CosineDecay(0.5, 12), prune_rate=0.5,
                   prune_mode=prune_mode, growth_mode=growth_mode, redistribution_mode="none", args=args)
    torch.manual_seed(123)
    if init == "snip" and torch.device(device).type == "cpu":
        # SNIP() runs one forward/backward of the network.  The product has no CPU depthwise operator (by design), so
        # on CPU tensors this test lends the model the oracle's conv for that one call; everything Masking itself does
        # (scores, threshold, layer sparsities, Bernoulli draw) is the code under test
        from oracle import slak_model as omodel
        from slak_b200 import ops
        saved = ops.depthwise_conv2d
        ops.depthwise_conv2d = omodel.dwconv
        try:
            mask.add_module(net)
        finally:
            ops.depthwise_conv2d = saved
    else:
        mask.add_module(net)
    assert sorted(mask.masks) == sorted(str(s) for s in z["mask_names"])

    def check(step):
        for n, m in mask.masks.items():
            assert np.array_equal(m.cpu().numpy(), z[f"mask{step}." + n]), (step, n)
        for n, p in net.named_parameters():
            got, want = p.detach().cpu().numpy(), z[f"w{step}." + n]
            if step == 0:
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (step, n)   # incl. -0.0
            else:
                np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7, err_msg=f"{step} {n}")
            # a pruned position holds an exact (signed) zero
            if n in mask.masks:
                assert np.all(got[z[f"mask{step}." + n] == 0] == 0)

    check(0)
    g = torch.Generator().manual_seed(99)
    for step in range(1, 7):
        for p in net.parameters():
            p.grad = (torch.randn(p.shape, generator=g) * 0.05).to(device)
        torch.manual_seed(1000 + step)
        mask.step()
        assert mask.prune_rate == float(z["prune_rates"][step - 1])
        check(step)
        for n, p in net.named_parameters():
            k = f"mom{step}." + n
            if k in z.files:
                np.testing.assert_allclose(opt.state[p]["momentum_buffer"].cpu().numpy(), z[k], rtol=1e-6, atol=1e-7)
    return mask
