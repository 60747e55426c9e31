"""Pin the oracle restatements to golden vectors produced by the REFERENCE's own classes
(oracle/gen_golden.py imports models/SLaK.py, sparse_core.py, funcs.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import dwconv as orc
from oracle import masking as omask
from oracle import slak_model as omodel

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def _sd(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}


def test_conv_grid_golden():
    z = _load("ref_conv_grid.npz")
    for seed in (0, 42):
        for k in (3, 7, 13, 31):
            x, w, y = (z[f"s{seed}_k{k}_{t}"] for t in "xwy")
            np.testing.assert_allclose(orc.fwd_c(x, w), y, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag", ["k13", "k51"])
def test_block_golden_forward_and_backward(tag):
    z = _load(f"ref_block_{tag}.npz")
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in _sd(z, "sd0.").items()}
    x = torch.from_numpy(z["x"]).requires_grad_(True)
    y = omodel.block(x, sd, "", training=True)
    np.testing.assert_allclose(y.detach().numpy(), z["y_train"], rtol=1e-4, atol=1e-5)
    (y * torch.from_numpy(z["cot"])).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), z["dx"], rtol=1e-3, atol=1e-4)
    for k in z.files:
        if k.startswith("grad."):
            np.testing.assert_allclose(sd[k[5:]].grad.numpy(), z[k], rtol=2e-3, atol=2e-4, err_msg=k)
    sd1 = _sd(z, "sd1.")
    with torch.no_grad():
        y_eval = omodel.block(torch.from_numpy(z["x"]), sd1, "", training=False)
    np.testing.assert_allclose(y_eval.numpy(), z["y_eval"], rtol=1e-4, atol=1e-5)


def test_narrow_model_golden():
    z = _load("ref_slak_narrow.npz")
    sd = _sd(z, "sd.")
    x = torch.from_numpy(z["x"])
    depths = [int(d) for d in z["depths"]]
    with torch.no_grad():
        np.testing.assert_allclose(omodel.forward(x, sd, depths, training=False).numpy(), z["logits_eval"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(omodel.forward(x, sd, depths, training=True).numpy(), z["logits_train"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("init", ["uniform", "ERK"])
@pytest.mark.parametrize("only_l", [False, True])
def test_masking_golden_sequence(init, only_l):
    """Replay the reference run: same CPU seeds, same pseudo-gradients, SGD(momentum .9);
    masks must be bit-identical, weights equal as float32 bit patterns."""
    z = _load(f"ref_masking_{init}_{'onlyL' if only_l else 'all'}.npz")
    names = [k[len("w_init."):] for k in z.files if k.startswith("w_init.")]
    weights = {n: torch.from_numpy(z["w_init." + n]).clone() for n in names}
    shapes = {n: tuple(weights[n].shape) for n in omask.maskable_names({n: w.shape for n, w in weights.items()}, only_l)}
    torch.manual_seed(123)
    masks = omask.init_uniform(shapes, 0.6) if init == "uniform" else omask.init_erk(shapes, 0.6)
    masks = omask.drop_dense(masks)
    omask.apply_mask(weights, masks)
    assert sorted(masks) == sorted(str(s) for s in z["mask_names"])
    for n in masks:
        assert np.array_equal(masks[n].numpy(), z["mask0." + n]), n
    for n in names:
        assert np.array_equal(weights[n].numpy().view(np.uint32), z["w0." + n].view(np.uint32)), n
    decay = omask.CosineDecay(0.5, 12)
    g = torch.Generator().manual_seed(99)
    mom = {}
    prune_rate = 0.5
    for step in range(1, 7):
        for n in names:                                   # torch.optim.SGD(lr=.1, momentum=.9)
            grad = torch.randn(weights[n].shape, generator=g) * 0.05
            mom[n] = grad.clone() if n not in mom else mom[n] * 0.9 + grad
            weights[n] = weights[n] - 0.1 * mom[n]
        torch.manual_seed(1000 + step)
        omask.apply_mask(weights, masks, mom)
        decay.step()
        prune_rate = decay.get_dr()
        if step % 2 == 0:
            omask.truncate_weights(weights, masks, prune_rate, mom)
        assert prune_rate == pytest.approx(float(z["prune_rates"][step - 1]), rel=0, abs=0)
        for n in masks:
            assert np.array_equal(masks[n].numpy(), z[f"mask{step}." + n]), (step, n)
        for n in names:
            np.testing.assert_allclose(weights[n].numpy(), z[f"w{step}." + n], rtol=1e-6, atol=1e-7, err_msg=f"{step} {n}")
