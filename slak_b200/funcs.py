"""Prune / growth / redistribution registries: host-side mirror of the reference's `funcs.py`
(registries at funcs.py:374-392).  `magnitude` prune and `random` growth -- the defaults of
main.py:211-212 -- run on the sm_100a mask kernels; the other modes keep their tensor-level
definitions (same torch primitives as the reference).
"""
from __future__ import annotations

import math

import torch

from . import _lib

# ---------------------------------------------------------------- redistribution (funcs.py:7-50)

def momentum_redistribution(masking, name, weight, mask):
    grad = masking.get_momentum_for_weight(weight)
    return torch.abs(grad[mask.bool()]).mean().item()


def magnitude_redistribution(masking, name, weight, mask):
    return torch.abs(weight)[mask.bool()].mean().item()


def nonzero_redistribution(masking, name, weight, mask):
    return (weight != 0.0).sum().item()


def no_redistribution(masking, name, weight, mask):
    return weight.numel() / float(masking.baseline_nonzero)


# ---------------------------------------------------------------- prune (funcs.py:56-126)
_prune_ws = {}


def magnitude_prune(masking, mask, weight, name):
    """Zero the mask at the k = ceil(zeros + ceil(rate*nonzeros)) smallest |w| (funcs.py:107-114)."""
    num_remove = math.ceil(masking.prune_rate * masking.name2nonzeros[name])
    num_zeros = masking.name2zeros[name]
    k = math.ceil(num_zeros + num_remove)
    if num_remove == 0.0:
        return weight.data != 0.0
    w = weight.data
    if w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and mask.is_contiguous():
        lib = _lib.load()
        dev = w.device
        ws = _prune_ws.get(dev.index)
        need = lib.slak_mask_prune_workspace(w.numel())
        if ws is None or ws.numel() < need:
            ws = torch.empty(max(need, 4096), dtype=torch.uint8, device=dev)
            _prune_ws[dev.index] = ws
        with torch.cuda.device(dev):
            rc = lib.slak_mask_prune_magnitude(w.data_ptr(), mask.data_ptr(), w.numel(), int(k), ws.data_ptr(),
                                               ws.numel(), _lib.current_stream_ptr())
        _lib.check(rc, "slak_mask_prune_magnitude")
        from . import ops
        ops._count(18)   # init + 8 x (histogram, pick) + write
        return mask
    # host-logic path for CPU tensors (unit tests of the controller without a GPU)
    _, idx = torch.sort(torch.abs(w.reshape(-1)), stable=True)
    mask.data.view(-1)[idx[:k]] = 0.0
    return mask


def magnitude_and_negativity_prune(masking, mask, weight, name):
    """'SET' pruning (funcs.py:128-157): remove the smallest positive and the largest negative weights."""
    num_remove = math.ceil(masking.prune_rate * masking.name2nonzeros[name])
    if num_remove == 0.0:
        return weight.data != 0.0
    num_zeros = masking.name2zeros[name]
    k = math.ceil(num_zeros + (num_remove / 2.0))
    x, idx = torch.sort(weight[weight > 0.0].data.view(-1))
    if x.numel():
        kk = min(math.ceil(num_remove / 2.0), x.shape[0] - 1)
        mask.data[(weight < x[kk].item()) & (weight > 0.0)] = 0.0
    x, idx = torch.sort(weight[weight < 0.0].view(-1))
    if x.numel():
        kk = min(math.ceil(num_remove / 2.0), x.shape[0] - 1)
        mask.data[(weight > x[kk].item()) & (weight < 0.0)] = 0.0
    return mask


# ---------------------------------------------------------------- growth (funcs.py:170-299)

def random_growth(masking, name, new_mask, total_regrowth, weight):
    """Bernoulli(total_regrowth / zeros) over the whole tensor, drawn with torch.rand on the CPU
    default generator exactly like `torch.rand(new_mask.shape).cuda()` (funcs.py:170-175)."""
    zeros = getattr(masking, "_zeros_after_prune", None)
    n = zeros[name] if zeros is not None and name in zeros else (new_mask == 0).sum().item()
    if n == 0:
        return new_mask
    p = total_regrowth / n
    grown = (torch.rand(new_mask.shape) < p).to(new_mask.device, non_blocking=True)
    return new_mask.bool() | grown


def gradient_growth(masking, name, new_mask, total_regrowth, weight):
    grad = masking.get_gradient_for_weights(weight)
    grad = grad * (new_mask == 0).to(grad.dtype)
    _, idx = torch.sort(torch.abs(grad).flatten(), descending=True)
    new_mask.data.view(-1)[idx[:total_regrowth]] = 1.0
    return new_mask


def momentum_growth(masking, name, new_mask, total_regrowth, weight):
    grad = masking.get_momentum_for_weight(weight)
    grad = grad * (new_mask == 0).to(grad.dtype)
    _, idx = torch.sort(torch.abs(grad).flatten(), descending=True)
    new_mask.data.view(-1)[idx[:total_regrowth]] = 1.0
    return new_mask


prune_funcs = {"magnitude": magnitude_prune, "SET": magnitude_and_negativity_prune}
growth_funcs = {"random": random_growth, "gradient": gradient_growth, "momentum": momentum_growth}
redistribution_funcs = {"momentum": momentum_redistribution, "nonzero": nonzero_redistribution,
                        "magnitude": magnitude_redistribution, "none": no_redistribution}
