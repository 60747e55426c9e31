"""Prune / growth / redistribution registries: host-side mirror of the reference's `funcs.py`
(registries at funcs.py:374-392).  `magnitude` prune, `random` growth (the defaults of main.py:211-212) and the
score-ranked growth modes `gradient` / `momentum` run on the sm_100a mask kernels (device radix select, no full
sort, no host round trips); the remaining modes keep their tensor-level definitions (same torch primitives as the
reference).  Modes whose reference implementation reads Masking attributes the reference never defines
(`global_magnitude`, `global_momentum_growth`: masking.tolerance / prune_threshold / growth_threshold) are registered
and raise a clear error instead of the reference's AttributeError.
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


def _layer_prune_rate(masking, name):
    # the reference reads masking.name2prune_rate[name] here, a table its Masking never fills; fall back to the one rate
    table = getattr(masking, "name2prune_rate", None)
    return table[name] if table and name in table else masking.prune_rate


def magnitude_and_negativity_prune(masking, mask, weight, name):
    """'SET' pruning exactly as funcs.py:128-157 states it: the k = ceil(zeros + remove/2) smallest |w| lose their mask,
    then the ceil(remove/2) most negative weights."""
    num_remove = math.ceil(_layer_prune_rate(masking, name) * masking.name2nonzeros[name])
    if num_remove == 0.0:
        return weight.data != 0.0
    num_zeros = masking.name2zeros[name]
    k = math.ceil(num_zeros + (num_remove / 2.0))
    _, idx = torch.sort(torch.abs(weight.data.reshape(-1)), stable=True)
    mask.data.view(-1)[idx[:k]] = 0.0
    _, idx = torch.sort(weight.data.reshape(-1), stable=True)
    mask.data.view(-1)[idx[:math.ceil(num_remove / 2.0)]] = 0.0
    return mask


def global_magnitude_prune(masking):
    raise NotImplementedError("prune mode 'global_magnitude' needs Masking.tolerance / prune_threshold / increment, which "
                              "the reference's Masking never defines (funcs.py:116-126 would raise AttributeError)")


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


def _grow_by_score(score, new_mask, total_regrowth):
    """new_mask = 1 at the `total_regrowth` positions of largest |score * (new_mask == 0)| (funcs.py:196-205, :293-299:
    sort descending, idx[:k]).  CUDA fp32: device radix select (slak_mask_grow_topk), ties to the lower index."""
    k = int(total_regrowth)
    if k <= 0:
        return new_mask
    if score.is_cuda and score.dtype == torch.float32:
        lib = _lib.load()
        dev = score.device
        m = new_mask.float().contiguous()
        sc = score.contiguous()
        ws = _prune_ws.get(dev.index)
        need = lib.slak_mask_prune_workspace(sc.numel())
        if ws is None or ws.numel() < need:
            ws = torch.empty(max(need, 4096), dtype=torch.uint8, device=dev)
            _prune_ws[dev.index] = ws
        with torch.cuda.device(dev):
            rc = lib.slak_mask_grow_topk(sc.data_ptr(), m.data_ptr(), sc.numel(), k, ws.data_ptr(), ws.numel(),
                                         _lib.current_stream_ptr())
        _lib.check(rc, "slak_mask_grow_topk")
        from . import ops
        ops._count(18)
        return m.to(new_mask.dtype) if new_mask.dtype != torch.float32 else m
    g = score * (new_mask == 0).to(score.dtype)
    _, idx = torch.sort(torch.abs(g).flatten(), descending=True, stable=True)
    new_mask.data.view(-1)[idx[:k]] = 1.0
    return new_mask


def gradient_growth(masking, name, new_mask, total_regrowth, weight):
    return _grow_by_score(masking.get_gradient_for_weights(weight), new_mask, total_regrowth)


def momentum_growth(masking, name, new_mask, total_regrowth, weight):
    return _grow_by_score(masking.get_momentum_for_weight(weight), new_mask, total_regrowth)


def random_unfired_growth(masking, name, new_mask, total_regrowth, weight):
    """funcs.py:176-194: prefer positions that were never active (masking.fired_masks, kept by fired_masks_update)."""
    n = (new_mask == 0).sum().item()
    if n == 0:
        return new_mask
    fired = masking.fired_masks[name]
    num_nonfired = (fired == 0).sum().item()
    if total_regrowth <= num_nonfired:
        idx = (fired.flatten() == 0).nonzero()
        pick = torch.randperm(len(idx))[:total_regrowth].to(idx.device)
        new_mask.data.view(-1)[idx[pick]] = 1.0
        return new_mask
    new_mask[fired == 0] = 1.0
    n = (new_mask == 0).sum().item()
    p = (total_regrowth - num_nonfired) / max(n, 1)
    grown = (torch.rand(new_mask.shape) < p).to(new_mask.device)
    return new_mask.bool() | grown


def mix_growth(masking, name, new_mask, total_regrowth, weight):
    """funcs.py:207-225: a fraction masking.mix of the regrowth by gradient, the rest at random."""
    gradient_grow = int(total_regrowth * getattr(masking, "mix", 0.0))
    random_grow = total_regrowth - gradient_grow
    new_mask = _grow_by_score(masking.get_gradient_for_weights(weight), new_mask, gradient_grow)
    n = (new_mask == 0).sum().item()
    p = random_grow / max(n, 1)
    grown = (torch.rand(new_mask.shape) < p).to(new_mask.device)
    return new_mask.bool() | grown


def momentum_neuron_growth(masking, name, new_mask, total_regrowth, weight):
    """funcs.py:301-329: regrowth shared between output neurons in proportion to their mean |momentum|."""
    grad = masking.get_momentum_for_weight(weight)
    M = torch.abs(grad)
    sum_dim = [1] if M.dim() == 2 else [1, 2, 3]
    v = M.mean(sum_dim).data
    v /= v.sum()
    slots_per_neuron = (new_mask == 0).sum(sum_dim)
    M = M * (new_mask == 0).float()
    new_mask = new_mask.bool()
    for i, fraction in enumerate(v):
        neuron_regrowth = math.floor(fraction.item() * total_regrowth)
        available = slots_per_neuron[i].item()
        y, _ = torch.sort(M[i].flatten())
        if neuron_regrowth > available:
            neuron_regrowth = available
        threshold = y[-neuron_regrowth].item() if neuron_regrowth > 0 else 0.0
        if threshold == 0.0 or neuron_regrowth < 10:
            continue
        new_mask[i] = new_mask[i] | (M[i] > threshold)
    return new_mask


def global_momentum_growth(masking, total_regrowth):
    raise NotImplementedError("growth mode 'global_momentum_growth' needs Masking.tolerance / growth_threshold, which the "
                              "reference's Masking never defines (funcs.py:332-372 would raise AttributeError)")


prune_funcs = {"magnitude": magnitude_prune, "SET": magnitude_and_negativity_prune, "global_magnitude": global_magnitude_prune}
growth_funcs = {"random": random_growth, "random_unfired": random_unfired_growth, "momentum": momentum_growth,
                "gradient": gradient_growth, "mix": mix_growth, "momentum_neuron": momentum_neuron_growth,
                "global_momentum_growth": global_momentum_growth}
redistribution_funcs = {"momentum": momentum_redistribution, "nonzero": nonzero_redistribution,
                        "magnitude": magnitude_redistribution, "none": no_redistribution}
