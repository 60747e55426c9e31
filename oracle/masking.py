"""CPU restatement of the reference's dynamic-sparse-training controller -- TEST INFRASTRUCTURE.

Functions operate on ordered dicts name -> torch CPU tensor and follow, line by line:
  apply_mask          sparse_core.py:316-333   (w = w*mask; SGD momentum_buffer *= mask)
  magnitude_prune     funcs.py:107-114
  random_growth       funcs.py:170-175         (CPU torch.rand stream, like the reference)
  truncate_weights    sparse_core.py:335-357   (all prunes first, then all growths, then apply)
  init_uniform        sparse_core.py:146-156
  init_erk            sparse_core.py:191-241
  drop_dense          sparse_core.py:243-259   (masks with density >= 0.99 are removed)
  CosineDecay         sparse_core.py:49-64
  step                sparse_core.py:300-313
Pinned by tests/golden/ref_masking_*.npz (generated from the reference's own classes).
"""
from __future__ import annotations

import math

import numpy as np
import torch


class CosineDecay:
    def __init__(self, prune_rate, T_max, eta_min=0.005, last_epoch=-1, init_step=0):
        self.sgd = torch.optim.SGD(torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(1))]), lr=prune_rate)
        self.cosine_stepper = torch.optim.lr_scheduler.CosineAnnealingLR(self.sgd, T_max, eta_min, last_epoch)
        for _ in range(init_step):
            self.cosine_stepper.step()

    def step(self):
        self.cosine_stepper.step()

    def get_dr(self, prune_rate=None):
        return self.sgd.param_groups[0]["lr"]


def apply_mask(weights, masks, momentum=None):
    for name, w in weights.items():
        if name in masks:
            weights[name] = w * masks[name]
            if momentum is not None and name in momentum:
                momentum[name] = momentum[name] * masks[name]


def magnitude_prune(mask, weight, prune_rate, nonzeros, zeros):
    num_remove = math.ceil(prune_rate * nonzeros)
    k = math.ceil(zeros + num_remove)
    if num_remove == 0.0:
        return (weight != 0.0).float()
    _, idx = torch.sort(torch.abs(weight.reshape(-1)), stable=True)
    mask = mask.clone()
    mask.view(-1)[idx[:k]] = 0.0
    return mask


def random_growth(new_mask, total_regrowth):
    n = (new_mask == 0).sum().item()
    if n == 0:
        return new_mask
    p = total_regrowth / n
    grown = torch.rand(new_mask.shape) < p
    return new_mask.bool() | grown


def truncate_weights(weights, masks, prune_rate, momentum=None):
    removed = {}
    for name in weights:
        if name not in masks:
            continue
        m = masks[name]
        nz = m.sum().item()
        z = m.numel() - nz
        new = magnitude_prune(m, weights[name], prune_rate, nz, z)
        removed[name] = nz - new.sum().item()
        masks[name] = new
    for name in weights:
        if name not in masks:
            continue
        new = random_growth(masks[name].byte(), math.floor(removed[name]))
        masks[name] = new.float()
    apply_mask(weights, masks, momentum)


def init_uniform(shapes, density):
    return {n: (torch.rand(s) < density).float() for n, s in shapes.items()}


def init_erk(shapes, density, erk_power_scale=1.0):
    dense = set()
    while True:
        divisor, rhs, raw = 0.0, 0.0, {}
        for n, s in shapes.items():
            n_param = np.prod(s)
            if n in dense:
                rhs -= n_param * (1 - density)
            else:
                rhs += n_param * density
                raw[n] = (np.sum(s) / np.prod(s)) ** erk_power_scale
                divisor += raw[n] * n_param
        eps = rhs / divisor
        mx = np.max(list(raw.values()))
        if mx * eps > 1:
            for n, v in raw.items():
                if v == mx:
                    dense.add(n)
        else:
            break
    masks = {}
    for n, s in shapes.items():
        d = 1.0 if n in dense else eps * raw[n]
        masks[n] = (torch.rand(s) < d).float()
    return masks


def drop_dense(masks):
    for n in [n for n, m in masks.items() if (m != 0).sum().int().item() / m.numel() >= 0.99]:
        masks.pop(n)
    return masks


def maskable_names(named_shapes, only_l):
    out = []
    for n, s in named_shapes.items():
        if len(s) in (2, 4) and (not only_l or "large_kernel.LoRA" in n):
            out.append(n)
    return out
