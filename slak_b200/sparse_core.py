"""Dynamic-sparse-training controller: host-side mirror of the reference's `sparse_core.py`
(`Masking`, `CosineDecay`, `SNIP`) with the per-step work moved into single CUDA launches.

Same public surface as the reference (sparse_core.py:67-407): constructor kwargs
(main.py:421-425), `.add_module(model)`, `.step()`, `.apply_mask()`, `.truncate_weights()`,
`.masks` (name -> fp32 0/1 tensor, read by model_sema.py:83-88), `.prune_rate`, `.steps`,
`.name2nonzeros/.name2zeros/.name2removed`, `.print_nonzero_counts()`.

What is different underneath (results are bit-identical, tests/test_masking_gpu.py):
  * apply_mask (:316-333): ONE kernel over all masked tensors (slak_mask_apply) instead of one
    multiply per tensor per step; `w*mask` keeps IEEE semantics (pruned negatives become -0.0).
  * synchronism_masks (:404-407): the reference broadcasts every fp32 mask on every step; masks
    only change in init()/truncate_weights(), so they are broadcast there, once, as one flat
    uint8 buffer, and every rank ends with rank 0's masks exactly as before.
  * magnitude_prune (funcs.py:107-114): device radix select of the k-th smallest |w|
    (slak_mask_prune_magnitude) instead of a full sort; counts are read back in one batched
    transfer instead of >= 3 `.item()` syncs per layer.
  * random_growth (funcs.py:170-175): the Bernoulli draw stays `torch.rand(shape)` on the CPU
    default generator, in the reference's layer order, so a given seed gives the same masks.
"""
from __future__ import annotations

import copy
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .funcs import growth_funcs, prune_funcs, redistribution_funcs


def SNIP(net, keep_ratio, train_dataloader, device, masks, args):
    """Layer-wise sparsities from |w * dL/dw| on one batch (sparse_core.py:11-47)."""
    if getattr(args, "distributed", False):
        train_dataloader.sampler.set_epoch(0)
    images, labels = next(iter(train_dataloader))
    images = images.to(device, non_blocking=True)
    labels = labels.to(device, non_blocking=True)
    net = copy.deepcopy(net)
    net.zero_grad()
    F.cross_entropy(net(images), labels).backward()
    scores = [torch.abs(w * w.grad) for name, w in net.named_parameters() if name in masks]
    flat = torch.cat([s.flatten() for s in scores])
    keep = int(len(flat) * keep_ratio)
    if flat.is_cuda and flat.dtype == torch.float32 and 1 <= keep <= flat.numel() < (1 << 32):
        # torch.topk(flat, keep)[-1] = the keep-th largest score: device radix select instead of sorting ~30 M scores
        lib = _lib.load()
        ws = torch.empty(max(int(lib.slak_mask_prune_workspace(flat.numel())), 4096), dtype=torch.uint8, device=flat.device)
        cut = torch.empty(1, dtype=torch.float32, device=flat.device)
        with torch.cuda.device(flat.device):
            rc = lib.slak_select_kth_largest_abs(flat.data_ptr(), flat.numel(), keep, ws.data_ptr(), ws.numel(), cut.data_ptr(),
                                                 _lib.current_stream_ptr())
        _lib.check(rc, "slak_select_kth_largest_abs")
        zeros = torch.stack([(s <= cut).sum() for s in scores]).tolist()      # one device -> host transfer
        out = [z / s.numel() for z, s in zip(zeros, scores)]
    else:
        threshold, _ = torch.topk(flat, keep, sorted=True)
        cut = threshold[-1]
        out = []
        for s in scores:
            m = (s > cut).float()
            out.append(float((m == 0).sum().item() / m.numel()))
    net.zero_grad()
    return out


class CosineDecay(object):
    """Cosine schedule of the prune rate; like the reference it IS torch's CosineAnnealingLR on a
    dummy SGD (sparse_core.py:49-64), so the floating-point values are the same."""

    def __init__(self, prune_rate, T_max, eta_min=0.005, last_epoch=-1, init_step=0):
        self.sgd = torch.optim.SGD(torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(1))]), lr=prune_rate)
        self.cosine_stepper = torch.optim.lr_scheduler.CosineAnnealingLR(self.sgd, T_max, eta_min, last_epoch)
        for _ in range(init_step):
            self.cosine_stepper.step()

    def step(self):
        self.cosine_stepper.step()

    def get_dr(self, prune_rate=None):
        return self.sgd.param_groups[0]["lr"]


class Masking(object):
    def __init__(self, optimizer, train_loader=None, prune_rate_decay=None, prune_rate=0.5, prune_mode="magnitude",
                 growth_mode="random", redistribution_mode="momentum", verbose=False, fp16=False, args=False):
        if growth_mode not in ("random", "momentum", "momentum_neuron", "gradient"):
            print("Growth mode: {0} not supported!".format(growth_mode))
        self.args = args
        self.device = torch.device(args.device)
        self.growth_mode = growth_mode
        self.prune_mode = prune_mode
        self.redistribution_mode = redistribution_mode
        self.prune_rate_decay = prune_rate_decay
        self.verbose = verbose
        self.train_loader = train_loader
        self.growth_func = growth_mode
        self.prune_func = prune_mode
        self.redistribution_func = redistribution_mode
        self.global_growth = False
        self.global_prune = False
        self.masks = {}
        self.modules = []
        self.names = []
        self.optimizer = optimizer
        self.baseline_nonzero = None
        self.name2zeros = {}
        self.name2nonzeros = {}
        self.name2removed = {}
        self.prune_rate = prune_rate
        self.steps = 0
        self.half = fp16
        self.name_to_32bit = {}
        self._table = None          # cached device pointer tables for the fused apply
        if self.args.fix:
            self.args.update_frequency = None

    # ------------------------------------------------------------------ construction
    def add_module(self, module):
        self.modules.append(module)
        self.module = module
        for name, tensor in module.named_parameters():
            if tensor.dim() in (2, 4) and (not self.args.only_L or "large_kernel.LoRA" in name):
                self.names.append(name)
                self.masks[name] = torch.zeros_like(tensor, dtype=torch.float32, requires_grad=False).to(self.device)
        self.init(mode=self.args.sparse_init, density=1 - self.args.sparsity)

    def init_optimizer(self):
        sd = self.optimizer.state_dict()
        if "fp32_from_fp16" in sd:
            for (name, _), t2 in zip(self.modules[0].named_parameters(), sd["fp32_from_fp16"][0]):
                self.name_to_32bit[name] = t2
            self.half = True

    def _resolve(self, table, key, what):
        if isinstance(key, str):
            if key not in table:
                raise Exception("Unknown {0} mode: {1}; known: {2}".format(what, key, sorted(table)))
            return table[key], ("global" in key)
        return key, False

    def init_growth_prune_and_redist(self):
        self.growth_func, self.global_growth = self._resolve(growth_funcs, self.growth_func, "growth")
        self.prune_func, self.global_prune = self._resolve(prune_funcs, self.prune_func, "prune")
        self.redistribution_func, _ = self._resolve(redistribution_funcs, self.redistribution_func, "redistribution")

    def init(self, mode="snip", density=0.05, erk_power_scale=1.0):
        self.init_growth_prune_and_redist()
        self.init_optimizer()
        self.density = density
        self.baseline_nonzero = 0
        named = [(n, w) for m in self.modules for n, w in m.named_parameters() if n in self.masks]

        if mode == "uniform":
            print("initialized with uniform")
            for name, weight in named:
                self.masks[name][:] = (torch.rand(weight.shape) < density).float().to(self.device)
                self.baseline_nonzero += weight.numel() * density
        elif mode == "resume":
            print("initialized with resume")
            for name, weight in named:
                self.masks[name][:] = (weight != 0.0).float().to(self.device)
                self.baseline_nonzero += weight.numel() * density
        elif mode == "snip":
            print("initialize by snip")
            sparsities = SNIP(self.module, density, self.train_loader, self.device, self.masks, self.args)
            for sp, name in zip(sparsities, self.masks):
                self.masks[name][:] = (torch.rand(self.masks[name].shape) < (1 - sp)).float().to(self.device)
        elif mode == "ERK":
            print("initialize by fixed_ERK")
            for name, m in self.masks.items():
                self.baseline_nonzero += m.numel() * density
            dense_layers = set()
            while True:
                divisor, rhs, raw = 0, 0, {}
                for name, m in self.masks.items():
                    n_param = np.prod(m.shape)
                    if name in dense_layers:
                        rhs -= n_param * (1 - density)
                    else:
                        rhs += n_param * density
                        raw[name] = (np.sum(m.shape) / np.prod(m.shape)) ** erk_power_scale
                        divisor += raw[name] * n_param
                epsilon = rhs / divisor
                max_prob = np.max(list(raw.values()))
                if max_prob * epsilon > 1:
                    for name, p in raw.items():
                        if p == max_prob:
                            print(f"Sparsity of var:{name} had to be set to 0.")
                            dense_layers.add(name)
                else:
                    break
            total_nonzero, total_params = 0.0, 0
            for name, m in self.masks.items():
                d = 1.0 if name in dense_layers else epsilon * raw[name]
                print(f"layer: {name}, shape: {m.shape}, density: {d}")
                self.masks[name][:] = (torch.rand(m.shape) < d).float().to(self.device)
                total_nonzero += d * m.numel()
                total_params += m.numel()
            print(f"Overall sparsity {total_nonzero / total_params}")

        # layers that came out (almost) dense lose their mask (sparse_core.py:243-259)
        total, nz_total, dense = 0, 0, []
        counts = self._counts([self.masks[n] for n in self.masks])
        for (name, m), nz in zip(self.masks.items(), counts):
            total += m.numel()
            nz_total += nz
            d = nz / m.numel()
            if d >= 0.99:
                dense.append(name)
            print(f"Density of layer {name} with tensor {m.size()} is {d}")
        print("Final sparsity level of {0}: {1}".format(1 - self.density, 1 - nz_total / max(total, 1)))
        for name in dense:
            self.masks.pop(name)
            print(f"pop out layer {name}")
        self._table = None
        self._sync_masks()
        self.apply_mask()

    # ------------------------------------------------------------------ per step
    def step(self):
        self.optimizer.step()
        if not getattr(self.optimizer, "fused_mask", False):     # slak_b200.optim.FusedAdamW applies the masks in its step
            self.apply_mask()
        self.advance()

    def advance(self):
        """The host part of step() (sparse_core.py:303-313) after optimizer.step() and apply_mask(): prune-rate
        schedule, step counter, prune-and-grow every `update_frequency` steps.  Split out so that a caller who replays
        `optimizer.step(); apply_mask()` inside a CUDA graph can run the schedule around the replays."""
        self.prune_rate_decay.step()
        self.prune_rate = self.prune_rate_decay.get_dr(self.prune_rate)
        self.steps += 1
        if self.args.update_frequency is not None and self.steps % self.args.update_frequency == 0:
            print("*********************************Dynamic Sparsity********************************")
            self.truncate_weights()
            self.print_nonzero_counts()

    def _masked_params(self):
        for module in self.modules:
            for name, tensor in module.named_parameters():
                if name in self.masks:
                    yield name, tensor

    def _build_table(self):
        """Device tables (weight ptr, mask ptr, momentum ptr, numel) for the fused apply kernel."""
        ws, ms, es, ns = [], [], [], []
        keep = []
        for name, p in self._masked_params():
            if p.dtype != torch.float32 or not p.is_cuda or not p.data.is_contiguous():
                return None
            mom = self.optimizer.state.get(p, {}).get("momentum_buffer", None)
            if mom is not None and (mom.dtype != torch.float32 or not mom.is_contiguous()):
                return None
            ws.append(p.data.data_ptr()); ms.append(self.masks[name].data_ptr())
            es.append(mom.data_ptr() if mom is not None else 0); ns.append(p.numel())
            keep.append((p.data, mom))
        if not ws:
            return {"count": 0}
        dev = self.device
        i64 = lambda v: torch.tensor(v, dtype=torch.int64).to(dev)
        return {"count": len(ws), "w": i64(ws), "m": i64(ms), "e": i64(es) if any(es) else None, "n": i64(ns),
                "max": max(ns), "sig": (tuple(ws), tuple(ms), tuple(es)), "keep": keep}

    def _table_valid(self):
        t = self._table
        if t is None:
            return False
        if t["count"] == 0:
            return True
        ws, ms, es = [], [], []
        for name, p in self._masked_params():
            mom = self.optimizer.state.get(p, {}).get("momentum_buffer", None)
            ws.append(p.data.data_ptr()); ms.append(self.masks[name].data_ptr())
            es.append(mom.data_ptr() if mom is not None else 0)
        return t["sig"] == (tuple(ws), tuple(ms), tuple(es))

    def apply_mask(self):
        if self.half or self.device.type != "cuda":
            return self._apply_mask_eager()
        if not self._table_valid():
            self._table = self._build_table()
        t = self._table
        if t is None:
            return self._apply_mask_eager()
        if t["count"] == 0:
            return
        lib = _lib.load()
        with torch.cuda.device(self.device):
            rc = lib.slak_mask_apply(t["w"].data_ptr(), t["m"].data_ptr(),
                                     t["e"].data_ptr() if t["e"] is not None else None, t["n"].data_ptr(),
                                     t["count"], t["max"], _lib.current_stream_ptr())
        _lib.check(rc, "slak_mask_apply")
        from . import ops
        ops._count(1)

    def _apply_mask_eager(self):
        """The reference's own statement (sparse_core.py:322-333); used for the fp16-master-copy
        mode and for non-CUDA parameters (CPU unit tests of the host logic)."""
        for name, tensor in self._masked_params():
            if not self.half:
                tensor.data = tensor.data * self.masks[name]
                st = self.optimizer.state.get(tensor, {})
                if "momentum_buffer" in st:
                    st["momentum_buffer"] = st["momentum_buffer"] * self.masks[name]
            else:
                tensor.data = tensor.data * self.masks[name].half()
                if name in self.name_to_32bit:
                    t2 = self.name_to_32bit[name]
                    t2.data = t2.data * self.masks[name]

    # ------------------------------------------------------------------ prune and grow
    def _counts(self, tensors):
        """Number of non-zeros of each tensor, one device->host transfer for all of them."""
        if not tensors:
            return []
        return torch.stack([(t != 0).sum() for t in tensors]).tolist()

    def truncate_weights(self):
        named = list(self._masked_params())
        nz = self._counts([self.masks[n] for n, _ in named])
        for (name, weight), c in zip(named, nz):
            self.name2nonzeros[name] = float(c)
            self.name2zeros[name] = self.masks[name].numel() - float(c)
        for name, weight in named:                                        # prune
            self.masks[name][:] = self.prune_func(self, self.masks[name], weight, name)
        after = self._counts([self.masks[n] for n, _ in named])
        for (name, _), c in zip(named, after):
            self.name2removed[name] = self.name2nonzeros[name] - float(c)
        self._zeros_after_prune = {name: self.masks[name].numel() - c for (name, _), c in zip(named, after)}
        for name, weight in named:                                        # grow
            new_mask = self.masks[name].data.byte()
            new_mask = self.growth_func(self, name, new_mask, math.floor(self.name2removed[name]), weight)
            self.masks[name][:] = new_mask.float()
        self._zeros_after_prune = None
        self._sync_masks()
        self.apply_mask()
        if hasattr(self.optimizer, "remask_ema"):      # slak_b200.optim.FusedAdamW with an EMA folded into its step
            self.optimizer.remask_ema()

    # ------------------------------------------------------------------ utilities
    def get_momentum_for_weight(self, weight):
        st = self.optimizer.state[weight]
        if "exp_avg" in st:
            return st["exp_avg"] / (torch.sqrt(st["exp_avg_sq"]) + 1e-08)
        return st["momentum_buffer"]

    def get_gradient_for_weights(self, weight):
        return weight.grad.clone()

    def print_nonzero_counts(self):
        named = list(self._masked_params())
        counts = self._counts([self.masks[n] for n, _ in named])
        for (name, _), c in zip(named, counts):
            m = self.masks[name]
            print("{0}: {1}->{2}, density: {3:.3f}".format(name, self.name2nonzeros[name], c, c / float(m.numel())))
        print("Prune rate: {0}\n".format(self.prune_rate))

    def fired_masks_update(self):
        """sparse_core.py:388-402: positions that have ever been active (used by the `random_unfired` growth mode)."""
        if not hasattr(self, "fired_masks"):
            self.fired_masks = {}
        layer_fired = {}
        fired_total, total = 0.0, 0.0
        for name, _ in self._masked_params():
            prev = self.fired_masks.get(name)
            cur = self.masks[name].data.byte()
            self.fired_masks[name] = cur if prev is None else (cur | prev.data.byte())
            f = float(self.fired_masks[name].sum().item())
            fired_total += f
            total += float(self.fired_masks[name].numel())
            layer_fired[name] = f / float(self.fired_masks[name].numel())
        total_fired = fired_total / max(total, 1.0)
        print("The percentage of the total fired weights is:", total_fired)
        return layer_fired, total_fired

    # ------------------------------------------------------------------ packed masks (checkpoint / broadcast)
    def _pack(self, m):
        """fp32 0/1 mask -> int32 words, 32 elements per word (device kernel on CUDA, torch ops on CPU)."""
        n = m.numel()
        words = torch.empty((n + 31) // 32, dtype=torch.int32, device=m.device)
        if m.is_cuda and m.dtype == torch.float32 and m.is_contiguous():
            with torch.cuda.device(m.device):
                rc = _lib.load().slak_mask_pack_bits(m.data_ptr(), words.data_ptr(), n, _lib.current_stream_ptr())
            _lib.check(rc, "slak_mask_pack_bits")
            return words
        bits = torch.zeros(words.numel() * 32, dtype=torch.int64, device=m.device)
        bits[:n] = (m.reshape(-1) != 0).long()
        w = (bits.view(-1, 32) << torch.arange(32, device=m.device)).sum(1)
        return torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32)

    def _unpack_into(self, words, m):
        n = m.numel()
        words = words.to(device=m.device, dtype=torch.int32).contiguous()
        if m.is_cuda and m.dtype == torch.float32 and m.is_contiguous():
            with torch.cuda.device(m.device):
                rc = _lib.load().slak_mask_unpack_bits(words.data_ptr(), m.data_ptr(), n, _lib.current_stream_ptr())
            _lib.check(rc, "slak_mask_unpack_bits")
            return
        w = words.long() & 0xFFFFFFFF
        bits = ((w.view(-1, 1) >> torch.arange(32, device=m.device)) & 1).reshape(-1)[:n]
        m.copy_(bits.view_as(m).float())

    def state_dict(self):
        """Masks as bit masks (32x smaller than the fp32 masks) plus the schedule state.  The reference saves no masks:
        `--sparse_init resume` rebuilds them as `weight != 0` (sparse_core.py:158-172), which loses every active weight
        that happens to be exactly zero and the prune-rate schedule position."""
        return {"steps": self.steps, "prune_rate": self.prune_rate,
                "decay": self.prune_rate_decay.cosine_stepper.state_dict() if self.prune_rate_decay is not None else None,
                "masks": {n: (self._pack(m).cpu(), tuple(m.shape)) for n, m in self.masks.items()}}

    def load_state_dict(self, sd):
        self.steps = int(sd["steps"])
        self.prune_rate = sd["prune_rate"]
        if sd.get("decay") is not None and self.prune_rate_decay is not None:
            self.prune_rate_decay.cosine_stepper.load_state_dict(sd["decay"])
        for n in list(self.masks):
            if n not in sd["masks"]:
                self.masks.pop(n)                      # a layer that went dense at init time
        for n, (words, shape) in sd["masks"].items():
            if n not in self.masks or tuple(self.masks[n].shape) != tuple(shape):
                raise KeyError(f"mask {n} {tuple(shape)} does not match the attached model")
            self._unpack_into(words, self.masks[n])
        self._table = None
        self.apply_mask()

    def _sync_masks(self):
        """Every rank takes rank 0's masks (sparse_core.py:404-407): one broadcast of the bit-packed masks (the
        reference broadcasts every fp32 mask separately, on every step)."""
        if not getattr(self.args, "distributed", False):
            return
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or not self.masks:
            return
        packed = [self._pack(m) for m in self.masks.values()]
        flat = torch.cat(packed)
        dist.broadcast(flat, src=0)
        off = 0
        for m, w in zip(self.masks.values(), packed):
            self._unpack_into(flat[off:off + w.numel()], m)
            off += w.numel()

    synchronism_masks = _sync_masks
