"""Fused AdamW + mask apply + mask-aware EMA for the sparse training step.

The reference's step is `optimizer.step()` (torch.optim.AdamW from optim_factory.py:149-150), then
`Masking.apply_mask()` -- one `p.data * mask` kernel per masked tensor (sparse_core.py:322-333) -- then
`ModelEma.update(model, mask)` walking the whole state_dict in Python (model_sema.py:67-91).  All three are
memory-bound elementwise passes over the same parameters; `FusedAdamW` does them in ONE launch of
`slak_adamw_mask_ema_step` (csrc/optim.cu), each element read and written once, with the step counter on the
device so the launch can sit inside a CUDA graph.

Surface: a torch.optim.Optimizer with torch.optim.AdamW's constructor (`params` or param groups, lr, betas, eps,
weight_decay; per-group `lr` / `weight_decay` may be changed between steps as engine.py:39-44 does), per-parameter
state `exp_avg`, `exp_avg_sq`, `step` (read by Masking.get_momentum_for_weight, sparse_core.py:359-366), plus
  .attach_masking(masking)   fold `p *= mask` of a slak_b200.sparse_core.Masking into the step
  .attach_ema(ema_module, decay)   fold ModelEma.update for the PARAMETERS into the step (buffers: ModelEma.update)
There is no CPU path: parameters must be fp32 CUDA tensors.
"""
from __future__ import annotations

import torch

from . import _lib

CHUNK = 8192


def _dev_ptr_table(ptrs, device):
    return torch.tensor(ptrs, dtype=torch.int64).to(device)


class _Tables:
    """Device pointer / size tables over a list of (p, g, m, v, mask, ema) tensors and the flat chunk list."""

    def __init__(self, ps, gs, ms, vs, masks, emas, device):
        self.keep = (ps, gs, ms, vs, masks, emas)
        ptr = lambda ts: None if ts is None else _dev_ptr_table([0 if t is None else t.data_ptr() for t in ts], device)
        self.p, self.g, self.m, self.v = ptr(ps), ptr(gs), ptr(ms), ptr(vs)
        self.mask = ptr(masks) if masks is not None and any(t is not None for t in masks) else None
        self.ema = ptr(emas) if emas is not None and any(t is not None for t in emas) else None
        numels = [t.numel() for t in ps]
        self.numel = torch.tensor(numels, dtype=torch.int64).to(device)
        ct, co = [], []
        for i, n in enumerate(numels):
            for off in range(0, n, CHUNK):
                ct.append(i)
                co.append(off)
        self.nchunks = len(ct)
        self.chunk_tensor = torch.tensor(ct, dtype=torch.int32).to(device)
        self.chunk_off = torch.tensor(co, dtype=torch.int64).to(device)
        self.sig = tuple(0 if t is None else t.data_ptr() for ts in self.keep if ts is not None for t in ts)


def _p(t):
    return None if t is None else t.data_ptr()


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        b = {tuple(g["betas"]) for g in self.param_groups}
        e = {float(g["eps"]) for g in self.param_groups}
        if len(b) != 1 or len(e) != 1:
            raise ValueError("FusedAdamW needs the same betas and eps in every param group")
        self._tables = None
        self._hyper = None            # (lr, wd) per tensor as last uploaded
        self._lr_dev = self._wd_dev = None
        self._masking = None
        self._ema = None              # (module, decay)
        self._step_dev = None
        self.fused_mask = False       # read by slak_b200.sparse_core.Masking.step: apply_mask is part of step()

    # ------------------------------------------------------------------ attachments
    def attach_masking(self, masking) -> None:
        self._masking = masking
        self.fused_mask = masking is not None
        self._tables = None

    def attach_ema(self, ema_module, decay=0.9999) -> None:
        self._ema = None if ema_module is None else (ema_module, float(decay))
        self._tables = None

    # ------------------------------------------------------------------ tables
    def _params(self):
        return [(g, p) for g in self.param_groups for p in g["params"] if p.grad is not None]

    def _build(self, plist):
        dev = plist[0][1].device
        if self._step_dev is None:
            self._step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        name_of = {}
        if self._masking is not None:
            for module in self._masking.modules:
                for n, t in module.named_parameters():
                    name_of[t] = n
        ema_of = {}
        if self._ema is not None:
            model_named = {}
            if self._masking is not None:
                model_named = {n: t for module in self._masking.modules for n, t in module.named_parameters()}
            ema_named = dict(self._ema[0].named_parameters())
            if not model_named:
                raise RuntimeError("attach_ema needs attach_masking(masking) first, or use set_model(model) to give the names")
            for n, t in model_named.items():
                k = n[7:] if n.startswith("module.") and n not in ema_named else n
                if k in ema_named:
                    ema_of[t] = ema_named[k].data
        ps, gs, ms, vs, masks, emas = [], [], [], [], [], []
        for g, p in plist:
            if p.dtype != torch.float32 or not p.is_cuda or not p.data.is_contiguous():
                raise RuntimeError("FusedAdamW: parameters must be contiguous fp32 CUDA tensors (no CPU path)")
            st = self.state[p]
            if "exp_avg" not in st:
                st["step"] = self._step_dev            # shared device counter (all parameters step together)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            ps.append(p.data); gs.append(grad); ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"])
            mk = None
            if self._masking is not None and p in name_of and name_of[p] in self._masking.masks:
                mk = self._masking.masks[name_of[p]]
            masks.append(mk)
            emas.append(ema_of.get(p))
        return _Tables(ps, gs, ms, vs, masks, emas, dev)

    @torch.no_grad()
    def remask_ema(self) -> None:
        """After a prune-and-grow (Masking.truncate_weights) the attached EMA takes the new masks at once, as
        ModelEma.update(model, mask) -- which the reference runs after mask.step() -- would (model_sema.py:83-88)."""
        if self._ema is None or self._tables is None:
            return
        _, _, _, _, masks, emas = self._tables.keep
        for m, e in zip(masks or [], emas or []):
            if m is not None and e is not None:
                e.mul_(m)

    def set_model(self, model) -> None:
        """Names for attach_ema without a Masking: wraps the model in a minimal name provider."""
        class _Names:
            modules = [model]
            masks = {}
        if self._masking is None:
            self._masking = _Names()
            self._tables = None

    def sync_hyperparams(self) -> None:
        """Upload per-tensor lr / weight_decay if the param groups changed (engine.py:39-44 rewrites them every
        iteration).  Called by step(); call it yourself between CUDA-graph replays after changing a group."""
        plist = self._params()
        hyper = tuple((float(g["lr"]), float(g["weight_decay"])) for g, _ in plist)
        if hyper != self._hyper:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("lr / weight_decay changed during CUDA-graph capture: call sync_hyperparams() outside")
            dev = plist[0][1].device
            self._lr_dev = torch.tensor([h[0] for h in hyper], dtype=torch.float64).to(dev)
            self._wd_dev = torch.tensor([h[1] for h in hyper], dtype=torch.float64).to(dev)
            self._hyper = hyper

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        plist = self._params()
        if not plist:
            return loss
        sig = tuple(p.data.data_ptr() for _, p in plist) + tuple(p.grad.data_ptr() for _, p in plist)
        if self._tables is None or getattr(self._tables, "psig", None) != sig:
            self._tables = self._build(plist)
            self._tables.psig = sig
            self._hyper = None
        self.sync_hyperparams()
        t = self._tables
        g0 = self.param_groups[0]
        lib = _lib.load()
        with torch.cuda.device(plist[0][1].device):
            rc = lib.slak_adamw_mask_ema_step(_p(t.p), _p(t.g), _p(t.m), _p(t.v), _p(t.mask), _p(t.ema), _p(t.numel),
                                              _p(self._lr_dev), _p(self._wd_dev), _p(t.chunk_tensor), _p(t.chunk_off),
                                              t.nchunks, CHUNK, float(g0["betas"][0]), float(g0["betas"][1]),
                                              float(g0["eps"]), self._ema[1] if self._ema else 0.0, _p(self._step_dev),
                                              1, _lib.current_stream_ptr())
        _lib.check(rc, "slak_adamw_mask_ema_step")
        from . import ops
        ops._count(2)
        return loss


class ModelEma:
    """model_sema.py:36-91 surface (`.ema`, `.decay`, `.update(model, mask)`): mask-aware exponential moving average of
    the whole state_dict.  Floating-point entries are updated by ONE launch of the fused kernel in EMA-only mode;
    when the optimizer is a FusedAdamW with this EMA attached, the parameters are already done inside optimizer.step()
    and update() only handles the buffers (BatchNorm running statistics, counters)."""

    def __init__(self, model, decay=0.9999, device="", resume=""):
        from copy import deepcopy
        self.ema = deepcopy(model)
        self.ema.eval()
        self.decay = decay
        self.device = device
        if device:
            self.ema.to(device=device)
        self.ema_has_module = hasattr(self.ema, "module")
        for p in self.ema.parameters():
            p.requires_grad_(False)
        self.params_in_optimizer = False
        self._tables = None
        self._zero = None

    @torch.no_grad()
    def update(self, model, mask=None):
        needs_module = hasattr(model, "module") and not self.ema_has_module
        msd = model.state_dict()
        esd = self.ema.state_dict()
        param_keys = {n for n, _ in self.ema.named_parameters()}
        fl_model, fl_ema, fl_mask = [], [], []
        for k, ema_v in esd.items():
            mk = "module." + k if needs_module else k
            model_v = msd[mk].detach()
            if self.params_in_optimizer and k in param_keys:
                continue
            fused_ok = (ema_v.is_cuda and model_v.is_cuda and ema_v.dtype == torch.float32 and model_v.dtype == torch.float32
                        and ema_v.is_contiguous() and model_v.is_contiguous() and ema_v.numel() > 0)
            m = mask.masks.get(mk) if mask else None
            if fused_ok:
                fl_model.append(model_v); fl_ema.append(ema_v); fl_mask.append(m)
            elif m is not None:       # the reference's expressions (model_sema.py:83-89)
                diff = ((ema_v.data != 0).byte() ^ m.data.byte()) & m.data.byte()
                ema_v.data.copy_((ema_v.data * self.decay + model_v * (1 - self.decay)).mul_(m.data).add_(diff * self.decay * model_v))
            else:
                ema_v.copy_(ema_v * self.decay + (1.0 - self.decay) * model_v)
        if fl_model:
            sig = tuple(t.data_ptr() for t in fl_model + fl_ema) + tuple(0 if t is None else t.data_ptr() for t in fl_mask)
            if self._tables is None or self._tables.psig != sig:
                dev = fl_model[0].device
                self._tables = _Tables(fl_model, None, None, None, fl_mask, fl_ema, dev)
                self._tables.psig = sig
                self._zero = torch.zeros(1, dtype=torch.int64, device=dev)
            t = self._tables
            lib = _lib.load()
            with torch.cuda.device(fl_model[0].device):
                # do_adam = 0: p (the model weight) is only read; a mask in the table would also be applied to it, which
                # is idempotent here (Masking.apply_mask has already run)
                rc = lib.slak_adamw_mask_ema_step(_p(t.p), None, None, None, _p(t.mask), _p(t.ema), _p(t.numel), None, None,
                                                  _p(t.chunk_tensor), _p(t.chunk_off), t.nchunks, CHUNK, 0.9, 0.999, 1e-8,
                                                  float(self.decay), _p(self._zero), 0, _lib.current_stream_ptr())
            _lib.check(rc, "slak_adamw_mask_ema_step")
            from . import ops
            ops._count(1)
