"""Data parallelism for the hot path = gradient all-reduce only (the reference wraps the model in
torch.nn.parallel.DistributedDataParallel, main.py:374-376, for exactly this effect).

B200-first restatement of what that wrapper does, sized for one NVSwitch box:
  * every gradient lives in ONE flat fp32 buffer (p.grad are views into it): no flatten / copy-back passes, the
    optimizer reads the reduced values in place;
  * the buffer is cut into buckets in reverse parameter order (the order backward produces gradients); when the
    last gradient of a bucket has been accumulated, the bucket is all-reduced (average) on a SIDE stream, so the
    NCCL transfer over NVLink overlaps the rest of backward; `finish()` joins the side stream before the optimizer;
  * everything is plain stream work (events + NCCL kernels), so the whole step -- backward with its overlapped
    buckets included -- can be captured in one CUDA graph and replayed;
  * under gradient accumulation (`update_freq` micro-steps, engine.py:52-81) buckets are only reduced on the last
    micro-step (`arm(last_micro_step=True)`); the reference reduces on every micro-step (no `no_sync()`), the result
    is the same.
Bucket size is chosen for launch latency and overlap, not link count: NVSwitch gives every GPU full bandwidth to
every peer.  Works on CPU tensors with the gloo backend as well (tests/test_dist_cpu.py), without streams.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGradients:
    """All gradients of a module as views of ONE flat buffer laid out in reverse parameter order (the order backward
    produces them): static addresses (CUDA-graph capture, pointer tables of the fused optimizer), one-pass zeroing."""

    def __init__(self, module: torch.nn.Module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise ValueError("all parameters must share one device and dtype (fp32 master weights)")
        self.order = list(reversed(self.params))
        self.offs, total = [], 0
        for p in self.order:
            self.offs.append(total)
            total += (p.numel() + 3) // 4 * 4          # 16-byte aligned slices
        self.flat = torch.zeros(total, dtype=dt, device=dev)
        for p, o in zip(self.order, self.offs):
            p.grad = self.flat[o:o + p.numel()].view_as(p)

    def zero_grad(self) -> None:
        """Gradients stay views of the flat buffer: zero it in one pass (instead of optimizer.zero_grad())."""
        self.flat.zero_()


class GradientAllReducer(FlatGradients):
    def __init__(self, module: torch.nn.Module, bucket_mb: float = 25.0, process_group=None, broadcast: bool = True):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("GradientAllReducer needs an initialised torch.distributed process group")
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        if broadcast:                      # identical initial weights and buffers on every rank (DDP's constructor)
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t.data, src=dist.get_global_rank(process_group, 0) if process_group else 0,
                                   group=process_group)
        super().__init__(module)
        dev = self.flat.device
        cap = max(1, int(bucket_mb * (1 << 20) / self.flat.element_size()))
        total = self.flat.numel()
        self.buckets = []                               # [start, end, n_params]
        self._bucket_of = {}
        start, count = 0, 0
        for p, o in zip(self.order, self.offs):
            self._bucket_of[p] = len(self.buckets)
            count += 1
            end = o + (p.numel() + 3) // 4 * 4
            if end - start >= cap:
                self.buckets.append([start, end, count])
                start, count = end, 0
        if count:
            self.buckets.append([start, total, count])
        self._pending = [b[2] for b in self.buckets]
        self._armed = False
        self.cuda = dev.type == "cuda"
        self.side = torch.cuda.Stream(device=dev) if self.cuda else None
        self.reduced_buckets = 0                        # statistics: buckets reduced since construction
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]

    # ------------------------------------------------------------------------------------------------
    def arm(self, last_micro_step: bool = True) -> None:
        """Call before every backward: buckets are reduced during that backward only when it is the last micro-step."""
        self._armed = bool(last_micro_step)
        self._pending = [b[2] for b in self.buckets]
        if self._armed and self.cuda:
            self.side.wait_stream(torch.cuda.current_stream())

    def _reduce(self, bi: int) -> None:
        s, e, _ = self.buckets[bi]
        view = self.flat[s:e]
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record()                                  # the bucket's last gradient has been written on this stream
            with torch.cuda.stream(self.side):
                self.side.wait_event(ev)
                dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(view, group=self.group)
            view.div_(self.world)
        self.reduced_buckets += 1

    def _hook(self, p) -> None:
        if not self._armed:
            return
        bi = self._bucket_of[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._reduce(bi)

    def finish(self) -> None:
        """After backward: reduce whatever was not triggered (parameters without a gradient this step) and make the
        current stream wait for the side stream."""
        if self._armed:
            for bi, n in enumerate(self._pending):
                if n > 0:
                    self._pending[bi] = 0
                    self._reduce(bi)
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.side)
        self._armed = False

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []
