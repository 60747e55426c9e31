"""Data parallelism for the hot path = gradient all-reduce only (the reference wraps the model in
torch.nn.parallel.DistributedDataParallel, main.py:374-376, for exactly this effect).

B200-first restatement of what that wrapper does, sized for one NVSwitch box:
  * every gradient ends up in ONE flat fp32 buffer: autograd hands each freshly computed gradient over (p.grad is None
    during backward, so there is no `grad += new` kernel per parameter) and one multi-tensor copy per bucket gathers
    them; the optimizer reads the reduced values in place through p.grad = views of the buffer;
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
    produces them): static addresses (CUDA-graph capture, pointer tables of the fused optimizer).

    A step is `zero_grad(); [arm(); loss.backward()] x micro-steps; finish()`.  zero_grad() DROPS the .grad references
    instead of zeroing memory: autograd then takes over each freshly computed gradient (no `grad += new` kernel per
    parameter -- 308 tiny launches per SLaK-T step) and finish() gathers them into the flat buffer with one
    multi-tensor copy, leaving p.grad = the flat views for the optimizer.  Without zero_grad() the views stay in
    place and autograd accumulates into them (plain torch semantics)."""

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
        self.views = {}
        for p, o in zip(self.order, self.offs):
            self.views[p] = self.flat[o:o + p.numel()].view_as(p)
            p.grad = self.views[p]

    def zero_grad(self) -> None:
        """Start a step (instead of optimizer.zero_grad()): the next backward's gradients are taken over as they are
        produced and gathered by finish(); nothing is written here."""
        for p in self.params:
            p.grad = None

    def arm(self, last_micro_step: bool = True) -> None:
        """Call before every backward (no-op without data parallelism)."""

    @torch.no_grad()
    def _gather(self, params) -> None:
        """Copy the gradients autograd left in p.grad into their slices of the flat buffer (one multi-tensor launch);
        a parameter without a gradient this step gets zeros."""
        src, dst = [], []
        for p in params:
            v, g = self.views[p], p.grad
            if g is None:
                v.zero_()
            elif g.data_ptr() != v.data_ptr():
                src.append(g if g.dtype == v.dtype else g.to(v.dtype))
                dst.append(v)
        if dst:
            torch._foreach_copy_(dst, src)

    def finish(self) -> None:
        """After the last backward of the step: everything into the flat buffer, p.grad = its views."""
        self._gather(self.params)
        for p in self.params:
            p.grad = self.views[p]


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
        self.bucket_params = []                         # parameters of each bucket
        self._bucket_of = {}
        start, members = 0, []
        for p, o in zip(self.order, self.offs):
            self._bucket_of[p] = len(self.buckets)
            members.append(p)
            end = o + (p.numel() + 3) // 4 * 4
            if end - start >= cap:
                self.buckets.append([start, end, len(members)])
                self.bucket_params.append(members)
                start, members = end, []
        if members:
            self.buckets.append([start, total, len(members)])
            self.bucket_params.append(members)
        self._pending = [b[2] for b in self.buckets]
        self._gathered = [False] * len(self.buckets)
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
        self._gathered = [False] * len(self.buckets)
        if self._armed and self.cuda:
            self.side.wait_stream(torch.cuda.current_stream())

    def _reduce(self, bi: int) -> None:
        s, e, _ = self.buckets[bi]
        self._gather(self.bucket_params[bi])             # the bucket's gradients into the flat buffer (current stream)
        self._gathered[bi] = True
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
        for bi, done in enumerate(self._gathered):       # (not armed: accumulation only, nothing was reduced)
            if not done:
                self._gather(self.bucket_params[bi])
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.side)
        for p in self.params:
            p.grad = self.views[p]
        self._armed = False

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []
