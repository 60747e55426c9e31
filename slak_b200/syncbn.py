"""SyncBatchNorm statistics exchange over NVLink peer memory (models/SLaK.py:24-28 makes every branch BatchNorm an
nn.SyncBatchNorm; under DDP that is 3 all_gathers per Block forward and 3 all_reduces per backward, all latency-bound).

`PeerExchange` owns one SYMMETRIC buffer per (process group, device): torch.distributed._symmetric_memory allocates it
and maps every rank's copy into every process (plumbing); the exchange itself is done by this library's kernels
(`slak_bn3_finalize_fwd_sync` / `_bwd_sync`, csrc/block_fused.cu): signal, wait, and a fixed-order sum of all ranks'
payloads read through the peer pointers, fused into the BatchNorm finalize.  One `site` per (Block, direction): a
payload slot, a row of flags and a device-side epoch counter, so the calls replay inside a CUDA graph.

If the symmetric allocation is not available (single process, > 8 ranks, non-NVLink peers, gloo) `get()` returns None and
the fused Block falls back to one NCCL all-reduce per direction (slak_b200/block.py).
"""
from __future__ import annotations

import ctypes
import os

import torch

SLOT_BYTES = 48 * 1024          # >= (6 C + 2) doubles at C = 768 (36.9 KB) / 6 C floats
FLAG_BYTES = 64                 # 16 uint32 flags per site
MAX_SITES = 160
_cache = {}


class PeerExchange:
    def __init__(self, group, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world < 2 or self.world > 8:
            raise RuntimeError("PeerExchange supports 2..8 ranks of one NVSwitch box")
        nbytes = MAX_SITES * (SLOT_BYTES + FLAG_BYTES)
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, self.group)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != self.buf.data_ptr():
            raise RuntimeError("unexpected symmetric-memory layout")
        self.ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        self.epochs = torch.zeros(MAX_SITES, dtype=torch.int32, device=device)
        self.n_sites = 0
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)             # every rank's buffer is zeroed before anybody signals

    def new_site(self) -> int:
        if self.n_sites >= MAX_SITES:
            raise RuntimeError("too many SyncBatchNorm call sites")
        self.n_sites += 1
        return self.n_sites - 1

    def slot_off(self, site):
        return site * SLOT_BYTES

    def flag_off(self, site):
        return MAX_SITES * SLOT_BYTES + site * FLAG_BYTES

    def slot(self, site, numel, dtype):
        nb = numel * torch.empty((), dtype=dtype).element_size()
        if nb > SLOT_BYTES:
            raise RuntimeError("payload larger than the exchange slot")
        o = self.slot_off(site)
        return self.buf[o:o + nb].view(dtype)

    def epoch_ptr(self, site):
        return self.epochs.data_ptr() + 4 * site


def get(group, device):
    """The PeerExchange of (group, device), or None when peer memory cannot be used (then NCCL does the exchange)."""
    if os.environ.get("SLAK_SYNCBN_NCCL", "0") == "1":
        return None
    import torch.distributed as dist
    g = group if group is not None else dist.group.WORLD
    key = (id(g), device.index)
    if key not in _cache:
        ex = None
        try:
            if dist.get_backend(g) == "nccl" and 2 <= dist.get_world_size(g) <= 8:
                ex = PeerExchange(g, device)
        except Exception as e:      # noqa: BLE001 -- any failure of the optional fast path selects the NCCL path
            if dist.get_rank(g) == 0:
                print(f"[slak_b200] SyncBN peer-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL all-reduce")
            ex = None
        # all ranks must take the same path
        ok = torch.tensor([1 if ex is not None else 0], device=device, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=g)
        _cache[key] = ex if int(ok.item()) == 1 else None
    return _cache[key]
