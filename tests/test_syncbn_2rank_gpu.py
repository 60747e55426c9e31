"""SyncBatchNorm numerics of the fused Block with two ranks (models/SLaK.py:24-28: every branch BN is an
nn.SyncBatchNorm under DDP).  Two processes share cuda:0 and talk over gloo (NCCL refuses two ranks on one device;
the collectives carry CUDA tensors either way), each runs the fused Block on its half of a batch.  Checked:

  * outputs of the two halves == the fused Block on the full batch in one process (statistics over the global batch);
  * for every parameter, grad_rank0 + grad_rank1 == the full-batch gradient.  A data-parallel wrapper averages, so
    the per-rank BN weight / bias gradients must be LOCAL sums, as torch's SyncBatchNorm returns them
    (torch/nn/modules/_functions.py:140-160) -- global sums on every rank would come out world_size times too large;
  * running statistics after the step == the full-batch ones;
  * the same two-rank run module by module (torch.nn.SyncBatchNorm itself) agrees with the fused node;
  * unequal per-rank batches (5 + 3).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
DIM, HW, KS = 16, 28, 49
SPLITS = [(4, 4), (5, 3)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_block(sync):
    from slak_b200 import slak
    torch.manual_seed(4)
    slak.use_sync_bn = sync
    blk = slak.Block(dim=DIM, drop_path=0.0, layer_scale_init_value=1.0, kernel_size=(KS, 5), Decom=True, bn=True)
    for p in blk.parameters():
        if p.dim() > 1:
            torch.nn.init.normal_(p, std=0.05)
    for m in blk.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            torch.nn.init.uniform_(m.weight, 0.5, 1.5)
            torch.nn.init.uniform_(m.bias, -0.5, 0.5)
    return blk.cuda().train()


def _data(n):
    g = torch.Generator().manual_seed(8)
    return torch.randn(n, DIM, HW, HW, generator=g), torch.randn(n, DIM, HW, HW, generator=g)


def _run(blk, x, cot, fused):
    from slak_b200 import slak
    slak.FUSED_BLOCK = fused
    try:
        xi = x.cuda().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = blk(xi)
        (y.float() * cot.cuda()).sum().backward()
    finally:
        slak.FUSED_BLOCK = True
    grads = {n: p.grad.detach().float().cpu() for n, p in blk.named_parameters()}
    bufs = {n: b.detach().float().cpu() for n, b in blk.named_buffers()}
    return y.detach().float().cpu(), xi.grad.float().cpu(), grads, bufs


def _np(res):
    """numpy copies for the multiprocessing queue (torch tensors would travel as file descriptors that die with the
    worker process)."""
    y, dx, g, b = res
    return y.numpy(), dx.numpy(), {k: v.numpy() for k, v in g.items()}, {k: v.numpy() for k, v in b.items()}


def _pt(res):
    y, dx, g, b = res
    t = torch.from_numpy
    return t(y), t(dx), {k: t(v) for k, v in g.items()}, {k: t(v) for k, v in b.items()}


def _worker(rank, world, port, split, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = sum(split)
        lo = sum(split[:rank])
        x, cot = _data(n)
        xs, cs = x[lo:lo + split[rank]], cot[lo:lo + split[rank]]
        res = {}
        for fused in (True, False):
            blk = _make_block(sync=True)
            res[fused] = _np(_run(blk, xs, cs, fused))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("split", SPLITS)
def test_fused_block_syncbn_two_ranks_equals_full_batch(split):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    n = sum(split)
    x, cot = _data(n)
    y_full, dx_full, g_full, b_full = _run(_make_block(sync=False), x, cot, fused=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, split, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
    for fused in (True, False):
        (y0, dx0, g0, b0), (y1, dx1, g1, b1) = _pt(got[0][fused]), _pt(got[1][fused])
        tag = "fused" if fused else "module-by-module (torch.nn.SyncBatchNorm)"
        tol = 2e-2 if fused else 5e-2        # the module path rounds y_i and the BN outputs to bf16 separately
        assert rel(torch.cat([y0, y1]), y_full) < tol, (tag, rel(torch.cat([y0, y1]), y_full))
        assert rel(torch.cat([dx0, dx1]), dx_full) < 2 * tol, (tag, rel(torch.cat([dx0, dx1]), dx_full))
        for name in g_full:
            r = rel(g0[name] + g1[name], g_full[name])
            assert r < 3 * tol, (tag, name, r)
        for name in b_full:
            if "num_batches" in name:
                assert torch.equal(b0[name], b_full[name])
            else:
                assert rel(b0[name], b_full[name]) < tol and torch.equal(b0[name], b1[name]), (tag, name)
    # the BN parameter gradients are the point of the test: tight comparison fused vs torch.nn.SyncBatchNorm per rank
    for r_ in (0, 1):
        gf, gm = _pt(got[r_][True])[2], _pt(got[r_][False])[2]
        for name in gf:
            if ".bn." in name:
                assert rel(gf[name], gm[name]) < 6e-2, (r_, name, rel(gf[name], gm[name]))
