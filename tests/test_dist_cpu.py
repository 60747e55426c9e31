"""Multi-rank host logic on CPU with the gloo backend (world_size 2): mask synchronisation from rank 0
(sparse_core.py:404-407 semantics) and the flattened gradient all-reduce bench.py uses for data parallelism."""
import os
import socket
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from slak_b200 import slak
        from slak_b200.sparse_core import CosineDecay, Masking
        torch.manual_seed(100 + rank)                 # main.py:232: seed + rank -> different masks before the sync
        slak.use_sync_bn = False
        net = torch.nn.Sequential()
        net.add_module("stages", torch.nn.Sequential(slak.Block(dim=8, kernel_size=(9, 5), Decom=True, bn=True)))
        with torch.no_grad():
            g = torch.Generator().manual_seed(7)      # identical weights on both ranks (as after DDP's broadcast)
            for p in net.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
        opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9)
        args = types.SimpleNamespace(device="cpu", fix=False, update_frequency=2, only_L=False, sparse_init="uniform",
                                     sparsity=0.4, distributed=True)
        mask = Masking(opt, None, CosineDecay(0.5, 10), prune_rate=0.5, prune_mode="magnitude", growth_mode="random",
                       redistribution_mode="none", args=args)
        mask.add_module(net)
        sig0 = torch.cat([m.flatten() for m in mask.masks.values()]).clone()
        for step in range(4):
            gg = torch.Generator().manual_seed(50 + step)
            for p in net.parameters():
                p.grad = torch.randn(p.shape, generator=gg) * 0.05
            mask.step()                               # includes prune-and-grow every 2 steps, rank-dependent RNG
        sig1 = torch.cat([m.flatten() for m in mask.masks.values()]).clone()
        w = torch.cat([p.detach().flatten() for p in net.parameters()])
        # flattened gradient all-reduce (bench.py:allreduce_grads)
        grads = [torch.full((3, 2), float(rank + 1)), torch.full((5,), 10.0 * (rank + 1))]
        flat = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(flat)
        flat.div_(world)
        for g_, f_ in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
            g_.copy_(f_)
        q.put((rank, sig0.numpy(), sig1.numpy(), w.numpy(), [g_.numpy() for g_ in grads]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_masks_follow_rank0_and_grads_average_under_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=100) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    import numpy as np
    (_, a0, a1, wa, ga), (_, b0, b1, wb, gb) = out
    assert np.array_equal(a0, b0) and np.array_equal(a1, b1)  # rank 1 holds rank 0's masks after init and after prune/grow
    assert 0 < a1.sum() < a1.size
    assert np.array_equal(wa, wb)                             # hence identical masked weights
    assert np.allclose(ga[0], 1.5) and np.allclose(gb[1], 15.0)
