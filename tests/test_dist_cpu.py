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
        # bucketed gradient all-reduce of the package (slak_b200/ddp.py; main.py:374-376 semantics): rank-dependent
        # data, two micro-steps of accumulation, tiny buckets so that several are reduced during backward
        from slak_b200.ddp import GradientAllReducer
        torch.manual_seed(300 + rank)
        mlp = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        dp = GradientAllReducer(mlp, bucket_mb=1e-4)           # broadcasts rank 0's weights
        wts = torch.cat([p.detach().flatten() for p in mlp.parameters()]).clone()
        assert len(dp.buckets) >= 2
        xs = [torch.randn(4, 6, generator=torch.Generator().manual_seed(10 * rank + k)) for k in range(2)]
        dp.zero_grad()
        for k in range(2):
            dp.arm(last_micro_step=(k == 1))
            (mlp(xs[k]).pow(2).sum() / 2).backward()
        dp.finish()
        assert dp.reduced_buckets == len(dp.buckets)
        assert all(p.grad.data_ptr() >= dp.flat.data_ptr() for p in mlp.parameters())     # still views of the flat buffer
        grads = [p.grad.detach().clone() for p in mlp.parameters()]
        q.put((rank, sig0.numpy(), sig1.numpy(), w.numpy(), [g_.numpy() for g_ in grads], wts.numpy()))
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
    (_, a0, a1, wa, ga, pa), (_, b0, b1, wb, gb, pb) = out
    assert np.array_equal(a0, b0) and np.array_equal(a1, b1)  # rank 1 holds rank 0's masks after init and after prune/grow
    assert 0 < a1.sum() < a1.size
    assert np.array_equal(wa, wb)                             # hence identical masked weights
    assert np.array_equal(pa, pb)                             # GradientAllReducer broadcast rank 0's weights
    # every rank ends with the same (averaged) gradients, equal to the single-process mean over both ranks' data
    for x, y in zip(ga, gb):
        assert np.array_equal(x, y)
    torch.manual_seed(300)
    mlp = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    tot = 0.0
    for r in range(2):
        for k in range(2):
            tot = tot + mlp(torch.randn(4, 6, generator=torch.Generator().manual_seed(10 * r + k))).pow(2).sum() / 2
    (tot / 2).backward()
    for g_, p_ in zip(ga, mlp.parameters()):
        assert np.allclose(g_, p_.grad.numpy(), rtol=1e-5, atol=1e-6)
