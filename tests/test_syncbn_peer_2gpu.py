"""SyncBatchNorm of the fused Block with the statistics exchange fused into the finalize kernels (one-shot all-reduce over
NVLink peer memory, slak_b200/syncbn.py + csrc/block_fused.cu) on TWO GPUs with NCCL: outputs, every gradient and the
running statistics against the single-process full-batch run, and against the NCCL all-reduce path of the same node.
Needs >= 2 CUDA devices (run with `gpurun --gpus 2`); skipped otherwise."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from test_syncbn_2rank_gpu import DIM, HW, _data, _free_port, _make_block, _np, _pt, _run

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, split, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from slak_b200 import syncbn
        n = sum(split)
        lo = sum(split[:rank])
        x, cot = _data(n)
        xs, cs = x[lo:lo + split[rank]], cot[lo:lo + split[rank]]
        res = {}
        for mode in ("peer", "nccl"):
            syncbn._cache.clear()
            os.environ["SLAK_SYNCBN_NCCL"] = "1" if mode == "nccl" else "0"
            blk = _make_block(sync=True)
            used_peer = syncbn.get(None, torch.device("cuda", rank)) is not None
            outs = []
            for _ in range(3):                       # several steps: the per-site epochs advance
                for p in blk.parameters():
                    p.grad = None
                outs.append(_run(blk, xs, cs, True))
            res[mode] = (_np(outs[0]), _np(outs[-1]), used_peer)
        q.put((rank, res))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("split", [(4, 4), (5, 3)])
def test_fused_block_syncbn_peer_memory_exchange_two_gpus(split):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    n = sum(split)
    x, cot = _data(n)
    y_full, dx_full, g_full, b_full = _run(_make_block(sync=False), x, cot, fused=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, split, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
    assert got[0]["peer"][2] and got[1]["peer"][2], "the peer-memory exchange was not used"
    assert not got[0]["nccl"][2]
    for mode in ("peer", "nccl"):
        (y0, dx0, g0, b0), (y1, dx1, g1, b1) = _pt(got[0][mode][0]), _pt(got[1][mode][0])
        assert rel(torch.cat([y0, y1]), y_full) < 2e-2, mode
        assert rel(torch.cat([dx0, dx1]), dx_full) < 4e-2, mode
        for name in g_full:
            assert rel(g0[name] + g1[name], g_full[name]) < 6e-2, (mode, name)
        for name in b_full:
            if "num_batches" not in name:
                assert rel(b0[name], b_full[name]) < 2e-2 and torch.equal(b0[name], b1[name]), (mode, name)
    # the two exchange paths add the same per-rank sums: results agree to rounding of the double sums
    for r_ in (0, 1):
        a, b = _pt(got[r_]["peer"][0]), _pt(got[r_]["nccl"][0])
        assert rel(a[0], b[0]) < 1e-3 and rel(a[1], b[1]) < 2e-3
        for name in a[2]:
            assert rel(a[2][name], b[2][name]) < 5e-3, name
