"""The tensor-core depthwise kernels at the sizes bench.py times (batch 128, the four SLaK-T stage geometries,
BASELINE.json configs[1]) directly against the oracle.

Depthwise channels are independent, so the oracle (fp64 F.conv2d on the bf16-rounded operands,
oracle/dwconv.py:fwd_torch / grads_torch, i.e. test_correctness.py:8-9 generalised by forward_fp32.cu:140-143) is
evaluated on a SUBSET of channels of the full-size GPU result: first / last channel, and channels in the middle where
the persistent kernels switch channels inside a CTA or split one channel over several CTAs.  The GPU kernels run at
the full shape, so the multi-split partition, the channel walk and the last partial unit are all exercised.
"""
import pytest
import torch

from oracle import dwconv as orc
from slak_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"

# N, C, H, W, KL : the four stages of SLaK-T at 224^2, per-GPU batch 128 (models/SLaK.py:181-186,264-268)
STAGES = [(128, 96, 56, 56, 51), (128, 192, 28, 28, 49), (128, 384, 14, 14, 47), (128, 768, 7, 7, 13)]
# plus SLaK-T 61x61 (config 5) at stage 1 and an odd batch (last unit partial in every class)
EXTRA = [(128, 96, 56, 56, 61), (77, 40, 14, 14, 47), (45, 24, 7, 7, 13), (51, 12, 28, 28, 49)]


def _subset(C):
    idx = sorted({0, 1, C // 3, C // 2, C // 2 + 1, C - 2, C - 1})
    return torch.tensor([i for i in idx if 0 <= i < C])


def _inputs(N, C, H, W, KL, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, C, H, W, generator=g).bfloat16()
    ws = [torch.randn(C, 1, *k, generator=g) * 0.05 for k in ((KL, 5), (5, KL), (5, 5))]
    return g, x, ws


@pytest.mark.parametrize("case", STAGES + EXTRA)
def test_forward_full_size_vs_oracle(case):
    N, C, H, W, KL = case
    _, x, ws = _inputs(N, C, H, W, KL, 11 + KL + N)
    xd = x.to(DEV)
    assert ops.lk_branches_uses_tc(xd, KL, 5)
    ys = ops.lk_branches_forward(xd, *[w.to(DEV) for w in ws])
    torch.cuda.synchronize()
    idx = _subset(C)
    xs = x[:, idx].double()
    for i, (w, y) in enumerate(zip(ws, ys)):
        ref = orc.fwd_torch(xs, orc.round_like(w[idx], torch.bfloat16).double())
        got = y[:, idx.to(DEV)].cpu().double()
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        assert err <= 2.0 ** -8 + 1e-5, (i, err)          # one bf16 rounding of an fp32-accumulated sum


@pytest.mark.parametrize("case", STAGES + EXTRA)
def test_backward_full_size_vs_oracle(case):
    N, C, H, W, KL = case
    g, x, ws = _inputs(N, C, H, W, KL, 23 + KL + N)
    dys = [torch.randn(N, C, H, W, generator=g).bfloat16() for _ in range(3)]
    xd = x.to(DEV)
    assert ops.lk_branches_bwd_uses_tc(xd, KL, 5)
    wd = [w.to(DEV) for w in ws]
    dyd = [d.to(DEV) for d in dys]
    dx = ops.lk_branches_backward_data(*dyd, *wd)
    dws = ops.lk_branches_backward_filter(xd, *dyd, KL, 5)
    dws2 = ops.lk_branches_backward_filter(xd, *dyd, KL, 5)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(dws, dws2))          # fixed-order reduction: bitwise repeatable
    idx = _subset(C)
    xs = x[:, idx].double()
    dx64 = torch.zeros_like(xs)
    for i, (w, dy) in enumerate(zip(ws, dys)):
        dxi, dwi = orc.grads_torch(xs, orc.round_like(w[idx], torch.bfloat16).double(), dy[:, idx].double())
        dx64 += dxi
        got = dws[i][idx.to(DEV)].cpu().double()
        e = (got - dwi).abs().max().item() / dwi.abs().max().item()
        # fp32 accumulation of N*H*W bf16 products: rtol 1e-4 is the reference's own wgrad tolerance
        # (test_correctness.py:90,127)
        assert e <= 1e-4, (i, e)
    got = dx[:, idx.to(DEV)].cpu().double()
    err = (got - dx64).abs().max().item() / dx64.abs().max().item()
    assert err <= 2.0 ** -7, err                                      # two bf16 roundings (5x5 partial, final sum)


def test_block_conv_statistics_additive_over_batch_halves_and_syncbn_finalize():
    """SyncBN numerics: the per-channel sums of two half batches (two ranks) added together must give what
    slak_bn3_finalize_fwd computes from the full batch (one rank) -- scale / shift / mean / istd -- with the global
    count passed on the device, as the 2-rank path does (slak_b200/block.py)."""
    import ctypes
    from slak_b200 import _lib
    lib = _lib.load()
    N, C, H, W, KL = 64, 24, 28, 28, 49
    _, x, ws = _inputs(N, C, H, W, KL, 5)
    xd = x.to(DEV)
    wd = [w.to(DEV) for w in ws]
    st = _lib.current_stream_ptr()

    def conv_sums(xpart):
        n = xpart.shape[0]
        ys = [torch.empty_like(xpart) for _ in range(3)]
        buf = torch.zeros(C * 6 + 1, dtype=torch.float64, device=DEV)
        need = lib.slak_block_conv_fwd_workspace(n, C, H, W)
        wsp = torch.empty(need, dtype=torch.uint8, device=DEV)
        _lib.check(lib.slak_block_conv_fwd(xpart.data_ptr(), wd[0].data_ptr(), wd[1].data_ptr(), wd[2].data_ptr(),
                                           ys[0].data_ptr(), ys[1].data_ptr(), ys[2].data_ptr(), buf.data_ptr(),
                                           wsp.data_ptr(), wsp.numel(), n, C, H, W, KL, st), "slak_block_conv_fwd")
        buf[C * 6] = float(n * H * W)
        return buf

    def finalize(buf, count, count_dev):
        bnw = [torch.rand(C, device=DEV, generator=torch.Generator(DEV).manual_seed(1 + i)) + 0.5 for i in range(3)]
        bnb = [torch.rand(C, device=DEV, generator=torch.Generator(DEV).manual_seed(9 + i)) - 0.5 for i in range(3)]
        p3 = lambda ts: (ctypes.c_void_p * 3)(*[t.data_ptr() if t is not None else None for t in ts])
        outs = [torch.empty((3, C), device=DEV), torch.empty((C,), device=DEV), torch.empty((3, C), device=DEV),
                torch.empty((3, C), device=DEV)]
        _lib.check(lib.slak_bn3_finalize_fwd(buf.data_ptr(), count, count_dev, p3(bnw), p3(bnb), p3([None] * 3),
                                             p3([None] * 3), 1e-5, 0.1, C, *[o.data_ptr() for o in outs], st),
                   "slak_bn3_finalize_fwd")
        return [o.clone() for o in outs]

    full = conv_sums(xd)
    a, b = conv_sums(xd[:40].contiguous()), conv_sums(xd[40:].contiguous())     # unequal "ranks"
    summed = a + b                                                             # what all_reduce(SUM) leaves on each rank
    assert torch.allclose(summed[:-1], full[:-1], rtol=1e-6, atol=1e-6)
    assert summed[-1].item() == float(N * H * W)
    ref = finalize(full, float(N * H * W), None)
    got = finalize(summed, 0.0, summed[C * 6:].data_ptr())                      # count read from the device
    for r, g_ in zip(ref, got):
        assert torch.allclose(r, g_, rtol=1e-5, atol=1e-6)
