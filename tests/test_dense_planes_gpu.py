"""Small planes as dense per-channel GEMMs (csrc/dwconv_tc_dense.cu): forward of the three Decom branches with the BatchNorm
sums, and their fused data gradient with the fp32 shortcut added, against the oracle (fp64 F.conv2d on the bf16-rounded
operands, oracle/dwconv.py) and against the banded-Toeplitz kernels they replace (SLAK_DENSE_PLANES=0).
Shapes: the two SLaK-T stages they serve at the bench batch (channel subsets checked), ragged batches (two image tiles,
the second partial), plane sizes with every alignment class (P % 8 == 0, % 4, even, odd) and all group sizes G."""
import ctypes
import os

import pytest
import torch

from oracle import dwconv as orc
from slak_b200 import _lib, ops

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = [(128, 384, 14, 14, 47), (128, 768, 7, 7, 13), (130, 6, 14, 14, 47), (5, 8, 7, 7, 13), (3, 4, 12, 12, 27), (2, 8, 10, 10, 21),
         (3, 16, 4, 4, 9), (2, 8, 8, 8, 17), (2, 8, 9, 9, 17), (2, 8, 13, 15, 31), (200, 16, 7, 7, 13)]


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _subset(C):
    idx = sorted({0, 1, C // 3, C // 2, C // 2 + 1, C - 2, C - 1})
    return torch.tensor([i for i in idx if 0 <= i < C])


def _run(N, C, H, W, KL, seed):
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, C, H, W, generator=g).bfloat16()
    ws = [torch.randn(C, 1, *k, generator=g) * 0.05 for k in ((KL, 5), (5, KL), (5, 5))]
    dys = [torch.randn(N, C, H, W, generator=g).bfloat16() for _ in range(3)]
    add = torch.randn(N, C, H, W, generator=g)
    xd, wd, dyd, addd = x.to(DEV), [w.to(DEV) for w in ws], [d.to(DEV) for d in dys], add.to(DEV)
    st = torch.cuda.current_stream().cuda_stream
    ys = [torch.empty_like(xd) for _ in range(3)]
    sums = torch.empty(C * 6, dtype=torch.float64, device=DEV)
    need = lib.slak_block_conv_fwd_workspace(N, C, H, W)
    wsb = torch.empty(max(need, 16), dtype=torch.uint8, device=DEV)
    _lib.check(lib.slak_block_conv_fwd(_p(xd), _p(wd[0]), _p(wd[1]), _p(wd[2]), _p(ys[0]), _p(ys[1]), _p(ys[2]), _p(sums), _p(wsb),
                                       wsb.numel(), N, C, H, W, KL, st), "slak_block_conv_fwd")
    dx = torch.empty(N, C, H, W, device=DEV)
    tmp = torch.empty_like(xd)
    _lib.check(lib.slak_lk_branches_bwd_data_f32(_p(dyd[0]), _p(dyd[1]), _p(dyd[2]), _p(wd[0]), _p(wd[1]), _p(wd[2]), _p(addd), _p(dx),
                                                 _p(tmp), N, C, H, W, KL, 5, st), "slak_lk_branches_bwd_data_f32")
    dws = ops.lk_branches_backward_filter(xd, *dyd, KL, 5)
    dws2 = ops.lk_branches_backward_filter(xd, *dyd, KL, 5)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(dws, dws2))          # fixed-order reduction: bitwise repeatable
    return x, ws, dys, add, [y.cpu() for y in ys], sums.cpu().view(C, 6), dx.cpu(), [d.cpu() for d in dws]


@pytest.mark.parametrize("case", CASES)
def test_dense_planes_vs_oracle(case):
    N, C, H, W, KL = case
    if not ops.lk_branches_uses_tc(torch.empty(N, C, H, W, dtype=torch.bfloat16, device=DEV), KL, 5):
        pytest.skip("shape outside the tensor-core classes")
    os.environ["SLAK_DENSE_PLANES"] = "1"
    x, ws, dys, add, ys, sums, dx, dws = _run(N, C, H, W, KL, 7 + N + C + KL)
    idx = _subset(C)
    xs = x[:, idx].double()
    dx64 = add[:, idx].double().clone()
    for i, w in enumerate(ws):
        wr = orc.round_like(w[idx], torch.bfloat16).double()
        ref = orc.fwd_torch(xs, wr)
        got = ys[i][:, idx].double()
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        assert err <= 2.0 ** -8 + 1e-5, ("y", i, err)          # one bf16 rounding of an fp32-accumulated sum
        # BatchNorm sums of the fp32 results (before rounding)
        s_ref, q_ref = ref.sum((0, 2, 3)), (ref * ref).sum((0, 2, 3))
        assert torch.allclose(sums[idx, 2 * i], s_ref, rtol=2e-4, atol=2e-3 * ref.abs().max().item() * (N * H * W) ** 0.5), ("sum", i)
        assert torch.allclose(sums[idx, 2 * i + 1], q_ref, rtol=2e-4), ("sumsq", i)
        dxi, dwi = orc.grads_torch(xs, wr, dys[i][:, idx].double())
        dx64 += dxi
        e = (dws[i][idx].double() - dwi).abs().max().item() / dwi.abs().max().item()
        assert e <= 1e-4, ("dw", i, e)                            # the reference's own wgrad tolerance (test_correctness.py:90,127)
    err = (dx[:, idx].double() - dx64).abs().max().item() / dx64.abs().max().item()
    assert err <= 2e-5, ("dx", err)                           # fp32 accumulation of bf16 products, fp32 out
    # the banded-Toeplitz kernels on the same inputs
    os.environ["SLAK_DENSE_PLANES"] = "0"
    try:
        _, _, _, _, ys0, sums0, dx0, dws0 = _run(N, C, H, W, KL, 7 + N + C + KL)
    finally:
        os.environ["SLAK_DENSE_PLANES"] = "1"
    for a, b in zip(ys, ys0):
        assert (a.float() - b.float()).abs().max().item() <= 2.0 ** -7 * b.float().abs().max().item()
    assert torch.allclose(sums, sums0, rtol=1e-3, atol=1e-2 * (N * H * W) ** 0.5)
    # (the banded path hands the 5 x 5 branch's partial gradient over in bf16: 2^-9 of that part)
    assert torch.allclose(dx, dx0, rtol=4e-3, atol=4e-3 * dx0.abs().max().item())
    for a, b in zip(dws, dws0):
        assert torch.allclose(a, b, rtol=2e-4, atol=2e-4 * b.abs().max().item())
