"""Fused three-branch forward (tcgen05 banded-Toeplitz kernel where the shape allows, CUDA-core
kernels otherwise) against the oracle, branch by branch."""
import pytest
import torch

from oracle import dwconv as orc
from slak_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = [
    # N, C, H, W, KL   (tensor-core path: bf16, 8 <= H,W <= 62, W % 8 == 0)
    (2, 3, 56, 56, 51), (5, 4, 56, 56, 51), (1, 2, 56, 56, 61), (3, 2, 48, 48, 51), (4, 3, 24, 24, 49),
    (3, 5, 40, 56, 31), (7, 2, 16, 8, 13), (2, 2, 56, 56, 5),
    # CUDA-core path
    (3, 4, 28, 28, 49), (2, 3, 14, 14, 47), (2, 2, 7, 7, 13), (2, 2, 96, 96, 51),
]


@pytest.mark.parametrize("case", CASES)
def test_three_branches_match_oracle(case):
    N, C, H, W, KL = case
    g = torch.Generator().manual_seed(77 + KL + N)
    x = torch.randn(N, C, H, W, generator=g).bfloat16()
    ws = [torch.randn(C, 1, *k, generator=g) * 0.05 for k in ((KL, 5), (5, KL), (5, 5))]
    ys = ops.lk_branches_forward(x.to(DEV), *[w.to(DEV) for w in ws])
    for i, (w, y) in enumerate(zip(ws, ys)):
        ref = orc.fwd_torch(x.double(), orc.round_like(w, torch.bfloat16).double())
        err = (y.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
        assert y.dtype == torch.bfloat16
        assert err <= 2.0 ** -8 + 1e-5, (i, err)


def test_headline_shape_uses_tensor_cores_and_matches_cuda_core_path():
    torch.manual_seed(3)
    N, C, H, W, KL = 32, 96, 56, 56, 51
    x = torch.randn(N, C, H, W, device=DEV).bfloat16()
    ws = [torch.randn(C, 1, *k, device=DEV) * 0.02 for k in ((KL, 5), (5, KL), (5, 5))]
    assert ops.lk_branches_uses_tc(x, KL, 5)
    ys = ops.lk_branches_forward(x, *ws)
    for w, y in zip(ws, ys):
        y_simt = ops.dwconv2d_forward(x, w)
        # same bf16 operands, fp32 accumulation in both: only summation order differs
        d = (y.float() - y_simt.float()).abs().max().item()
        assert d <= 2.0 ** -7 * y_simt.float().abs().max().item(), d


TC_CASES = [(2, 3, 56, 56, 51), (5, 4, 56, 56, 51), (1, 2, 56, 56, 61), (3, 2, 48, 48, 51), (4, 3, 24, 24, 49),
            (3, 5, 40, 56, 31), (7, 2, 16, 8, 13), (2, 2, 56, 56, 5), (33, 2, 56, 56, 51),
            (9, 3, 28, 28, 49), (6, 2, 14, 14, 47), (11, 3, 7, 7, 13), (5, 2, 30, 22, 21), (3, 2, 13, 9, 9),
            (17, 2, 28, 28, 5)]


@pytest.mark.parametrize("case", TC_CASES)
def test_fused_backward_data_and_filter_match_oracle(case):
    N, C, H, W, KL = case
    g = torch.Generator().manual_seed(99 + KL + N)
    x = torch.randn(N, C, H, W, generator=g).bfloat16()
    dys = [torch.randn(N, C, H, W, generator=g).bfloat16() for _ in range(3)]
    ws = [torch.randn(C, 1, *k, generator=g) * 0.05 for k in ((KL, 5), (5, KL), (5, 5))]
    assert ops.lk_branches_bwd_uses_tc(x.to(DEV), KL, 5)
    dx64 = torch.zeros(N, C, H, W, dtype=torch.float64)
    dw64 = []
    for w, dy in zip(ws, dys):
        dxi, dwi = orc.grads_torch(x.double(), orc.round_like(w, torch.bfloat16).double(), dy.double())
        dx64 += dxi
        dw64.append(dwi)
    wd = [w.to(DEV) for w in ws]
    dyd = [d.to(DEV) for d in dys]
    dx = ops.lk_branches_backward_data(*dyd, *wd).cpu().double()
    err = (dx - dx64).abs().max().item() / dx64.abs().max().item()
    # two bf16 roundings (the 5x5 branch is rounded once before the final sum)
    assert err <= 2.0 ** -7, err
    dws = ops.lk_branches_backward_filter(x.to(DEV), *dyd, KL, 5)
    for i, (dw, ref) in enumerate(zip(dws, dw64)):
        assert dw.dtype == torch.float32 and tuple(dw.shape) == tuple(ref.shape)
        e = (dw.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
        assert e <= 1e-4, (i, e)
    # deterministic (fixed-order reduction)
    dws2 = ops.lk_branches_backward_filter(x.to(DEV), *dyd, KL, 5)
    assert all(torch.equal(a, b) for a, b in zip(dws, dws2))


# BatchNorm partial sums gathered in the forward epilogue (slak_block_conv_fwd): [C][6] = (sum, sum of squares) of
# y1, y2, y3 over (N, H, W), from the fp32 accumulators.  Geometries chosen so that both epilogue groups of a CTA,
# several CTAs per channel (T = 64) and several channels per CTA (small classes) all contribute.
STAT_CASES = [(9, 3, 56, 56, 51), (1, 2, 56, 56, 51), (37, 5, 14, 14, 47), (70, 3, 28, 28, 49), (45, 4, 7, 7, 13),
              (300, 2, 56, 56, 5)]


@pytest.mark.parametrize("case", STAT_CASES)
def test_forward_epilogue_statistics_match_float64_sums(case):
    from slak_b200 import _lib
    lib = _lib.load()
    N, C, H, W, KL = case
    g = torch.Generator().manual_seed(5 + N + KL)
    x = torch.randn(N, C, H, W, generator=g).bfloat16().to(DEV)
    ws = [(torch.randn(C, 1, *k, generator=g) * 0.05).to(DEV) for k in ((KL, 5), (5, KL), (5, 5))]
    ys = [torch.empty_like(x) for _ in range(3)]
    sums = torch.empty((C, 6), dtype=torch.float64, device=DEV)
    need = lib.slak_block_conv_fwd_workspace(N, C, H, W)
    assert need > 0
    wsp = torch.empty(need, dtype=torch.uint8, device=DEV)
    rc = lib.slak_block_conv_fwd(x.data_ptr(), ws[0].data_ptr(), ws[1].data_ptr(), ws[2].data_ptr(), ys[0].data_ptr(),
                                 ys[1].data_ptr(), ys[2].data_ptr(), sums.data_ptr(), wsp.data_ptr(), wsp.numel(),
                                 N, C, H, W, KL, _lib.current_stream_ptr())
    _lib.check(rc, "slak_block_conv_fwd")
    ref = ops.lk_branches_forward(x, *ws)
    for i in range(3):
        assert torch.equal(ys[i], ref[i])                       # same kernel with and without the statistics
        # oracle for the sums: the fp32 conv of the bf16-rounded operands, accumulated in float64
        w16 = ws[i].bfloat16().double().cpu()
        y64 = torch.nn.functional.conv2d(x.double().cpu(), w16, padding=(w16.shape[2] // 2, w16.shape[3] // 2), groups=C)
        s_ref = y64.sum((0, 2, 3))
        q_ref = (y64 * y64).sum((0, 2, 3))
        got = sums.cpu()
        scale = float(N * H * W) ** 0.5
        assert torch.allclose(got[:, 2 * i], s_ref, rtol=1e-4, atol=1e-4 * scale), (i, got[:, 2 * i], s_ref)
        assert torch.allclose(got[:, 2 * i + 1], q_ref, rtol=1e-4, atol=1e-5 * scale)
