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
