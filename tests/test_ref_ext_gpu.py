"""This repo's kernels against the reference's OWN CUDA operator on the same GPU: the MegEngine-CUTLASS example-19
extension built for sm_100a from the reference's sources (oracle/build_ref_ext.py -> oracle/_ref/ext, shipped with the
snapshot).  north_star: "outputs match the reference CUTLASS path within 1e-3 rel fp32".  Skipped when the
extension was not built."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXT_DIR = os.path.join(ROOT, "oracle", "_ref", "ext")
DEV = "cuda"


def _ext():
    if not os.path.exists(os.path.join(EXT_DIR, "_depthwise_conv2d_implicit_gemm_C.so")):
        pytest.skip("oracle/_ref/ext not built (python oracle/build_ref_ext.py)")
    if EXT_DIR not in sys.path:
        sys.path.insert(0, EXT_DIR)
    import _depthwise_conv2d_implicit_gemm_C as ext
    return ext


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


# the reference's own test grid (test_correctness.py:16-35) plus the SLaK geometries it never tests
CASES = [(1, 64, 16, 3, 3), (16, 64, 32, 7, 7), (16, 192, 16, 13, 13), (1, 192, 32, 31, 31),
         (4, 96, 56, 51, 5), (4, 96, 56, 5, 51), (4, 96, 56, 5, 5), (8, 192, 28, 49, 5), (8, 384, 14, 5, 47),
         (16, 768, 7, 13, 5)]


@pytest.mark.parametrize("case", CASES)
def test_fp32_ops_match_reference_extension(case):
    from slak_b200 import ops
    ext = _ext()
    N, C, HW, kh, kw = case
    torch.manual_seed(kh * 100 + kw + N)
    x = torch.randn(N, C, HW, HW, device=DEV)
    g = torch.randn(N, C, HW, HW, device=DEV)
    w = torch.randn(C, 1, kh, kw, device=DEV) * 0.05
    assert _rel(ops.dwconv2d_forward(x, w), ext.forward_fp32(x, w)) < 1e-5
    assert _rel(ops.dwconv2d_backward_data(g, w), ext.backward_data_fp32(g, w)) < 1e-5
    # the reference accumulates with fp32 atomics in a launch-dependent order: its own tolerance is rtol 1e-4
    assert _rel(ops.dwconv2d_backward_filter(g, x, w), ext.backward_filter_fp32(g, x, w)) < 1e-4


@pytest.mark.parametrize("case", [(4, 96, 56, 51), (8, 192, 28, 49), (8, 384, 14, 47), (16, 768, 7, 13)])
def test_bf16_tensor_core_branches_within_1e3_of_reference_extension_fp32(case):
    """The tcgen05 path (bf16 operands, fp32 accumulate) against the reference extension run in fp32 on the SAME
    bf16-representable inputs: only the accumulation order and the final bf16 rounding differ."""
    from slak_b200 import ops
    ext = _ext()
    N, C, HW, KL = case
    torch.manual_seed(KL)
    x = torch.randn(N, C, HW, HW, device=DEV).bfloat16()
    ws = [(torch.randn(C, 1, *k, device=DEV) * 0.05) for k in ((KL, 5), (5, KL), (5, 5))]
    ys = ops.lk_branches_forward(x, *ws)
    for w, y in zip(ws, ys):
        ref = ext.forward_fp32(x.float(), w.bfloat16().float())
        assert _rel(y.float(), ref) <= 2.0 ** -8 + 1e-5          # one bf16 rounding of the output
        assert (y.float() - ref).abs().mean().item() / ref.abs().mean().item() < 1e-3 * 3
