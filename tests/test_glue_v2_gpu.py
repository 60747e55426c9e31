"""Register-resident glue kernels (csrc/block_glue2.cu) against fp64 torch restatements of the same passes
(models/SLaK.py:89-100 BN + sum, :153-166 LayerNorm / gamma / residual) and against the shared-memory-tile kernels they
replace (SLAK_GLUE_V1=1), on aligned, 8-byte-aligned, unaligned (7 x 7) and ragged planes."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [(3, 16, 64, "small LW8"), (2, 24, 36, "small LW4"), (3, 40, 49, "small LW1"), (2, 768, 64, "widest C"),
         (5, 96, 200, "ragged last tile"), (3, 384, 196, "stage 3 plane"), (2, 768, 49, "stage 4 plane"), (2, 192, 784, "stage 2 plane")]


@pytest.mark.parametrize("N,C,HW,tag", CASES)
def test_glue_v2_matches_fp64(N, C, HW, tag):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import glue_bench
    old = os.environ.get("SLAK_GLUE_V1")
    try:
        out = glue_bench.check(N, C, HW, tag)
    finally:
        if old is None:
            os.environ.pop("SLAK_GLUE_V1", None)
        else:
            os.environ["SLAK_GLUE_V1"] = old
    for k, v in out["v2"].items():
        # tensors stored as bf16: one rounding (2^-9 relative) per element; fp32 results and fixed-order sums: fp32 round-off
        bound = 4e-3 if k in ("ln_fwd.xn", "res_fwd.bf16", "res_bwd.dh2", "ln_bwd.du") else 2e-5
        assert v < bound, f"{k}: rel L2 error {v:.3e} (bound {bound}) [{tag}]  (v1: {out['v1'][k]:.3e})"


@pytest.mark.parametrize("N,C,HW", [(3, 40, 49), (2, 768, 49), (2, 24, 25), (2, 16, 64), (3, 24, 36)])
def test_bn3_bwd_apply_all_plane_alignments(N, C, HW):
    """dy_i = A_i*du + B_i*y_i + C_i (BatchNorm backward of the three branches, models/SLaK.py:89-100 under autograd): the
    vector widths 8 / 4 and the flat walk over unaligned (7 x 7, 5 x 5) planes against the same expression in torch."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import ctypes
    from slak_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N + C + HW)
    du, y1, y2, y3 = [torch.randn(N, C, HW, generator=g).to(dev).bfloat16() for _ in range(4)]
    coef = torch.randn(9, C, generator=g).to(dev)
    outs = [torch.empty_like(du) for _ in range(3)]
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(lib.slak_bn3_bwd_apply(P(du), P(y1), P(y2), P(y3), P(coef), P(outs[0]), P(outs[1]), P(outs[2]), N, C, HW,
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "slak_bn3_bwd_apply")
    torch.cuda.synchronize()
    for i, y in enumerate((y1, y2, y3)):
        ref = coef[i][None, :, None] * du.float() + coef[3 + i][None, :, None] * y.float() + coef[6 + i][None, :, None]
        assert torch.equal(outs[i], ref.bfloat16()) or (outs[i].float() - ref).abs().max() <= 2.0 ** -7 * ref.abs().max()
