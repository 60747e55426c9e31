"""Parity of the round-2 DRAFT tcgen05 pointwise-MLP GEMMs (csrc/mlp_tc.cu) against the torch expressions they
replace.  The kernels were written after the round's GPU budget was spent and have not run on hardware yet, so these
tests are opt-in: SLAK_FUSED_MLP_TEST=1 python -m pytest tests/test_mlp_draft_gpu.py -m gpu
(first thing to run in round 2; once green, drop the gate and flip slak_b200.block.FUSED_MLP)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("SLAK_FUSED_MLP_TEST", "0") != "1",
                                 reason="round-2 draft, opt-in (SLAK_FUSED_MLP_TEST=1)")]

SHAPES = [(256, 384, 96), (1000, 768, 192), (128 * 49, 3072, 768), (77, 128, 8), (4096, 1536, 384)]   # M, N, K


@pytest.mark.parametrize("shape", SHAPES)
def test_fc1_gelu_fwd_matches_linear_then_gelu(shape):
    from slak_b200 import _lib
    lib = _lib.load()
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).bfloat16().cuda()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16().cuda()
    b = torch.randn(N, generator=g).cuda()
    h = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    a = torch.empty_like(h)
    rc = lib.slak_mlp_fc1_gelu_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), h.data_ptr(), a.data_ptr(), M, N, K,
                                   _lib.current_stream_ptr())
    _lib.check(rc, "slak_mlp_fc1_gelu_fwd")
    h_ref = (x.float() @ w.float().t() + b.bfloat16().float())          # fp32 accumulate, bf16 bias as in the module path
    # H: one bf16 rounding of the fp32 result (accumulation order differs: 1 ulp of bf16)
    assert torch.allclose(h.float(), h_ref, rtol=2 ** -7, atol=1e-2)
    # A: exact-erf GELU of the STORED h, rounded to bf16
    a_ref = F.gelu(h.float())
    assert torch.allclose(a.float(), a_ref, rtol=2 ** -7, atol=2e-3)


@pytest.mark.parametrize("shape", SHAPES)
def test_fc2_dgelu_bwd_matches_matmul_then_gelu_grad(shape):
    from slak_b200 import _lib
    lib = _lib.load()
    M, N, K = shape
    g = torch.Generator().manual_seed(7 + M + N + K)
    dh2 = torch.randn(M, K, generator=g).bfloat16().cuda()
    wt = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16().cuda()          # W2^T
    h = torch.randn(M, N, generator=g).bfloat16().cuda()
    dh = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    parts = lib.slak_mlp_parts(M, N)
    part = torch.empty(parts, N, dtype=torch.float32, device="cuda")
    rc = lib.slak_mlp_fc2_dgelu_bwd(dh2.data_ptr(), wt.data_ptr(), h.data_ptr(), dh.data_ptr(), part.data_ptr(), M, N, K,
                                    _lib.current_stream_ptr())
    _lib.check(rc, "slak_mlp_fc2_dgelu_bwd")
    hf = h.float().requires_grad_(True)
    F.gelu(hf).backward(dh2.float() @ wt.float().t())
    assert torch.allclose(dh.float(), hf.grad, rtol=2 ** -6, atol=2e-2)
    db = part.sum(0)
    assert torch.allclose(db, dh.float().sum(0), rtol=1e-3, atol=1e-2 * M ** 0.5)   # sums of the stored (rounded) dh
