"""Host control flow of the fused Block node (slak_b200/block.py) without a GPU: every C-ABI call is replaced by a
stub that returns success, so the kernels compute nothing, but the whole Python side of forward and backward runs --
argument lists, saved tensors, shapes and dtypes of everything autograd hands back, the eval path and the opt-in
tcgen05-MLP path.  Numerics are the business of the gpu-marked parity tests; this catches a broken call site before a
GPU is involved."""
import contextlib

import pytest
import torch


class _StubLib:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def f(*a, **k):
            self.calls.append(name)
            if name.endswith(("_parts", "_workspace", "_splits")):
                return 4
            if name.endswith("uses_tc"):
                return 1
            return 0
        return f


@pytest.fixture
def stubbed(monkeypatch):
    from slak_b200 import _lib, ops, slak
    lib = _StubLib()
    monkeypatch.setattr(_lib, "load", lambda: lib)
    monkeypatch.setattr(_lib, "current_stream_ptr", lambda: None)
    monkeypatch.setattr(_lib, "check", lambda rc, what: None)
    monkeypatch.setattr(ops, "_check_input", lambda t, n: None)
    monkeypatch.setattr(ops, "_workspace", lambda n, dev: torch.empty(max(int(n), 1), dtype=torch.uint8))
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(slak, "use_sync_bn", False)
    return lib


@pytest.mark.parametrize("dim,fused_mlp", [(8, False), (32, True)])
def test_forward_backward_and_eval_call_sequences(stubbed, monkeypatch, dim, fused_mlp):
    from slak_b200 import block as B
    from slak_b200 import slak
    monkeypatch.setattr(B, "FUSED_MLP", fused_mlp)
    blk = slak.Block(dim=dim, drop_path=0.1, kernel_size=(13, 5), Decom=True, bn=True).train()
    x = torch.randn(2, dim, 16, 16, requires_grad=True)
    y = B.fused_block_forward(blk, x)
    assert y.shape == x.shape and y.dtype == torch.float32
    fwd = list(stubbed.calls)
    assert fwd[:4] == ["slak_block_conv_fwd_workspace", "slak_block_conv_fwd", "slak_bn3_finalize_fwd", "slak_bn3_sum_ln_fwd"]
    assert ("slak_mlp_gemm_nt" in fwd) == fused_mlp and fwd[-1] == "slak_block_residual_fwd"
    del stubbed.calls[:]
    y.backward(torch.ones_like(y))
    bwd = list(stubbed.calls)
    for name in ("slak_block_residual_bwd", "slak_colsum_f32", "slak_bn3_sum_ln_bwd", "slak_bn3_finalize_bwd",
                 "slak_bn3_bwd_apply", "slak_lk_branches_bwd_data_f32", "slak_lk_branches_bwd_filter"):
        assert name in bwd, name
    assert ("slak_mlp_gemm_tn_splitk" in bwd) == fused_mlp and ("slak_gelu_bwd_bias" in bwd) == (not fused_mlp)
    assert bwd.count("slak_mlp_gemm_nt") == (2 if fused_mlp else 0) and bwd.count("slak_mlp_gemm_tn_splitk") == (2 if fused_mlp else 0)
    # one gradient per parameter, with the parameter's shape and dtype
    assert x.grad.shape == x.shape
    for n, p in blk.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape and p.grad.dtype == p.dtype, n
    blk.eval()
    del stubbed.calls[:]
    assert B.fused_block_forward(blk, x.detach()).shape == x.shape
    assert "slak_bn3_eval_affine" in stubbed.calls and "slak_block_conv_fwd" not in stubbed.calls
