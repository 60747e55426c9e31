"""Parity of the tcgen05 pointwise-MLP GEMMs (csrc/mlp_tc.cu: every GEMM of pwconv1 -> GELU -> pwconv2 and of its
backward, models/SLaK.py:157-160) against the torch fp32 expressions they replace: F.linear / F.gelu (exact erf) and
autograd's weight gradients, at the four SLaK-T widths (C = 96, 192, 384, 768 -> hidden 384 .. 3072), with token counts
that are not tile multiples, and through the fused Block against the cuBLAS route."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"

# M (tokens), C : hidden = 4C.  Real stage widths; M chosen ragged (tails in every tile dimension)
SHAPES = [(3136 + 77, 96), (2 * 784 + 5, 192), (6 * 196 + 3, 384), (128 * 49, 768), (77, 8), (1, 16), (300, 40)]


def _lib():
    from slak_b200 import _lib as L
    return L, L.load()


def _nt(lib, L, epi, a, b, bias, aux, out0, out1, colpart, M, N, K):
    p = lambda t: None if t is None else t.data_ptr()
    L.check(lib.slak_mlp_gemm_nt(epi, p(a), p(b), p(bias), p(aux), p(out0), p(out1), p(colpart), M, N, K,
                                 L.current_stream_ptr()), "slak_mlp_gemm_nt")


@pytest.mark.parametrize("shape", SHAPES)
def test_pwconv1_bias_gelu_forward(shape):
    L, lib = _lib()
    M, C = shape
    N, K = 4 * C, C
    g = torch.Generator().manual_seed(M + C)
    x = torch.randn(M, K, generator=g).bfloat16().to(DEV)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16().to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    h = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
    a = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
    _nt(lib, L, 0, x, w, b, None, h, a, None, M, N, K)
    h_ref = F.linear(x.float(), w.float(), b.bfloat16().float())       # fp32 accumulate, bf16 bias as under autocast
    assert torch.allclose(h.float(), h_ref, rtol=2 ** -7, atol=1e-2)   # one bf16 rounding
    assert torch.allclose(a.float(), F.gelu(h.float()), rtol=2 ** -7, atol=2e-3)   # exact-erf GELU of the stored H
    # inference form: only the activation is written
    a2 = torch.empty_like(a)
    _nt(lib, L, 0, x, w, b, None, None, a2, None, M, N, K)
    assert torch.equal(a, a2)


@pytest.mark.parametrize("shape", SHAPES)
def test_pwconv2_bias_forward_and_plain(shape):
    L, lib = _lib()
    M, C = shape
    N, K = C, 4 * C
    g = torch.Generator().manual_seed(3 + M + C)
    a = torch.randn(M, K, generator=g).bfloat16().to(DEV)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16().to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
    _nt(lib, L, 1, a, w, b, None, out, None, None, M, N, K)
    assert torch.allclose(out.float(), F.linear(a.float(), w.float(), b.bfloat16().float()), rtol=2 ** -7, atol=1e-2)
    out2 = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
    _nt(lib, L, 3, a, w, None, None, out2, None, None, M, N, K)
    assert torch.allclose(out2.float(), a.float() @ w.float().t(), rtol=2 ** -7, atol=1e-2)


@pytest.mark.parametrize("shape", SHAPES)
def test_dgelu_backward_and_bias_gradient(shape):
    L, lib = _lib()
    M, C = shape
    N, K = 4 * C, C
    g = torch.Generator().manual_seed(7 + M + C)
    dh2 = torch.randn(M, K, generator=g).bfloat16().to(DEV)
    wt = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16().to(DEV)          # W2^T
    h = torch.randn(M, N, generator=g).bfloat16().to(DEV)
    dh = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
    parts = lib.slak_mlp_parts(M, N)
    part = torch.full((parts, N), float("nan"), dtype=torch.float32, device=DEV)
    _nt(lib, L, 2, dh2, wt, None, h, dh, None, part, M, N, K)
    hf = h.float().requires_grad_(True)
    F.gelu(hf).backward(dh2.float() @ wt.float().t())
    assert torch.allclose(dh.float(), hf.grad, rtol=2 ** -6, atol=2e-2)
    assert torch.allclose(part.sum(0), dh.float().sum(0), rtol=1e-3, atol=1e-2 * M ** 0.5)   # sums of the stored dh


@pytest.mark.parametrize("shape", SHAPES + [(401408 // 8, 96)])
def test_weight_gradients_split_k_over_tokens(shape):
    """dW2 = dH2^T A [C, 4C] and dW1 = dH^T Xn [4C, C]: contraction over the tokens, fp32, deterministic."""
    L, lib = _lib()
    M, C = shape
    g = torch.Generator().manual_seed(11 + M + C)
    for Ma, Nb in ((C, 4 * C), (4 * C, C)):
        p = torch.randn(M, Ma, generator=g).bfloat16().to(DEV)
        q = torch.randn(M, Nb, generator=g).bfloat16().to(DEV)
        splits = lib.slak_mlp_wgrad_splits(M, Ma, Nb)
        assert splits >= 1
        outs = []
        for _ in range(2):
            part = torch.full((splits, Ma * Nb), float("nan"), dtype=torch.float32, device=DEV)
            L.check(lib.slak_mlp_gemm_tn_splitk(p.data_ptr(), q.data_ptr(), part.data_ptr(), M, Ma, Nb,
                                                L.current_stream_ptr()), "slak_mlp_gemm_tn_splitk")
            out = torch.empty(Ma * Nb, dtype=torch.float32, device=DEV)
            L.check(lib.slak_colsum_f32(part.data_ptr(), splits, Ma * Nb, out.data_ptr(), L.current_stream_ptr()), "colsum")
            outs.append(out.view(Ma, Nb))
        assert torch.equal(outs[0], outs[1])                               # fixed order: bitwise repeatable
        ref = (p.double().t() @ q.double())
        err = (outs[0].double() - ref).abs().max().item() / ref.abs().max().item()
        assert err < 1e-4, (Ma, Nb, err)                                   # fp32 accumulation of exact bf16 products


@pytest.mark.parametrize("dim,hw", [(96, 56), (192, 28), (384, 14), (768, 7)])
def test_fused_block_tcgen05_mlp_equals_cublas_route(dim, hw):
    """The whole fused Block (forward + every gradient) with the MLP on this library's GEMMs against the same node
    with torch.mm / F.gelu (cuBLAS) for the MLP: same bf16 operands and fp32 accumulation, so only summation order and
    the erf approximation (1.5e-7) differ."""
    import copy
    from slak_b200 import block as B
    from slak_b200 import slak
    torch.manual_seed(5)
    slak.use_sync_bn = False
    ks = {56: 51, 28: 49, 14: 47, 7: 13}[hw]
    blk = slak.Block(dim=dim, drop_path=0.0, layer_scale_init_value=1.0, kernel_size=(ks, 5), Decom=True, bn=True).to(DEV).train()
    ref = copy.deepcopy(blk)
    n = 4 if hw >= 28 else 16
    x = torch.randn(n, dim, hw, hw, device=DEV)
    cot = torch.randn(n, dim, hw, hw, device=DEV)
    outs = []
    for fused, m in ((True, blk), (False, ref)):
        B.FUSED_MLP = fused
        try:
            xi = x.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = m(xi)
            (y * cot).sum().backward()
            outs.append((y.detach(), xi.grad, {k: p.grad for k, p in m.named_parameters()}))
        finally:
            B.FUSED_MLP = True
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
    (y0, dx0, g0), (y1, dx1, g1) = outs
    assert rel(y0, y1) < 3e-3, rel(y0, y1)
    assert rel(dx0, dx1) < 1e-2, rel(dx0, dx1)
    for k in g0:
        assert rel(g0[k], g1[k]) < 1e-2, (k, rel(g0[k], g1[k]))


def test_cast_transpose_and_wide_fold():
    L, lib = _lib()
    for R, Cc in [(384, 96), (96, 384), (3072, 768), (40, 8), (33, 17)]:
        w = torch.randn(R, Cc, device=DEV)
        wb = torch.empty(R, Cc, dtype=torch.bfloat16, device=DEV)
        wt = torch.empty(Cc, R, dtype=torch.bfloat16, device=DEV)
        L.check(lib.slak_cast_transpose_bf16(w.data_ptr(), wb.data_ptr(), wt.data_ptr(), R, Cc, L.current_stream_ptr()), "ct")
        assert torch.equal(wb, w.bfloat16()) and torch.equal(wt, w.bfloat16().t().contiguous())
    # split-K partial fold over few rows / very many columns (the weight-gradient partials)
    for rows, cols in [(2, 768 * 3072), (8, 384 * 1536), (1, 16384), (3, 20000)]:
        part = torch.randn(rows, cols, device=DEV)
        out = torch.empty(cols, device=DEV)
        L.check(lib.slak_colsum_f32(part.data_ptr(), rows, cols, out.data_ptr(), L.current_stream_ptr()), "colsum")
        ref = part[0].clone()
        for r in range(1, rows):
            ref += part[r]
        assert torch.equal(out, ref)            # rows added in order: bitwise the sequential sum
