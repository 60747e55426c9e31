"""LayerNorm over channels of NCHW (slak_layernorm2d_*) against the oracle's restatement of models/SLaK.py:256-261
(oracle/slak_model.py::layer_norm_cf, evaluated in fp64 on the CPU) and autograd through that formula."""
import pytest
import torch

from oracle import slak_model as omodel

pytestmark = pytest.mark.gpu

CASES = [
    # N, C, H, W
    (2, 96, 56, 56),
    (3, 192, 28, 28),
    (2, 384, 14, 14),
    (5, 7, 3, 5),        # ragged: fewer pixels than one warp per image, odd C
    (1, 1, 1, 1),
    (2, 40, 9, 33),
]


def _ref(x, w, b, g):
    x64 = x.detach().double().cpu().requires_grad_(True)
    w64 = w.detach().double().cpu().requires_grad_(True)
    b64 = b.detach().double().cpu().requires_grad_(True)
    y = omodel.layer_norm_cf(x64, w64, b64, eps=1e-6)
    y.backward(g.detach().double().cpu())
    return y.detach(), x64.grad, w64.grad, b64.grad


@pytest.mark.parametrize("shape", CASES)
@pytest.mark.parametrize("xdt,ydt", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                     (torch.bfloat16, torch.float32), (torch.bfloat16, torch.bfloat16)])
def test_layernorm2d_matches_oracle(shape, xdt, ydt):
    from slak_b200 import ops
    torch.manual_seed(sum(shape))
    N, C, H, W = shape
    x = (torch.randn(N, C, H, W, device="cuda") * 2 + 0.5).to(xdt).requires_grad_(True)
    w = (torch.randn(C, device="cuda") * 0.5 + 1).requires_grad_(True)
    b = torch.randn(C, device="cuda").requires_grad_(True)
    g = torch.randn(N, C, H, W, device="cuda").to(ydt)
    y = ops.layernorm2d(x, w, b, 1e-6, ydt)
    assert y.dtype == ydt and y.shape == x.shape
    y.backward(g)
    y_ref, dx_ref, dw_ref, db_ref = _ref(x, w, b, g)
    # tolerance: fp32 arithmetic (1e-5 relative to the row scale) plus one rounding of the stored dtype
    tol_y = 2e-5 if ydt == torch.float32 else 1e-2
    tol_x = 5e-5 if xdt == torch.float32 else 1.5e-2
    assert torch.allclose(y.double().cpu(), y_ref, rtol=tol_y, atol=tol_y * 4)
    if C > 1:
        assert torch.allclose(x.grad.double().cpu(), dx_ref, rtol=tol_x, atol=tol_x * 4)
    red = max(1.0, (N * H * W) ** 0.5)
    assert torch.allclose(w.grad.double().cpu(), dw_ref, rtol=1e-4, atol=2e-5 * red)
    assert torch.allclose(b.grad.double().cpu(), db_ref, rtol=1e-4, atol=2e-5 * red)


def test_layernorm2d_is_deterministic():
    from slak_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(8, 96, 28, 28, device="cuda", requires_grad=True)
    w = torch.randn(96, device="cuda", requires_grad=True)
    b = torch.randn(96, device="cuda", requires_grad=True)
    g = torch.randn_like(x)
    outs = []
    for _ in range(2):
        for t in (x, w, b):
            t.grad = None
        ops.layernorm2d(x, w, b, 1e-6, None).backward(g)
        outs.append((x.grad.clone(), w.grad.clone(), b.grad.clone()))
    for a, c in zip(*outs):
        assert torch.equal(a, c)


def test_module_channels_first_uses_kernel_and_matches_formula():
    from slak_b200 import ops, slak
    torch.manual_seed(1)
    ln = slak.LayerNorm(48, eps=1e-6, data_format="channels_first").cuda()
    with torch.no_grad():
        ln.weight.normal_(1, 0.2)
        ln.bias.normal_()
    x = torch.randn(4, 48, 10, 12, device="cuda")
    n0 = ops.launch_count()
    y = ln(x)
    assert ops.launch_count() == n0 + 1
    ref = omodel.layer_norm_cf(x.double().cpu(), ln.weight.detach().double().cpu(), ln.bias.detach().double().cpu(), eps=1e-6)
    assert torch.allclose(y.double().cpu(), ref, rtol=2e-5, atol=1e-4)
