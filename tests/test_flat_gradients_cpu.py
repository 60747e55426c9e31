"""FlatGradients (slak_b200/ddp.py): a step is zero_grad(); backward(s); finish() -- gradients are taken over from autograd and
gathered into the flat buffer; results equal plain torch accumulation (main.py:374-376 wraps the model for this effect only)."""
import torch


def _net():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2, bias=False))


def test_flat_gradients_match_plain_accumulation():
    from slak_b200.ddp import FlatGradients
    a, b = _net(), _net()
    xs = [torch.randn(4, 6, generator=torch.Generator().manual_seed(k)) for k in range(2)]
    for x in xs:                                          # reference: two micro-steps accumulated by autograd
        a(x).pow(2).sum().backward()
    flat = FlatGradients(b)
    assert all(p.grad.data_ptr() >= flat.flat.data_ptr() for p in b.parameters())
    for step in range(2):                                 # the second step must not see the first one's gradients
        flat.zero_grad()
        assert all(p.grad is None for p in b.parameters())
        for x in xs:
            flat.arm(last_micro_step=False)
            b(x).pow(2).sum().backward()
        flat.finish()
        for pa, pb in zip(a.parameters(), b.parameters()):
            assert pb.grad.data_ptr() == flat.views[pb].data_ptr()          # views of the flat buffer again
            assert torch.equal(pa.grad, pb.grad)


def test_flat_gradients_unused_parameter_and_in_place_mode():
    from slak_b200.ddp import FlatGradients
    net = _net()
    extra = torch.nn.Linear(2, 2)                         # never used in forward
    mod = torch.nn.ModuleList([net, extra])
    flat = FlatGradients(mod)
    flat.flat.fill_(7.0)                                  # stale values must not survive a step
    flat.zero_grad()
    net(torch.ones(2, 6)).sum().backward()
    flat.finish()
    assert all(float(p.grad.abs().sum()) == 0.0 for p in extra.parameters())
    g1 = [p.grad.clone() for p in net.parameters()]
    # without zero_grad() the views stay in place and autograd accumulates into them
    net(torch.ones(2, 6)).sum().backward()
    flat.finish()
    for g, p in zip(g1, net.parameters()):
        assert torch.allclose(p.grad, 2 * g)
