import sys, math, torch
sys.path.insert(0, "/root/repo")
torch.manual_seed(0)
p0 = (torch.randn(200000) * 0.1).cuda(); g = (torch.randn(200000) * 0.01).cuda()
pa = torch.nn.Parameter(p0.clone()); pa.grad = g.clone()
lr, wd, b1, b2, eps = 4e-3, 0.05, 0.9, 0.999, 1e-8
ref = torch.optim.AdamW([pa], lr=lr, weight_decay=wd, foreach=False, fused=False)
ref.step()
m = ref.state[pa]["exp_avg"]; v = ref.state[pa]["exp_avg_sq"]
p1 = p0 * (1 - lr * wd)
bc1 = 1 - b1 ** 1; bc2 = 1 - b2 ** 1
step_size = lr / bc1; bc2s = bc2 ** 0.5
dens = {"recip": (v.sqrt() * (torch.tensor(1.0, dtype=torch.float32) / torch.tensor(bc2s, dtype=torch.float32)).item()) + eps,
        "truediv_tensor": (v.sqrt() / torch.full_like(v, bc2s)) + eps,
        "torchdiv": (v.sqrt() / bc2s).add_(eps)}
for dn, den in dens.items():
    q = m / den
    cands = {"mul_then_add": p1 + q * (-step_size),
             "addcdiv": p1.clone().addcdiv_(m, den, value=-step_size),
             "fma_f64": (p1.double() + q.double() * float(torch.tensor(-step_size, dtype=torch.float32))).float()}
    for cn, c in cands.items():
        print(dn, cn, "maxdiff vs ref", (c - pa.data).abs().max().item(), "equal", torch.equal(c, pa.data), flush=True)
print("dens equal recip/torchdiv", torch.equal(dens["recip"], dens["torchdiv"]), "truediv/torchdiv", torch.equal(dens["truediv_tensor"], dens["torchdiv"]))
