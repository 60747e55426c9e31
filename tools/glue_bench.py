"""A/B of the Block's glue kernels: register-resident (block_glue2.cu) vs shared-memory tile (block_fused.cu, SLAK_GLUE_V1=1).
Each kernel is checked against an fp64 torch restatement on a small case and on the four SLaK-T stage shapes, then timed
(CUDA events, inputs larger than L2 rotated between launches).   usage: python tools/glue_bench.py [--quick]"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slak_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def P(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def ck(rc, what):
    _lib.check(rc, what)


def make(N, C, HW, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    d = dict(
        y=[(r(N, C, HW) * (0.5 + i)).to(dev).bfloat16() for i in range(3)],
        scale=(r(3, C) * 0.3 + 1).to(dev), shift=(r(C) * 0.5).to(dev), lnw=(r(C) * 0.2 + 1).to(dev), lnb=(r(C) * 0.1).to(dev),
        x=r(N, C, HW).to(dev), h2=r(N, HW, C).to(dev).bfloat16(), gamma=(r(C) * 0.1).to(dev), dp=(torch.rand(N, generator=g) + 0.5).to(dev),
        dout=r(N, C, HW).to(dev), dxn=r(N, HW, C).to(dev).bfloat16())
    return d


def run_ln_fwd(d, N, C, HW):
    xn = torch.empty(N, HW, C, device=dev, dtype=torch.bfloat16)
    mu = torch.empty(N * HW, device=dev)
    rstd = torch.empty(N * HW, device=dev)
    f = lambda: ck(lib.slak_bn3_sum_ln_fwd(P(d["y"][0]), P(d["y"][1]), P(d["y"][2]), P(d["scale"]), P(d["shift"]), P(d["lnw"]), P(d["lnb"]),
                                           1e-6, P(xn), P(mu), P(rstd), N, C, HW, st), "ln_fwd")
    return f, (xn, mu, rstd)


def ref_ln_fwd(d, N, C, HW):
    u = sum(d["scale"][i].double()[None, :, None] * d["y"][i].double() for i in range(3)) + d["shift"].double()[None, :, None]
    m = u.mean(1, keepdim=True)
    v = ((u - m) ** 2).mean(1, keepdim=True)
    r = (v + 1e-6).rsqrt()
    xn = (u - m) * r * d["lnw"].double()[None, :, None] + d["lnb"].double()[None, :, None]
    return xn.permute(0, 2, 1).contiguous(), m.reshape(-1), r.reshape(-1), u


def run_res_fwd(d, N, C, HW):
    out = torch.empty(N, C, HW, device=dev)
    ob = torch.empty(N, C, HW, device=dev, dtype=torch.bfloat16)
    f = lambda: ck(lib.slak_block_residual_fwd(P(d["x"]), P(d["h2"]), P(d["gamma"]), P(d["dp"]), P(out), P(ob), N, C, HW, st), "res_fwd")
    return f, (out, ob)


def run_res_bwd(d, N, C, HW):
    parts = lib.slak_block_residual_bwd_parts(N, C, HW)
    dh2 = torch.empty(N, HW, C, device=dev, dtype=torch.bfloat16)
    part = torch.empty(parts, 2, C, device=dev)
    f = lambda: ck(lib.slak_block_residual_bwd(P(d["dout"]), P(d["h2"]), P(d["gamma"]), P(d["dp"]), P(dh2), P(part), N, C, HW, st), "res_bwd")
    return f, (dh2, part)


def run_ln_bwd(d, N, C, HW, mu, rstd):
    parts = lib.slak_bn3_sum_ln_bwd_parts(N, C, HW)
    du = torch.empty(N, C, HW, device=dev, dtype=torch.bfloat16)
    part = torch.empty(parts, 6, C, device=dev)
    f = lambda: ck(lib.slak_bn3_sum_ln_bwd(P(d["dxn"]), P(d["y"][0]), P(d["y"][1]), P(d["y"][2]), P(d["scale"]), P(d["shift"]), P(d["lnw"]),
                                           P(mu), P(rstd), P(du), P(part), N, C, HW, st), "ln_bwd")
    return f, (du, part)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def check(N, C, HW, tag):
    d = make(N, C, HW, seed=C + HW)
    out = {}
    for mode in ("v2", "v1"):
        os.environ["SLAK_GLUE_V1"] = "1" if mode == "v1" else "0"
        xr, mr, rr, u = ref_ln_fwd(d, N, C, HW)
        f, (xn, mu, rstd) = run_ln_fwd(d, N, C, HW); f(); torch.cuda.synchronize()
        e = {"ln_fwd.xn": rel(xn, xr), "ln_fwd.mu": rel(mu, mr), "ln_fwd.rstd": rel(rstd, rr)}
        f, (o, ob) = run_res_fwd(d, N, C, HW); f(); torch.cuda.synchronize()
        oref = d["x"].double() + d["dp"].double()[:, None, None] * d["gamma"].double()[None, :, None] * d["h2"].double().permute(0, 2, 1)
        e["res_fwd.out"] = rel(o, oref); e["res_fwd.bf16"] = rel(ob, oref)
        f, (dh2, part) = run_res_bwd(d, N, C, HW); f(); torch.cuda.synchronize()
        g0 = d["dout"].double() * d["dp"].double()[:, None, None]
        dref = (g0 * d["gamma"].double()[None, :, None]).permute(0, 2, 1)
        e["res_bwd.dh2"] = rel(dh2, dref)
        e["res_bwd.dgamma"] = rel(part[:, 0].double().sum(0), (g0 * d["h2"].double().permute(0, 2, 1)).sum((0, 2)))
        e["res_bwd.colsum"] = rel(part[:, 1].double().sum(0), dh2.double().sum((0, 1)))
        # LayerNorm backward against autograd in fp64
        mu32, rstd32 = mr.float().contiguous(), rr.float().contiguous()
        f, (du, part) = run_ln_bwd(d, N, C, HW, mu32, rstd32); f(); torch.cuda.synchronize()
        uu = u.clone().requires_grad_(True)
        w = d["lnw"].double().clone().requires_grad_(True)
        b = d["lnb"].double().clone().requires_grad_(True)
        m = uu.mean(1, keepdim=True); v = ((uu - m) ** 2).mean(1, keepdim=True)
        xn2 = (uu - m) * (v + 1e-6).rsqrt() * w[None, :, None] + b[None, :, None]
        xn2.backward(d["dxn"].double().permute(0, 2, 1))
        e["ln_bwd.du"] = rel(du, uu.grad)
        e["ln_bwd.dlnw"] = rel(part[:, 0].double().sum(0), w.grad)
        e["ln_bwd.dlnb"] = rel(part[:, 1].double().sum(0), b.grad)
        dub = du.double()
        e["ln_bwd.S0"] = rel(part[:, 2].double().sum(0), dub.sum((0, 2)))
        for i in range(3):
            e[f"ln_bwd.S{i + 1}"] = rel(part[:, 3 + i].double().sum(0), (dub * d["y"][i].double()).sum((0, 2)))
        out[mode] = e
    worst = max(out["v2"].values())
    print(f"[{tag}] N{N} C{C} HW{HW}: worst v2 rel err {worst:.2e}")
    for k in out["v2"]:
        flag = "" if out["v2"][k] < 2.5 * max(out["v1"][k], 4e-3 if ".du" in k or "xn" in k or "bf16" in k or "dh2" in k else 1e-4) else "   <-- CHECK"
        print(f"    {k:16s} v2 {out['v2'][k]:.2e}   v1 {out['v1'][k]:.2e}{flag}")
    return out


def timeit(N, C, HW, reps=20):
    nset = max(2, int(400e6 / (N * C * HW * 12)) + 1)          # rotate > 400 MB of inputs: nothing survives in L2
    sets = [make(N, C, HW, seed=i) for i in range(min(nset, 6))]
    res = {}
    for mode in ("v1", "v2"):
        os.environ["SLAK_GLUE_V1"] = "1" if mode == "v1" else "0"
        for name in ("ln_fwd", "res_fwd", "res_bwd", "ln_bwd"):
            fs = []
            for d in sets:
                if name == "ln_fwd":
                    fs.append(run_ln_fwd(d, N, C, HW)[0])
                elif name == "res_fwd":
                    fs.append(run_res_fwd(d, N, C, HW)[0])
                elif name == "res_bwd":
                    fs.append(run_res_bwd(d, N, C, HW)[0])
                else:
                    mu = torch.randn(N * HW, device=dev) * 0.1
                    rstd = torch.rand(N * HW, device=dev) + 0.5
                    fs.append(run_ln_bwd(d, N, C, HW, mu, rstd)[0])
            for f in fs:
                f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps):
                fs[i % len(fs)]()
            e1.record(); torch.cuda.synchronize()
            res[(mode, name)] = e0.elapsed_time(e1) * 1e3 / reps
    bytes_per = {"ln_fwd": 8, "res_fwd": 12, "res_bwd": 8, "ln_bwd": 10}
    row = {}
    for name in ("ln_fwd", "res_fwd", "res_bwd", "ln_bwd"):
        v1, v2 = res[("v1", name)], res[("v2", name)]
        gbs = N * C * HW * bytes_per[name] / v2 / 1e3
        print(f"  {name:8s} N{N} C{C} HW{HW}:  v1 {v1:7.1f} us   v2 {v2:7.1f} us   ({v1 / v2:4.2f}x)  {gbs:7.0f} GB/s")
        row[name] = {"v1_us": round(v1, 1), "v2_us": round(v2, 1), "v2_GBps": round(gbs)}
    return row


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    for (N, C, HW, tag) in [(3, 16, 64, "small LW8"), (2, 24, 36, "small LW4"), (3, 40, 49, "small LW1"), (2, 768, 64, "wide"),
                            (5, 96, 200, "ragged LW8"), (3, 384, 196, "stage3"), (2, 768, 49, "stage4")]:
        check(N, C, HW, tag)
    if not quick:
        table = {}
        for (C, HW) in [(96, 3136), (192, 784), (384, 196), (768, 49)]:
            table[f"C{C}_HW{HW}"] = timeit(128, C, HW)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(table, open("gpurun_out/glue_bench.json", "w"), indent=1)
