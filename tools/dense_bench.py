"""Timing of the dense small-plane kernels alone (forward + statistics, fused data gradient) at the two SLaK-T shapes,
with the SLAK_DENSE_DBG experiments: 0 = product, 1 = no matrix build, 2 = no global stores, 3 = both."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slak_b200 import _lib  # noqa: E402

lib = _lib.load()
DEV = "cuda"


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def run(N, C, H, W, KL, reps=10):
    g = torch.Generator().manual_seed(1)
    nset = 4
    xs = [torch.randn(N, C, H, W, generator=g).bfloat16().to(DEV) for _ in range(nset)]
    dys = [[torch.randn(N, C, H, W, generator=g).bfloat16().to(DEV) for _ in range(3)] for _ in range(nset)]
    ws = [(torch.randn(C, 1, *k, generator=g) * 0.05).to(DEV) for k in ((KL, 5), (5, KL), (5, 5))]
    add = torch.randn(N, C, H, W, device=DEV)
    ys = [torch.empty_like(xs[0]) for _ in range(3)]
    dx = torch.empty_like(add)
    tmp = torch.empty_like(xs[0])
    sums = torch.empty(C * 6, dtype=torch.float64, device=DEV)
    need = lib.slak_block_conv_fwd_workspace(N, C, H, W)
    wsb = torch.empty(max(need, 16), dtype=torch.uint8, device=DEV)
    st = torch.cuda.current_stream().cuda_stream

    def fwd(i):
        _lib.check(lib.slak_block_conv_fwd(P(xs[i % nset]), P(ws[0]), P(ws[1]), P(ws[2]), P(ys[0]), P(ys[1]), P(ys[2]), P(sums), P(wsb),
                                           wsb.numel(), N, C, H, W, KL, st), "fwd")

    def dgrad(i):
        d = dys[i % nset]
        _lib.check(lib.slak_lk_branches_bwd_data_f32(P(d[0]), P(d[1]), P(d[2]), P(ws[0]), P(ws[1]), P(ws[2]), P(add), P(dx), P(tmp),
                                                     N, C, H, W, KL, 5, st), "dgrad")

    from slak_b200 import ops

    def wgrad(i):
        ops.lk_branches_backward_filter(xs[i % nset], *dys[i % nset], KL, 5)

    out = {}
    for name, f in (("fwd", fwd), ("dgrad", dgrad), ("wgrad", wgrad)):
        f(0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            f(i)
        e1.record(); torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) * 1e3 / reps
    return out


if __name__ == "__main__":
    for shape in [(128, 384, 14, 14, 47), (128, 768, 7, 7, 13)]:
        for dense, dbg in ((0, 0), (1, 0), (1, 3)):
            os.environ["SLAK_DENSE_PLANES"] = str(dense)
            os.environ["SLAK_DENSE_DBG"] = str(dbg)
            r = run(*shape)
            print(f"N{shape[0]} C{shape[1]} {shape[2]}x{shape[3]} K{shape[4]}  dense={dense} dbg={dbg}:  fwd {r['fwd']:7.1f} us   dgrad {r['dgrad']:7.1f} us   wgrad {r['wgrad']:7.1f} us", flush=True)
