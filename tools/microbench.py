"""Kernel microbenchmarks on one B200 (CUDA events, L2 flushed between iterations).
usage: python tools/microbench.py [fwd_tc|simt|all] ...   prints one JSON line per kernel."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from slak_b200 import ops

DEV = "cuda"
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = None


def timeit(fn, iters=20, warm=3):
    global flush
    if flush is None:
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def report(name, us, best, nbytes, flops):
    print(json.dumps({"kernel": name, "median_us": round(us, 1), "best_us": round(best, 1),
                      "GBps": round(nbytes / us / 1e3, 1), "hbm_frac": round(nbytes / us / 1e3 / PEAK, 4),
                      "TFLOPs": round(flops / us / 1e6, 2)}), flush=True)


def eff(H, k):
    return sum(min(H, p + k // 2 + 1) - max(0, p - k // 2) for p in range(H))


def stage(N, C, H, KL, which):
    x = torch.randn(N, C, H, H, device=DEV).bfloat16()
    dy = torch.randn(N, C, H, H, device=DEV).bfloat16()
    w1 = torch.randn(C, 1, KL, 5, device=DEV) * 0.02
    w2 = torch.randn(C, 1, 5, KL, device=DEV) * 0.02
    w3 = torch.randn(C, 1, 5, 5, device=DEV) * 0.02
    e = x.numel() * 2
    f1 = 2 * N * C * eff(H, KL) * eff(H, 5)
    f3 = 2 * N * C * eff(H, 5) * eff(H, 5)
    tag = f"N{N}_C{C}_{H}x{H}_K{KL}"
    if which in ("simt", "all"):
        us, b = timeit(lambda: ops.dwconv2d_forward(x, w1)); report(f"simt_fwd_{KL}x5_{tag}", us, b, 2 * e, f1)
        us, b = timeit(lambda: ops.dwconv2d_forward(x, w2)); report(f"simt_fwd_5x{KL}_{tag}", us, b, 2 * e, f1)
        us, b = timeit(lambda: ops.dwconv2d_forward(x, w3)); report(f"simt_fwd_5x5_{tag}", us, b, 2 * e, f3)
        us, b = timeit(lambda: ops.dwconv2d_backward_data(dy, w1)); report(f"simt_dgrad_{KL}x5_{tag}", us, b, 2 * e, f1)
        us, b = timeit(lambda: ops.dwconv2d_backward_filter(dy, x, w1)); report(f"simt_wgrad_{KL}x5_{tag}", us, b, 2 * e, f1)
        us, b = timeit(lambda: ops.dwconv2d_backward_filter(dy, x, w2)); report(f"simt_wgrad_5x{KL}_{tag}", us, b, 2 * e, f1)
    if which in ("fwd_tc", "all"):
        us, b = timeit(lambda: ops.lk_branches_forward(x, w1, w2, w3))
        report(f"lk3_fwd_{'tc' if ops.lk_branches_uses_tc(x, KL, 5) else 'simt'}_{tag}", us, b, 4 * e, 2 * f1 + f3)
    if which in ("bwd_tc", "all") and hasattr(ops, "lk_branches_backward_data"):
        us, b = timeit(lambda: ops.lk_branches_backward_data(dy, dy, dy, w1, w2, w3))
        report(f"lk3_dgrad_{tag}", us, b, 4 * e, 2 * f1 + f3)
        us, b = timeit(lambda: ops.lk_branches_backward_filter(x, dy, dy, dy, KL, 5))
        report(f"lk3_wgrad_{tag}", us, b, 4 * e, 2 * f1 + f3)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    shapes = [(128, 96, 56, 51)]
    if "--all-stages" in sys.argv:
        shapes += [(128, 192, 28, 49), (128, 384, 14, 47), (128, 768, 7, 13)]
    for (N, C, H, KL) in shapes:
        stage(N, C, H, KL, which)
