#!/usr/bin/env python
"""Run every single-convolution op (fwd / dgrad / wgrad, bf16) at the four SLaK-T stage shapes in its own subprocess
with a timeout, and print its duration: finds a hanging or pathologically slow shape without losing the whole run."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from slak_b200 import ops
C, HW, kh, kw, op, N = [int(a) for a in sys.argv[1:7]]
x = torch.randn(N, C, HW, HW, device="cuda").bfloat16()
g = torch.randn(N, C, HW, HW, device="cuda").bfloat16()
w = torch.randn(C, 1, kh, kw, device="cuda") * 0.02
fn = [lambda: ops.dwconv2d_forward(x, w), lambda: ops.dwconv2d_backward_data(g, w), lambda: ops.dwconv2d_backward_filter(g, x, w)][op]
fn(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fn(); e1.record(); torch.cuda.synchronize()
print("%%.1f us" %% (e0.elapsed_time(e1) * 1e3))
''' % ROOT

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for C, HW, K in ((96, 56, 51), (192, 28, 49), (384, 14, 47), (768, 7, 13)):
    for kh, kw in ((K, 5), (5, K), (5, 5)):
        for op, name in enumerate(("fwd", "dgrad", "wgrad")):
            try:
                r = subprocess.run([sys.executable, "-c", CHILD, str(C), str(HW), str(kh), str(kw), str(op), str(N)],
                                   capture_output=True, text=True, timeout=40)
                out = (r.stdout.strip() or r.stderr.strip()[-200:])
            except subprocess.TimeoutExpired:
                out = "TIMEOUT"
            print(f"C{C} {HW}x{HW} {kh}x{kw} {name}: {out}", flush=True)
