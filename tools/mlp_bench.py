#!/usr/bin/env python
"""Microbenchmark of the tcgen05 MLP GEMMs (csrc/mlp_tc.cu) at the four SLaK-T stage shapes (batch 128), each kernel
alone with the L2 flushed between launches, next to the torch (cuBLAS) expression it replaces.
Prints one line per (stage, kernel): microseconds, algorithmic GB/s and TFLOP/s."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

from slak_b200 import _lib

L = _lib
lib = _lib.load()
DEV = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, reps=7):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def p(t):
    return None if t is None else t.data_ptr()


def nt(epi, a, b, bias, aux, o0, o1, cp, M, N, K):
    L.check(lib.slak_mlp_gemm_nt(epi, p(a), p(b), p(bias), p(aux), p(o0), p(o1), p(cp), M, N, K, L.current_stream_ptr()), "nt")


def main():
    only = sys.argv[1:] or None
    rows = []
    for C, HW in ((96, 56), (192, 28), (384, 14), (768, 7)):
        M, H4 = 128 * HW * HW, 4 * C
        bf = torch.bfloat16
        x = torch.randn(M, C, device=DEV).to(bf)
        W1 = (torch.randn(H4, C, device=DEV) * C ** -0.5).to(bf)
        W2 = (torch.randn(C, H4, device=DEV) * H4 ** -0.5).to(bf)
        W1t, W2t = W1.t().contiguous(), W2.t().contiguous()
        b1, b2 = torch.randn(H4, device=DEV), torch.randn(C, device=DEV)
        h, a = torch.empty(M, H4, device=DEV, dtype=bf), torch.empty(M, H4, device=DEV, dtype=bf)
        h2 = torch.empty(M, C, device=DEV, dtype=bf)
        dh2 = torch.randn(M, C, device=DEV).to(bf)
        dh = torch.empty(M, H4, device=DEV, dtype=bf)
        dxn = torch.empty(M, C, device=DEV, dtype=bf)
        cp = torch.empty(lib.slak_mlp_parts(M, H4), H4, device=DEV)
        s2, s1 = lib.slak_mlp_wgrad_splits(M, C, H4), lib.slak_mlp_wgrad_splits(M, H4, C)
        part2 = torch.empty(s2, C * H4, device=DEV)
        part1 = torch.empty(s1, H4 * C, device=DEV)
        e = 2
        cases = [
            ("fc1+gelu", lambda: nt(0, x, W1, b1, None, h, a, None, M, H4, C), (M * C + 2 * M * H4) * e, 2 * M * C * H4,
             lambda: F.gelu(torch.addmm(b1.to(bf), x, W1.t()))),
            ("fc2+bias", lambda: nt(1, a, W2, b2, None, h2, None, None, M, C, H4), (M * H4 + M * C) * e, 2 * M * C * H4,
             lambda: torch.addmm(b2.to(bf), a, W2.t())),
            ("dgelu", lambda: nt(2, dh2, W2t, None, h, dh, None, cp, M, H4, C), (M * C + 2 * M * H4) * e, 2 * M * C * H4,
             lambda: torch.mm(dh2, W2)),
            ("dxn", lambda: nt(3, dh, W1t, None, None, dxn, None, None, M, C, H4), (M * H4 + M * C) * e, 2 * M * C * H4,
             lambda: torch.mm(dh, W1)),
            ("dW2", lambda: L.check(lib.slak_mlp_gemm_tn_splitk(p(dh2), p(a), p(part2), M, C, H4, L.current_stream_ptr()), "w"),
             (M * C + M * H4) * e, 2 * M * C * H4, lambda: torch.mm(dh2.t(), a)),
            ("dW1", lambda: L.check(lib.slak_mlp_gemm_tn_splitk(p(dh), p(x), p(part1), M, H4, C, L.current_stream_ptr()), "w"),
             (M * C + M * H4) * e, 2 * M * C * H4, lambda: torch.mm(dh.t(), x)),
        ]
        for name, fn, byts, flops, ref in cases:
            if only and name not in only:
                continue
            us = timeit(fn)
            us_ref = timeit(ref)
            rows.append(dict(stage=f"C{C} M{M}", kernel=name, us=round(us, 1), gbs=round(byts / us / 1e3, 1),
                             tflops=round(flops / us / 1e6, 1), torch_us=round(us_ref, 1)))
            print(rows[-1], flush=True)
    out = os.path.join(ROOT, "gpurun_out", "mlp_bench.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(rows, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
