#!/usr/bin/env python
"""Time the reference's OWN CUDA operator on this GPU -- the MegEngine-CUTLASS example-19 torch extension built for
sm_100a by oracle/build_ref_ext.py (oracle/_ref/ext/_depthwise_conv2d_implicit_gemm_C.so) -- next to this repo's
kernels: north_star's comparison target ("the reference CUTLASS-ext build on the same B200 box").

BASELINE / TEST INFRASTRUCTURE: nothing under slak_b200/ imports this.  /root/reference does not exist on the GPU
box, so the reference's 82-line Python module (depthwise_conv2d_implicit_gemm.py) is not imported; the binding below
calls the extension's six functions (frontend.h:3-10) the way that module does (fp32 path for fp32 input, fp16 path
for fp16 input, dw returned in fp32) and stock PyTorch modules do everything else, as in models/SLaK.py.

  python tools/ref_ext_bench.py --ops            per-op table: fwd / dgrad / wgrad, K x 5 | 5 x K | 5 x 5, 4 stages,
                                                 fp32 + fp16 (reference ext) vs bf16 (this repo), microseconds
  python tools/ref_ext_bench.py --model-only     whole-model training step (fwd+bwd+AdamW) images/s, reference ext
                                                 under fp16 autocast + GradScaler-free fp32 master (its AMP flow) and fp32
Prints one JSON line (last line of stdout) and, with --out, writes a markdown table.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "ext"))

import torch
import torch.nn as nn
import torch.nn.functional as F


def load_ext():
    import _depthwise_conv2d_implicit_gemm_C as ext     # noqa: F401  (needs torch imported first)
    return ext


class _RefDW(torch.autograd.Function):
    """forward_* / backward_data_* / backward_filter_* of the reference extension by input dtype."""

    @staticmethod
    def forward(ctx, x, w):
        ext = load_ext()
        half = x.dtype == torch.float16
        w_ = w.half() if half else w
        ctx.save_for_backward(x, w_)
        ctx.half = half
        return ext.forward_fp16(x, w_) if half else ext.forward_fp32(x, w_)

    @staticmethod
    def backward(ctx, g):
        ext = load_ext()
        x, w = ctx.saved_tensors
        g = g.contiguous()
        if ctx.half:
            return ext.backward_data_fp16(g, w), ext.backward_filter_fp16(g, x, w)
        return ext.backward_data_fp32(g, w), ext.backward_filter_fp32(g, x, w)


class RefDepthWiseConv2dImplicitGEMM(nn.Conv2d):
    def __init__(self, channels, kernel, bias=False):
        super().__init__(channels, channels, kernel, groups=channels, bias=bias)

    def forward(self, x):
        if x.dtype not in (torch.float32, torch.float16):
            raise TypeError("Only support fp32 and fp16, get {}".format(x.dtype))
        # the reference pins fp32 inputs to the fp32 kernel even under autocast (custom_fwd(cast_inputs=float32))
        with torch.autocast("cuda", enabled=False):
            y = _RefDW.apply(x.contiguous(), self.weight)
        return y


def _time(fn, reps, flush):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()                                   # > L2: the next launch starts cold
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


STAGES = [(96, 56, 51), (192, 28, 49), (384, 14, 47), (768, 7, 13)]


def run_ops(N, reps):
    from slak_b200 import ops
    ext = load_ext()
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for C, HW, KL in STAGES:
        x32 = torch.randn(N, C, HW, HW, device=dev)
        g32 = torch.randn(N, C, HW, HW, device=dev)
        for kh, kw in ((KL, 5), (5, KL), (5, 5)):
            w = torch.randn(C, 1, kh, kw, device=dev) * 0.02
            row = {"stage": f"C{C} {HW}x{HW}", "kernel": f"{kh}x{kw}", "N": N}
            for name, dt in (("fp32", torch.float32), ("fp16", torch.float16)):
                x, g, w_ = x32.to(dt), g32.to(dt), w.to(dt)
                f = {"fp32": (ext.forward_fp32, ext.backward_data_fp32, ext.backward_filter_fp32),
                     "fp16": (ext.forward_fp16, ext.backward_data_fp16, ext.backward_filter_fp16)}[name]
                row[f"ref_{name}_fwd_us"] = round(_time(lambda: f[0](x, w_), reps, flush), 1)
                row[f"ref_{name}_dgrad_us"] = round(_time(lambda: f[1](g, w_), reps, flush), 1)
                row[f"ref_{name}_wgrad_us"] = round(_time(lambda: f[2](g, x, w_), reps, flush), 1)
            xb, gb = x32.bfloat16(), g32.bfloat16()
            row["ours_bf16_fwd_us"] = round(_time(lambda: ops.dwconv2d_forward(xb, w), reps, flush), 1)
            row["ours_bf16_dgrad_us"] = round(_time(lambda: ops.dwconv2d_backward_data(gb, w), reps, flush), 1)
            row["ours_bf16_wgrad_us"] = round(_time(lambda: ops.dwconv2d_backward_filter(gb, xb, w), reps, flush), 1)
            row["ours_fp32_fwd_us"] = round(_time(lambda: ops.dwconv2d_forward(x32, w), reps, flush), 1)
            row["ours_fp32_dgrad_us"] = round(_time(lambda: ops.dwconv2d_backward_data(g32, w), reps, flush), 1)
            row["ours_fp32_wgrad_us"] = round(_time(lambda: ops.dwconv2d_backward_filter(g32, x32, w), reps, flush), 1)
            # parity of the two implementations on the same inputs (fp32: both are fp32-FMA kernels)
            y_ref, y_our = ext.forward_fp32(x32, w), ops.dwconv2d_forward(x32, w)
            row["fwd_fp32_max_rel_diff"] = float(((y_ref - y_our).abs().max() / y_ref.abs().max()).item())
            rows.append(row)
        # the fused three-branch kernels of this repo against the SUM of the reference's three launches
        ws = [torch.randn(C, 1, *k, device=dev) * 0.02 for k in ((KL, 5), (5, KL), (5, 5))]
        xb = x32.bfloat16()
        gs = [torch.randn(N, C, HW, HW, device=dev).bfloat16() for _ in range(3)]
        rows.append({"stage": f"C{C} {HW}x{HW}", "kernel": "fused 3 branches (tcgen05)", "N": N,
                     "ours_bf16_fwd_us": round(_time(lambda: ops.lk_branches_forward(xb, *ws), reps, flush), 1),
                     "ours_bf16_dgrad_us": round(_time(lambda: ops.lk_branches_backward_data(*gs, *ws), reps, flush), 1),
                     "ours_bf16_wgrad_us": round(_time(lambda: ops.lk_branches_backward_filter(xb, *gs, KL, 5), reps, flush), 1)})
    return rows


def run_model(config, batch, steps, amp):
    """SLaK under stock PyTorch modules (nn.BatchNorm2d, F.layer_norm, nn.Linear, nn.GELU: the module-by-module path of
    slak_b200.slak with FUSED_BLOCK off, i.e. models/SLaK.py) around the reference extension."""
    import bench
    from slak_b200 import slak
    cfg = bench.CONFIGS[config]
    slak.FUSED_BLOCK = False
    slak.use_sync_bn = False
    slak.DepthWiseConv2dImplicitGEMM = RefDepthWiseConv2dImplicitGEMM          # get_conv2d() builds this one
    # the channels_first LayerNorm of the stem / downsampling layers goes back to the reference's expression too
    def ln_forward(self, x):
        if self.data_format == "channels_last":
            return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        u = x.mean(1, keepdim=True)
        s = (x - u).pow(2).mean(1, keepdim=True)
        x = (x - u) / torch.sqrt(s + self.eps)
        return self.weight[:, None, None] * x + self.bias[:, None, None]
    slak.LayerNorm.forward = ln_forward
    orig_block_forward = slak.Block.forward

    def block_forward(self, x):           # models/SLaK.py:153-166, no autocast-eligible depthwise branch
        inp = x
        x = self.large_kernel(x)
        x = x.permute(0, 2, 3, 1)
        x = self.norm(x)
        x = self.pwconv2(self.act(self.pwconv1(x)))
        if self.gamma is not None:
            x = self.gamma * x
        x = x.permute(0, 3, 1, 2)
        return inp + self.drop_path(x)
    slak.Block.forward = block_forward

    def lk_forward(self, inputs):         # models/SLaK.py:89-100
        outs = [b(inputs) for b in self.branches()]
        out = outs[0]
        for o in outs[1:]:
            out = out + o
        return out
    slak.ReparamLargeKernelConv.forward = lk_forward
    torch.manual_seed(0)
    dev = "cuda"
    net = bench.build_model(cfg, 1.0, 0.1).to(dev).train()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=0.05)
    scaler = torch.amp.GradScaler("cuda", enabled=(amp == "fp16"))
    x = torch.randn(batch, 3, cfg["img"], cfg["img"], device=dev)
    y = torch.randint(0, 1000, (batch,), device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16, enabled=(amp == "fp16")):       # engine.py:53
            loss = F.cross_entropy(net(x).float(), y)
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        return loss

    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    slak.Block.forward = orig_block_forward
    return {"images_per_s": batch / (ms / 1e3), "ms_per_step": ms, "batch": batch, "steps": steps,
            "loss_finite": bool(torch.isfinite(loss).item())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", action="store_true")
    ap.add_argument("--model-only", action="store_true")
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--amp", default="both", choices=["fp16", "fp32", "both"])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    out = {"what": "reference CUTLASS example-19 extension (sm_100a build of the reference's own sources) under stock PyTorch",
           "gpu": torch.cuda.get_device_name(0)}
    if a.ops or not a.model_only:
        out["ops"] = run_ops(a.batch, a.reps)
    if a.model_only or not a.ops:
        out["model"] = {}
        for amp in (("fp16", "fp32") if a.amp == "both" else (a.amp,)):
            b = a.batch
            while True:
                try:
                    out["model"][amp] = run_model(a.config, b, a.steps, amp)
                    break
                except torch.OutOfMemoryError:
                    torch.cuda.empty_cache()
                    b //= 2
                    if b < 8:
                        out["model"][amp] = {"unavailable": "out of memory"}
                        break
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
