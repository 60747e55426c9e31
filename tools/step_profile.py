"""Per-kernel device-time breakdown of one SLaK-T training step (torch.profiler / CUPTI, eager launches).
Prints the kernels sorted by total time; used to decide what to optimise next.  Not a bench: the profiler
serialises nothing but adds host overhead, so only the per-kernel device times are meaningful."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from slak_b200 import slak  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--top", type=int, default=45)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    net = slak.SLaK_tiny(kernel_size=[51, 49, 47, 13, 5], Decom=True, bn=True, drop_path_rate=0.1).to(dev).train()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=0.05, fused=True, capturable=True)
    x = torch.randn(a.batch, 3, 224, 224, device=dev)
    y = torch.randint(0, 1000, (a.batch,), device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.cross_entropy(net(x).float(), y)
        loss.backward()
        opt.step()

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = e.cuda_time_total
        if t > 0:
            rows.append((t / a.steps, e.count / a.steps, e.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"total device time per step: {tot / 1e3:.2f} ms over {sum(r[1] for r in rows):.0f} kernels")
    acc = 0.0
    for t, c, k in rows[:a.top]:
        acc += t
        print(f"{t / 1e3:8.3f} ms {100 * t / tot:5.1f}% (cum {100 * acc / tot:5.1f}%) x{c:6.1f}  {k[:110]}")


if __name__ == "__main__":
    main()
