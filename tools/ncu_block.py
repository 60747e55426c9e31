"""One SLaK Block per stage geometry, forward + backward under bf16 autocast: the launch sequence ncu captures
for the per-kernel profiles under profiles/ (run under `ncu -k regex:slak --launch-skip S --launch-count C`).
usage: python tools/ncu_block.py [iters] [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from slak_b200 import ops, slak  # noqa: E402

STAGES = [(96, 56, 51), (192, 28, 49), (384, 14, 47), (768, 7, 13)]


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    blocks, xs = [], []
    for dim, hw, kl in STAGES:
        blocks.append(slak.Block(dim, drop_path=0.0, kernel_size=(kl, 5), Decom=True, bn=True).to(dev).train())
        xs.append(torch.randn(batch, dim, hw, hw, device=dev, requires_grad=True))
    for it in range(iters):
        if it == iters - 1:
            torch.cuda.profiler.start()      # ncu --profile-from-start off: only the last iteration is captured
        n0 = ops.launch_count()
        for blk, x in zip(blocks, xs):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = blk(x)
            y.backward(torch.ones_like(y))
        torch.cuda.synchronize()
        if it == iters - 1:
            torch.cuda.profiler.stop()
        print(f"iteration {it}: {ops.launch_count() - n0} launches of this repo's kernels", flush=True)


if __name__ == "__main__":
    main()
