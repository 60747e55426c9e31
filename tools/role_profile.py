import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slak_b200 import ops
for (N, C, H, KL) in [(128, 96, 56, 51), (128, 384, 14, 47), (128, 768, 7, 13)]:
    x = torch.randn(N, C, H, H, device="cuda").bfloat16()
    w1 = torch.randn(C, 1, KL, 5, device="cuda") * 0.02
    w2 = torch.randn(C, 1, 5, KL, device="cuda") * 0.02
    w3 = torch.randn(C, 1, 5, 5, device="cuda") * 0.02
    print(f"stage {H}x{H}", flush=True)
    ops.lk_branches_forward(x, w1, w2, w3)
    torch.cuda.synchronize()
