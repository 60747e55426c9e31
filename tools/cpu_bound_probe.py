"""Is the training step launch-bound?  Host enqueue time vs device time for the bench step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from slak_b200 import slak
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda")
slak.use_sync_bn = False
net = slak.SLaK_tiny(kernel_size=[51, 49, 47, 13, 5], Decom=True, bn=True, drop_path_rate=0.1, num_classes=1000).to(dev).train()
opt = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=0.05, fused=True)
x = torch.randn(128, 3, 224, 224, device=dev); y = torch.randint(0, 1000, (128,), device=dev)
def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = F.cross_entropy(net(x).float(), y)
    loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/10:.1f} ms/step, total {1e3*(t2-t0)/10:.1f} ms/step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:25]
tot = sum(e.device_time_total for e in prof.key_averages())
print(f"device time {tot/3/1e3:.1f} ms/step")
for e in rows:
    print(f"{e.device_time_total/3/1e3:7.2f} ms {e.count//3:5d}x {e.key[:90]}")
