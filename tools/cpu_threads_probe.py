"""Host-side probe: how many threads does the CPU arm of bench.py really get on this box, and what thread
count is fastest for one fwd+bwd+AdamW step of the oracle model?  (CPU only; run on the GPU box once.)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except OSError as e:
    print("cpu.max", e)
print("host_cores", bench.host_cores(), "OMP_NUM_THREADS", os.environ.get("OMP_NUM_THREADS"))
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for th in (8, 16, 32, 64):
    os.environ["SLAK_CPU_THREADS"] = str(th)
    t0 = time.perf_counter()
    ips, cores, sps = bench.time_cpu(1.0, batch, 1, 1)
    print(f"threads {th}: {ips:.3f} img/s  {sps:.2f} s/step (wall incl. warm-up {time.perf_counter() - t0:.1f} s)", flush=True)
    if time.perf_counter() - t0 > 60:
        break
