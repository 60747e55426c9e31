#!/usr/bin/env python
"""bench.py -- SLaK-T 51x51 224^2 bf16 training throughput (images/s) on N B200s, plus the
depthwise-kernel HBM roofline and the reference's CPU path timed on the same box.

  python bench.py [--gpus N --steps K --warmup W]            # this repo's CUDA path
  python bench.py --impl reference [...]                      # the reference's CPU (nn.Conv2d) path
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

One JSON line on stdout (rank 0).  A "step" is one fwd+bwd+AdamW pass of the hot path over
one synthetic batch.  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch
import torch.nn.functional as F

KERNEL_SIZE = [51, 49, 47, 13, 5]
DEPTHS = [3, 3, 9, 3]
PER_GPU_BATCH = 128
IMG = 224
NUM_CLASSES = 1000
HEADLINE = dict(N=PER_GPU_BATCH, C=96, H=56, W=56, kh=51, kw=5)   # the north-star kernel


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=20)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="per-GPU batch")
    p.add_argument("--width-factor", type=float, default=1.0)
    p.add_argument("--cpu-batch", type=int, default=32, help="images per CPU-baseline step")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--watchdog", type=float, default=1500.0, help="abort the process after this many seconds")
    p.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    return p.parse_args()


# ---------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi fields through NVML), runs only during the timed region
# ---------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag = [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = None

    def _run(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        if self.nv is not None:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()

    def stop(self):
        self.stop_flag.set()
        if self.t is not None:
            self.t.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------
# the reference's CPU path (oracle restatement, fp32, nn.Conv2d semantics) -- checker / baseline only
# ---------------------------------------------------------------------------------------
def host_cores():
    """Threads the CPU arm may really use: the scheduler affinity, clipped by the cgroup CPU quota when the
    container has one (spinning 128 OpenMP threads on a smaller quota is what makes a CPU run crawl), or
    SLAK_CPU_THREADS when set."""
    if os.environ.get("SLAK_CPU_THREADS"):
        return max(1, int(os.environ["SLAK_CPU_THREADS"]))
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, min(n, 64))


def cpu_training_step_factory(width_factor, batch):
    """Returns (step_fn, cores): one fwd+bwd+AdamW step of SLaK-T on the host cores through the
    oracle's functional restatement of models/SLaK.py (F.conv2d depthwise, train-mode BN)."""
    from oracle import slak_model as omodel
    from slak_b200 import slak
    cores = host_cores()
    torch.set_num_threads(cores)     # explicit: torchrun exports OMP_NUM_THREADS=1 to its workers
    torch.manual_seed(0)
    slak.use_sync_bn = False
    net = slak.SLaK_tiny(kernel_size=KERNEL_SIZE, Decom=True, bn=True, drop_path_rate=0.0,
                         width_factor=width_factor, num_classes=NUM_CLASSES)
    sd = {}
    leaves = []
    for k, v in net.state_dict().items():
        t = v.detach().clone()
        if v.dtype.is_floating_point and "running_" not in k:
            t.requires_grad_(True)
            leaves.append(t)
        sd[k] = t
    opt = torch.optim.AdamW(leaves, lr=1e-3, weight_decay=0.05)
    x = torch.randn(batch, 3, IMG, IMG)
    y = torch.randint(0, NUM_CLASSES, (batch,))

    def step():
        out = omodel.forward(x, sd, DEPTHS, training=True)
        loss = F.cross_entropy(out, y)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss.item()

    return step, cores


def time_cpu(width_factor, batch, steps, warmup):
    step, cores = cpu_training_step_factory(width_factor, batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, cores, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 6))
    warmup = max(1, min(args.warmup, 1))
    ips, cores, sps = time_cpu(args.width_factor, args.cpu_batch, steps, warmup)
    sample = f"{steps} steps x {args.cpu_batch} images of the same SLaK-T 224^2 fwd+bwd+AdamW step, fp32, {cores} threads"
    line = {
        "impl": "reference", "metric": "SLaK-T 51x51 224x224 training images/sec", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    return {
        "workload": f"SLaK-T kernel_size={KERNEL_SIZE} Decom=True bn=True width_factor={args.width_factor} "
                    f"224x224, fwd+bwd+AdamW, batch {args.batch}/GPU (BASELINE.json configs[1])",
        "global_batch": args.batch * n, "per_gpu_batch": args.batch, "parallelism": f"dp{n}",
        "autocast": "bf16 (fp32 master weights, fp32 residual stream as in the reference's AMP flow)",
        "l2": "no explicit flush: one step streams >10 GB of activations, far above the 126 MB L2",
    }


# ---------------------------------------------------------------------------------------
# this repo's CUDA path
# ---------------------------------------------------------------------------------------
def run_ours(args):
    from slak_b200 import _lib, ops, slak
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if not _lib.load().slak_device_ok():
        raise SystemExit("libslak_b200.so: no sm_100 device visible")
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True

    torch.manual_seed(0 + rank)    # main.py:232 seeds seed + rank
    slak.use_sync_bn = world > 1
    net = slak.SLaK_tiny(kernel_size=KERNEL_SIZE, Decom=True, bn=True, drop_path_rate=0.1,
                         width_factor=args.width_factor, num_classes=NUM_CLASSES).to(dev)
    net.train()
    params = [p for p in net.parameters()]
    if world > 1:                      # identical initial weights on every rank (what DDP's constructor does)
        for p in params:
            dist.broadcast(p.data, src=0)
        for b_ in net.buffers():
            dist.broadcast(b_, src=0)
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0.05, fused=True, capturable=True)

    B = args.batch
    x_host = torch.randn(B, 3, IMG, IMG).pin_memory()
    y_host = torch.randint(0, NUM_CLASSES, (B,)).pin_memory()
    x_dev = x_host.to(dev)             # static input buffers (also the CUDA-graph inputs)
    y_dev = y_host.to(dev)

    def allreduce_grads():
        """Data parallelism = gradient all-reduce only (main.py:374-376 wraps DDP for the same effect): one NCCL
        all-reduce over NVLink of the flattened fp32 gradients, averaged."""
        grads = [p.grad for p in params if p.grad is not None]
        flat = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(flat)
        flat.div_(world)
        for g_, f_ in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
            g_.copy_(f_)

    def step_eager():
        # optimizer.zero_grad() as in engine.py:74-86 (set_to_none is the torch default): backward then writes fresh gradients instead of
        # accumulating into zeroed ones; under graph capture they live in the graph's private pool
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(x_dev)
            loss = F.cross_entropy(out.float(), y_dev)
        loss.backward()
        if world > 1:
            allreduce_grads()
        opt.step()
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (side stream, as CUDA-graph capture requires), then capture the whole step -------------
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(args.warmup, 3)):
            step_eager()
    torch.cuda.current_stream().wait_stream(side)
    barrier()

    graph, static_loss, graph_note = None, None, "eager (no CUDA graph)"
    prof_events = []
    use_graph = not args.no_graph
    if use_graph:
        try:
            ops.profile_reset(dict(HEADLINE, external_events=True))
            l_before = ops.launch_count()
            graph = torch.cuda.CUDAGraph()
            # thread_local: the NCCL watchdog thread's event queries must not invalidate this thread's capture
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                static_loss = step_eager()
            prof_events = list(ops._prof["events"])
            launches_per_step = ops.launch_count() - l_before
            ops.profile_reset(None)
            graph_note = "whole step (fwd+bwd+grad all-reduce+AdamW) captured in one CUDA graph and replayed"
        except Exception as ex:      # capture not possible on this software stack: fall back to eager launches
            graph, static_loss = None, None
            ops.profile_reset(None)
            torch.cuda.synchronize()
            graph_note = f"eager (CUDA graph capture failed: {type(ex).__name__})"
            print(f"[bench] rank {rank}: graph capture failed: {ex!r}", file=sys.stderr, flush=True)
        if world > 1:                # every rank must replay, or none: the captured collectives have to pair up
            ok = torch.tensor([1 if graph is not None else 0], device=dev, dtype=torch.int32)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0 and graph is not None:
                graph, static_loss = None, None
                graph_note = "eager (CUDA graph capture failed on another rank)"

    def run_step():
        if graph is not None:
            graph.replay()
            return static_loss
        return step_eager()

    for _ in range(3):
        run_step()
    barrier()

    # ---- timed region 1: inputs resident in HBM ------------------------------------------
    if graph is None:
        ops.profile_reset(HEADLINE)
    sampler = ClockSampler(local_rank)
    launches0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    e0.record()
    for _ in range(args.steps):
        run_step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if graph is None:
        launches = ops.launch_count() - launches0
        prof = ops.profile_collect()
        ops.profile_reset(None)
        prof_how = "CUDA events around each launch of the kernel inside the timed region (eager launches)"
    else:
        launches = launches_per_step * args.steps      # kernels of this repo inside each replayed step graph
        # the events sit inside the captured graph: read them after the last timed replay and after a few more
        tot, cnt = 0.0, 0
        for rep in range(4):
            if rep:
                run_step()
                torch.cuda.synchronize()
            for a_, b_ in prof_events:
                tot += a_.elapsed_time(b_)
                cnt += 1
        prof = {"count": cnt, "ms_total": tot}
        prof_how = ("CUDA events (external) recorded around the kernel INSIDE the captured step graph; read after the "
                    "last timed replay and 3 further replays")

    # ---- timed region 2: end to end through the public API with host buffers --------------
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        x_dev.copy_(x_host, non_blocking=True)       # pinned host -> device, every step
        y_dev.copy_(y_host, non_blocking=True)
        loss = run_step()
        loss_host = loss.item()                      # device -> host read of the step's result
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    total_images = B * world * args.steps
    value = total_images / (ms / 1e3)
    e2e = total_images / (ms_e2e / 1e3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        h = HEADLINE
        # fused three-branch forward: x read once, y1/y2/y3 written once (bf16) + the fp32 taps
        alg_bytes = 4 * h["N"] * h["C"] * h["H"] * h["W"] * 2 + h["C"] * (2 * h["kh"] * h["kw"] + 25) * 4
        roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                "kernel": "lk3_fwd_tc_kernel<64,16,TMA> (stage-1 fused 51x5 + 5x51 + 5x5 forward, tcgen05)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launches_timed": prof["count"]}
        try:    # DRAM bytes of one launch of this kernel from the committed `ncu --set full` capture (not measured live)
            tr = json.load(open(os.path.join(ROOT, "profiles", "headline_traffic.json")))
            if B == h["N"]:
                roof["traffic"] = tr["dram_bytes"]
                roof["traffic_source"] = f"{tr['source']} (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
        except Exception:
            pass
        if prof["count"] > 0 and B == h["N"]:
            us = prof["ms_total"] * 1e3 / prof["count"]
            roof["avg_us"] = us
            roof["achieved"] = alg_bytes / (us * 1e-6) / 1e9
            roof["frac"] = roof["achieved"] / peak
            roof["share_of_step"] = (us * 1e-3 * 3) / (ms / args.steps)     # 3 stage-1 Blocks per step
            roof["timing"] = prof_how
        line = {
            "metric": "SLaK-T 51x51 224x224 bf16 training images/sec", "value": value, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(args, world), launch=graph_note),
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "images/s",
                    "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8, "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "roofline": roof,
        }
        if world == 1 and not args.no_cpu_baseline:
            ips, cores, _ = time_cpu(args.width_factor, args.cpu_batch, 3, 1)
            line["cpu_baseline"] = {
                "value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                "sample": f"3 steps x {args.cpu_batch} images of the same SLaK-T 224^2 fwd+bwd+AdamW step through "
                          f"oracle/slak_model.py (F.conv2d depthwise, fp32), {cores} threads"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        # tear down without ncclCommDestroy: with the step graph (which holds the captured all-reduces) alive the
        # communicator teardown blocks forever; the process exits right after, which frees everything
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        if graph is not None:
            graph.reset()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    a = parse()
    # a run that cannot finish (a wedged collective, a dead peer rank) must not hold the GPU box forever
    _wd = threading.Timer(a.watchdog, lambda: (sys.stderr.write("[bench] watchdog: no result in time, aborting\n"), os._exit(3)))
    _wd.daemon = True
    _wd.start()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
