"""bench.py contract checks that need no GPU: the CPU arm (`--impl reference`) prints ONE JSON line with the keys the
driver reads, rank != 0 exits silently, the thread count honours the cgroup quota, and the GPU arm refuses to run
without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"]


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, env=e, cwd=ROOT)


@pytest.mark.timeout(900)
def test_reference_arm_prints_one_contract_line():
    # width 0.25 and one image keep the CPU work to a few seconds; the keys and their meaning do not depend on it
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-batch", "1", "--width-factor", "0.25"],
             env={"SLAK_CPU_THREADS": "4"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 4
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # the driver forms the ours/reference ratio only when both arms print the SAME metric string, and checks steps/warmup
    sys.path.insert(0, ROOT)
    import bench
    assert d["metric"] == bench.METRIC == bench.CONFIGS[2]["metric"]
    assert d["steps"] == 1 and d["warmup"] == 1                      # honoured, not silently capped
    assert d["config"]["cpu_images_per_step"] == 1 and "1 images" in d["cpu_baseline"]["sample"]


def test_every_config_names_the_baseline_entry_and_shares_metric_between_arms():
    sys.path.insert(0, ROOT)
    import bench
    assert sorted(bench.CONFIGS) == [2, 3, 4, 5]
    for k, c in bench.CONFIGS.items():
        assert f"configs[{k - 1}]" in c["what"] and c["metric"].endswith("images/sec") or "images/sec" in c["metric"]
    assert bench.CONFIGS[5]["kernel_size"][0] == 61 and bench.CONFIGS[5]["sparse"]
    assert bench.CONFIGS[4]["img"] == 384 and bench.CONFIGS[4]["batch"] == 32
    assert bench.CONFIGS[3]["update_freq"] == 4


def test_reference_arm_other_ranks_do_nothing():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"},
             timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_host_cores_honours_override_and_quota(monkeypatch, tmp_path):
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setenv("SLAK_CPU_THREADS", "3")
    assert bench.host_cores() == 3
    monkeypatch.delenv("SLAK_CPU_THREADS")
    n = bench.host_cores()
    assert 1 <= n <= 64 and n <= len(os.sched_getaffinity(0))


def test_gpu_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "1"], timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)
