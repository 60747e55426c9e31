#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_optim_gpu.py -m gpu -q -s 2>&1 | grep -E "passed|failed|ulp|Error|assert" | head -20 | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_optim_gpu.py 2>&1 | tail -6 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; tail -c 400 gpurun_out/g_bench.err; python tools/show_bench.py gpurun_out/g_bench.json
