#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mlp_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 400 python tools/mlp_bench.py 2>&1 | tail -30 | cut -c1-150
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-ext --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; tail -c 300 gpurun_out/c_bench.err; python tools/show_bench.py gpurun_out/c_bench.json
