#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_fullsize_gpu.py tests/test_lk_branches_gpu.py tests/test_model_gpu.py tests/test_mlp_gpu.py -m gpu -q -x 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-ext --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; tail -c 300 gpurun_out/d_bench.err; python tools/show_bench.py gpurun_out/d_bench.json
