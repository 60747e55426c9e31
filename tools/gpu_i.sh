#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_slak_tiny_step_gpu.py tests/test_mlp_gpu.py tests/test_merge_decom.py tests/test_syncbn_2rank_gpu.py -m gpu -q -x 2>&1 | tail -5 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-ext --no-cpu-baseline > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; tail -c 400 gpurun_out/i_bench.err; python tools/show_bench.py gpurun_out/i_bench.json | head -3
bash tools/gpu_cfg.sh
