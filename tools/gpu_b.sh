#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== mlp" ; timeout 600 python -m pytest tests/test_mlp_gpu.py tests/test_syncbn_2rank_gpu.py -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/b_mlp.txt
echo "== model" ; timeout 600 python -m pytest tests/test_model_gpu.py tests/test_slak_tiny_step_gpu.py -m gpu -q -s 2>&1 | grep -E "passed|failed|worst|Error|error" | cut -c1-1500 | tee gpurun_out/b_model.txt
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-ext > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; tail -c 600 gpurun_out/b_bench.err; python tools/show_bench.py gpurun_out/b_bench.json
