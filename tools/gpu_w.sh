#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_glue_v2_gpu.py tests/test_downsample_gpu.py tests/test_model_gpu.py tests/test_slak_tiny_step_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -5 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-ext --no-cpu-baseline > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err; tail -c 300 gpurun_out/w_bench.err
python tools/show_bench.py gpurun_out/w_bench.json > gpurun_out/w_show.txt; head -1 gpurun_out/w_show.txt; grep "HW49" gpurun_out/w_show.txt
exit 0
