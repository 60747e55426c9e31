#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/dense_bench.py 2>&1 | tee gpurun_out/r_dense.txt | tail -12
timeout 300 python -m pytest tests/test_dense_planes_gpu.py -m gpu -q -x --timeout 60 2>&1 | tail -6 | cut -c1-300
timeout 600 python -m pytest tests/test_tc_fullsize_gpu.py tests/test_model_gpu.py tests/test_slak_tiny_step_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -5 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-ext --no-cpu-baseline > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err; tail -c 300 gpurun_out/r_bench.err
python tools/show_bench.py gpurun_out/r_bench.json > gpurun_out/r_show.txt; head -1 gpurun_out/r_show.txt; grep "dw_" gpurun_out/r_show.txt
exit 0
