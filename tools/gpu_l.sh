#!/bin/bash
# glue v2: parity + A/B timing, then the model tests and the bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/glue_bench.py 2>&1 | tee gpurun_out/l_glue.txt | tail -120 | cut -c1-200
timeout 900 python -m pytest tests/test_glue_v2_gpu.py tests/test_model_gpu.py tests/test_slak_tiny_step_gpu.py -m gpu -q -x 2>&1 | tail -6 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-ext --no-cpu-baseline > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; tail -c 300 gpurun_out/l_bench.err
python tools/show_bench.py gpurun_out/l_bench.json > gpurun_out/l_show.txt; head -1 gpurun_out/l_show.txt; grep glue gpurun_out/l_show.txt
