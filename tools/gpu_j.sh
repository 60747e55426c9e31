#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_slak_tiny_step_gpu.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-300
for KB in 24 48; do
  SLAK_GLUE_TILE_KB=$KB timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-ext --no-cpu-baseline > gpurun_out/j_bench_$KB.json 2> gpurun_out/j_bench.err; tail -c 300 gpurun_out/j_bench.err
  python tools/show_bench.py gpurun_out/j_bench_$KB.json > gpurun_out/j_show_$KB.txt; head -1 gpurun_out/j_show_$KB.txt; grep glue gpurun_out/j_show_$KB.txt
done
