#!/bin/bash
# dense small-plane kernels: parity first (each case in its own timeout), then model tests and the bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dense_planes_gpu.py -m gpu -q -x --timeout 60 2>&1 | tail -25 | cut -c1-300
timeout 600 python -m pytest tests/test_tc_fullsize_gpu.py tests/test_model_gpu.py tests/test_slak_tiny_step_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -8 | cut -c1-300
for DP in 1 0; do
  SLAK_DENSE_PLANES=$DP timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-ext --no-cpu-baseline > gpurun_out/p_bench_$DP.json 2> gpurun_out/p_bench.err; tail -c 300 gpurun_out/p_bench.err
  python tools/show_bench.py gpurun_out/p_bench_$DP.json > gpurun_out/p_show_$DP.txt; head -1 gpurun_out/p_show_$DP.txt; grep "dw_" gpurun_out/p_show_$DP.txt
done
exit 0
