#!/bin/bash
# downsampling layers on own kernels: parity, model tests, bench line with down_* rows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_downsample_gpu.py tests/test_glue_v2_gpu.py -m gpu -q -x 2>&1 | tail -15 | cut -c1-300
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_slak_tiny_step_gpu.py -m gpu -q -x 2>&1 | tail -6 | cut -c1-300
for FD in 1 0; do
  SLAK_FUSED_DOWNSAMPLE=$FD timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-ext --no-cpu-baseline > gpurun_out/m_bench_$FD.json 2> gpurun_out/m_bench.err; tail -c 300 gpurun_out/m_bench.err
  python tools/show_bench.py gpurun_out/m_bench_$FD.json > gpurun_out/m_show_$FD.txt; head -1 gpurun_out/m_show_$FD.txt; grep down_ gpurun_out/m_show_$FD.txt
done
