#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/dense_bench.py 2>&1 | tee gpurun_out/q_dense.txt | tail -12
timeout 300 python -m pytest tests/test_dense_planes_gpu.py -m gpu -q -x --timeout 60 2>&1 | tail -6 | cut -c1-300
exit 0
