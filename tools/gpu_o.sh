#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 60 tools/probes/tma_unaligned_probe 2>&1 | tee gpurun_out/o_tma_probe.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-ext --no-cpu-baseline > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err; tail -c 300 gpurun_out/o_bench.err
python tools/show_bench.py gpurun_out/o_bench.json > gpurun_out/o_show.txt; head -1 gpurun_out/o_show.txt
timeout 600 python -m pytest tests/test_optim_gpu.py tests/test_slak_tiny_step_gpu.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-300
exit 0
