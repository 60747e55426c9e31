#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for FL in 1 4; do
  SLAK_RES_FLAT_LW=$FL timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-ext --no-cpu-baseline > gpurun_out/x_bench_$FL.json 2> gpurun_out/x_bench.err; tail -c 200 gpurun_out/x_bench.err
  python tools/show_bench.py gpurun_out/x_bench_$FL.json > gpurun_out/x_show_$FL.txt; head -1 gpurun_out/x_show_$FL.txt | cut -c1-120; grep "glue_res_fwd\|down_out_fwd" gpurun_out/x_show_$FL.txt
done
SLAK_RES_FLAT_LW=4 timeout 300 python -m pytest tests/test_glue_v2_gpu.py tests/test_downsample_gpu.py -m gpu -q -x 2>&1 | tail -2
exit 0
