#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/dense_bench.py 2>&1 | tee gpurun_out/s_dense.txt | tail -8
timeout 300 python -m pytest tests/test_dense_planes_gpu.py -m gpu -q -x --timeout 60 2>&1 | tail -4 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-ext --no-cpu-baseline > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; tail -c 300 gpurun_out/s_bench.err
python tools/show_bench.py gpurun_out/s_bench.json > gpurun_out/s_show.txt; head -1 gpurun_out/s_show.txt; grep "dw_" gpurun_out/s_show.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-ref-ext > gpurun_out/s_ncu_bench.log 2>&1
gzip -f gpurun_out/s_launches.csv
python tools/launch_list_summary.py gpurun_out/s_launches.csv.gz 60
exit 0
