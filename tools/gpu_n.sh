#!/bin/bash
# stem on own kernels: parity, model tests, bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_downsample_gpu.py -m gpu -q -x 2>&1 | tail -15 | cut -c1-300
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_slak_tiny_step_gpu.py -m gpu -q -x 2>&1 | tail -6 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-ext --no-cpu-baseline > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; tail -c 300 gpurun_out/n_bench.err
python tools/show_bench.py gpurun_out/n_bench.json > gpurun_out/n_show.txt; head -1 gpurun_out/n_show.txt; grep "stem_\|down_" gpurun_out/n_show.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/n_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-ref-ext > gpurun_out/n_ncu_bench.log 2>&1
gzip -f gpurun_out/n_launches.csv
python tools/launch_list_summary.py gpurun_out/n_launches.csv.gz 40
exit 0
