#!/bin/bash
# instruction-rate probe + ncu launch list of one eager training step (what is left outside this repo's kernels)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 60 tools/probes/fhfma_probe | tee gpurun_out/k_fhfma.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/k_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-ref-ext > gpurun_out/k_bench.log 2>&1
tail -c 400 gpurun_out/k_bench.log
gzip -f gpurun_out/k_launches.csv
python tools/launch_list_summary.py gpurun_out/k_launches.csv.gz 45
