#!/bin/bash
# bench lines of the other BASELINE configs at N = 1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for CFG in 3 4 5; do
  STEPS=10; if [ $CFG = 5 ]; then STEPS=200; fi
  timeout 900 python bench.py --config $CFG --steps $STEPS --warmup 3 --no-ref-ext > gpurun_out/cfg${CFG}_bench.json 2> gpurun_out/cfg${CFG}_bench.err
  tail -c 300 gpurun_out/cfg${CFG}_bench.err
  python tools/show_bench.py gpurun_out/cfg${CFG}_bench.json | head -3
done
