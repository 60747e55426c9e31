#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/dbg_optim.py 2>&1 | tail -12
timeout 900 python -m pytest tests/test_merge_decom.py tests/test_masking_gpu.py tests/test_optim_gpu.py tests/test_mlp_gpu.py -m gpu -q 2>&1 | grep -v "Warning\|warn\|cosine_stepper\|^$" | tail -30 | cut -c1-400
timeout 300 python tools/mlp_bench.py dgelu fc1+gelu 2>&1 | tail -10 | cut -c1-150
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-ext --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; tail -c 400 gpurun_out/f_bench.err; python tools/show_bench.py gpurun_out/f_bench.json
