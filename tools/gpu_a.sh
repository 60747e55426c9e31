#!/bin/bash
# round-2 GPU call A: new parity tests, MLP draft tests, reference ext timing, bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
echo "== mlp draft" ; SLAK_FUSED_MLP_TEST=1 timeout 300 python -m pytest tests/test_mlp_draft_gpu.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/a_mlp.txt
echo "== new tests" ; timeout 900 python -m pytest tests/test_tc_fullsize_gpu.py tests/test_slak_tiny_step_gpu.py tests/test_syncbn_2rank_gpu.py tests/test_ref_ext_gpu.py -m gpu -q -s 2>&1 | tail -40 | tee gpurun_out/a_new.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/a_smoke.txt
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; tail -c 600 gpurun_out/a_bench.err; head -c 1500 gpurun_out/a_bench.json
echo "== ref ext ops" ; timeout 600 python tools/ref_ext_bench.py --ops --batch 128 --out gpurun_out/a_ref_ext_ops.json > /dev/null 2> gpurun_out/a_ref_ext_ops.err; tail -c 400 gpurun_out/a_ref_ext_ops.err
echo "== old suite" ; timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_tc_fullsize_gpu.py --deselect tests/test_slak_tiny_step_gpu.py --deselect tests/test_syncbn_2rank_gpu.py --deselect tests/test_ref_ext_gpu.py 2>&1 | tail -8 | tee gpurun_out/a_old.txt
