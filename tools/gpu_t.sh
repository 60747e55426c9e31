#!/bin/bash
# full GPU suite, smoke, then the bench line (with e2e pipelining) and the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -8 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; tail -c 300 gpurun_out/t_bench.err
python tools/show_bench.py gpurun_out/t_bench.json > gpurun_out/t_show.txt; head -1 gpurun_out/t_show.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/t_ref.json 2> gpurun_out/t_ref.err; tail -c 200 gpurun_out/t_ref.err; cut -c1-400 gpurun_out/t_ref.json
exit 0
