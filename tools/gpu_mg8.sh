#!/bin/bash
# N-GPU check of the bench line (peer-memory SyncBN + bucketed all-reduce in the captured step)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 10 --warmup 3 --no-ref-ext --no-cpu-baseline > gpurun_out/mg${N}_bench.json 2> gpurun_out/mg${N}_bench.err
tail -c 400 gpurun_out/mg${N}_bench.err
grep "^{" gpurun_out/mg${N}_bench.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
exit 0
