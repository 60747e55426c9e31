#!/bin/bash
# bucket size of the gradient all-reduce at N GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
for MB in ${2:-8 25 64 200}; do
  SLAK_BUCKET_MB=$MB timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N \
    bench.py --gpus $N --steps 10 --warmup 3 --no-ref-ext --no-cpu-baseline > gpurun_out/bucket_${N}_$MB.json 2> gpurun_out/bucket_${N}_$MB.err
  grep "^{" gpurun_out/bucket_${N}_$MB.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bucket_mb', $MB, 'N', d['n_gpus'], 'ms', round(d['ms_per_step'],3), 'img/s', round(d['value'],1))"
done
exit 0
