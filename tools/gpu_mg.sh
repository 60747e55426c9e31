#!/bin/bash
# 2-GPU check: SyncBN peer-memory exchange parity, then the bench line with the peer exchange and with the NCCL fallback
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_syncbn_peer_2gpu.py -m gpu -q -x 2>&1 | tail -8 | cut -c1-300
for MODE in 0 1; do
  SLAK_SYNCBN_NCCL=$MODE timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$MODE \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/mg${N}_syncbn_nccl$MODE.json 2> gpurun_out/mg${N}_syncbn_nccl$MODE.err
  tail -c 300 gpurun_out/mg${N}_syncbn_nccl$MODE.err
  python tools/show_bench.py gpurun_out/mg${N}_syncbn_nccl$MODE.json | head -1
done
