#!/bin/bash
# 8-GPU confirmation of the peer-memory transport: correctness check (20 M facets) + weak-scaling bench (10 M / GPU)
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
  tools/multi_halo_check.py 2500000 > gpurun_out/halo_check_$N.log 2> gpurun_out/halo_check_$N.err
echo "check rc=$?"; cat gpurun_out/halo_check_$N.log; grep -i "error\|assert\|Traceback" gpurun_out/halo_check_$N.err | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
  bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/multi_${N}_peer.log 2> gpurun_out/multi_${N}_peer.err
echo "bench rc=$?"; cat gpurun_out/multi_${N}_peer.log; grep -i "error\|assert\|Traceback" gpurun_out/multi_${N}_peer.err | head -5
