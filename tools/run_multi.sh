#!/bin/bash
# N-GPU bench (weak scaling): bash tools/run_multi.sh N [facets-per-gpu]
set -u
N=${1:-2}
F=${2:-10000000}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 3 --facets $F > gpurun_out/multi_$N.log 2> gpurun_out/multi_$N.err
echo "rc=$?"
cat gpurun_out/multi_$N.log
tail -20 gpurun_out/multi_$N.err
