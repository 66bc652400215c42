#!/bin/bash
# A/B of the overlapped halo exchange: bash tools/run_multi_ab.sh N facets-per-gpu
set -u
N=${1:-2}; F=${2:-10000000}
mkdir -p gpurun_out
for cfg in "1 4" "1 0" "0 0" "1 8"; do
  set -- $cfg
  MS_OVERLAP=$1 MS_RESERVE_SMS=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 30 --warmup 5 --facets $F 2>gpurun_out/ab.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('overlap=$1 reserve=$2', round(d['value'],3), 'Gf/s', round(d['ms_per_step'],4), 'ms', 'e2e', round(d['e2e']['ms_per_step'],3))
"
done
tail -3 gpurun_out/ab.err
