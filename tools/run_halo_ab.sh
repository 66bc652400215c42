#!/bin/bash
# N-GPU check of the peer-memory halo (bitwise vs NCCL) and the weak-scaling bench with either transport
set -u
N=${1:-2}
F=${2:-10000000}
mkdir -p gpurun_out
nvidia-smi -L | head -$N
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
  tools/multi_halo_check.py 2500000 > gpurun_out/halo_check_$N.log 2> gpurun_out/halo_check_$N.err
echo "check rc=$?"; cat gpurun_out/halo_check_$N.log; grep -v "^W\|^\[W\|warn" gpurun_out/halo_check_$N.err | tail -8
for T in peer nccl; do
  MS_HALO=$T timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --gpus $N --steps 20 --warmup 3 --facets $F > gpurun_out/multi_${N}_$T.log 2> gpurun_out/multi_${N}_$T.err
  echo "bench $T rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/multi_${N}_$T.log"):
    try: d = json.loads(l)
    except Exception: continue
    print("$T", d["value"], d["ms_per_step"], d["collectives_per_step"], "e2e", d["e2e"]["value"])
PY
  tail -3 gpurun_out/multi_${N}_$T.err
done
