#!/bin/bash
# quick GPU check: parity subset + bench at the default and alternative pack parameters
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -40 gpurun_out/pytest_gpu.log | cut -c1-220
for cfg in "${@:-128 512 896}"; do
  set -- $cfg
  echo "== threads=$1 owned=$2 local=$3 fill=${4:-default} repair=${5:-0}"
  python bench.py --steps 20 --warmup 3 --no-cpu --threads $1 --max-owned $2 --max-local $3 ${4:+--fill $4} ${5:+--repair $5} 2>>gpurun_out/bench.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); p = d['pack']
    print({k: d[k] for k in ('value','ms_per_step','kernels_ms')}, 'patches', p['n_patches'], 'listed', p['n_listed'], 'rounds', p['max_rounds'], 'fill', round(p['n_listed']/max(1,p['n_slots']),3), 'hw_excess', round(p['n_hw_excess']/max(1,p['n_hw_groups']),3), 'pack_s', round(p['seconds'],1), 'e2e_ms', d['e2e']['ms_per_step'])
"
done 2>&1 | tee gpurun_out/quick.log
tail -3 gpurun_out/bench.err
