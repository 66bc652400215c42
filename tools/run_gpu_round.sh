#!/bin/bash
# First GPU pass: parity tests, fp64/HBM ceilings, bench, pack-parameter sweep, ncu launch list.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
tools/_build/fp64_peak > gpurun_out/fp64_peak.json 2>&1; cat gpurun_out/fp64_peak.json
python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
for cfg in "128 512 896" "256 512 896" "256 1024 1600" "128 256 512" "256 768 1200" "192 384 700"; do
  set -- $cfg
  echo "== threads=$1 owned=$2 local=$3"
  python bench.py --steps 20 --warmup 3 --no-cpu --threads $1 --max-owned $2 --max-local $3 2>>gpurun_out/bench.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('value','ms_per_step','kernels_ms')}, d['pack']['n_patches'], d['pack']['n_listed'], d['pack']['max_rounds'], d['pack']['seconds'])
"
done 2>&1 | tee gpurun_out/sweep.log
