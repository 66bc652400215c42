#!/bin/bash
# A/B of the pipelined host evaluation on one box: the new parity test, then bench e2e with and without it
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or full_size or icosphere or trial" > gpurun_out/pytest_pipe.log 2>&1
tail -15 gpurun_out/pytest_pipe.log | cut -c1-220
for mode in pipe nopipe pipe nopipe; do
  if [ $mode = nopipe ]; then export MS_NO_PIPELINE=1; else unset MS_NO_PIPELINE; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu 2>>gpurun_out/bench.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$mode', {k: d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e'])
"
done 2>&1 | tee gpurun_out/pipe_ab.log
tail -3 gpurun_out/bench.err
