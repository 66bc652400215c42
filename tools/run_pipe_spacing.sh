#!/bin/bash
# e2e of bench.py: quadratic vs uniform chunk spacing of the pipelined host evaluation (+ the parity test)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k pipelined 2>&1 | tail -2
for mode in quadratic uniform quadratic uniform; do
  MS_PIPE_SPACING=$mode timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu 2>>gpurun_out/bench.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$mode', round(d['ms_per_step'], 4), 'e2e ms', round(d['e2e']['ms_per_step'], 4), round(d['e2e']['value'], 4))
"
done 2>&1 | tee gpurun_out/pipe_spacing.log
