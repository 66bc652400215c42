#!/bin/bash
# e2e of bench.py for several chunk counts of the pipelined host evaluation
set -u
mkdir -p gpurun_out
for n in 8 16 24 8 16 24; do
  MS_PIPE_CHUNKS=$n timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu 2>>gpurun_out/bench.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('chunks $n', round(d['ms_per_step'], 4), 'e2e ms', round(d['e2e']['ms_per_step'], 4), round(d['e2e']['value'], 4))
"
done 2>&1 | tee gpurun_out/pipe_chunks.log
