#!/bin/bash
# quick GPU check: smoke, parity tests, bench at the default and alternative pack parameters
# usage: tools/run_quick.sh ["threads owned local [lib]" ...]
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log | cut -c1-300
if [ "${SKIP_TESTS:-0}" = "0" ]; then timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log | cut -c1-220; fi
for cfg in "${@:-192 448 768}"; do
  set -- $cfg
  echo "== threads=$1 owned=$2 local=$3 lib=${4:-default} teams=${5:-auto}"
  MS_B200_LIB=${4:+$PWD/membrane_solver_b200/$4} MS_TEAMS=${5:-0} timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --threads $1 --max-owned $2 --max-local $3 2>>gpurun_out/bench.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); p = d['pack']
    print({k: d[k] for k in ('value','ms_per_step','kernels_ms')}, 'patches', p['n_patches'], 'listed', p['n_listed'], 'teams', p['teams'], 'steps', p['max_steps'], 'events', p['n_events'], 'maxE', p['max_events'], 'maxW', p['max_words'], 'maxL', p['max_local'], 'gather_excess', round(p['n_gather_excess']/max(1,p['n_gather_groups']),3), 'pack_s', round(p['seconds'],1), 'e2e_ms', d['e2e']['ms_per_step'], d['energies'])
"
done 2>&1 | tee gpurun_out/quick.log
tail -3 gpurun_out/bench.err
