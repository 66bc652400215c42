#!/bin/bash
# Packer parameter sweep on the headline workload: fill percentage, round width, patch size.
# usage: bash tools/pack_sweep.sh "<bench args>" ...   (one bench run per argument string)
for args in "$@"; do
  python bench.py --steps 10 --warmup 3 --no-cpu --no-parity --no-full-mesh --strong-facets 0 $args 2>/dev/null |
    python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
k, p = d['kernels_ms'], d['pack']
print('%-40s step %.4f ms  A %.4f B %.4f  patches %d listed %.3f fill %.3f conflicts/hw-group %.3f' % ('$args', d['ms_per_step'], k['pass_a'], k['pass_b'], p['n_patches'], p['n_listed'] / p['nf'], p['n_listed'] / p['n_slots'], p['n_hw_excess'] / max(1, p['n_hw_groups'])))"
done
