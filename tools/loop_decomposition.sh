#!/bin/bash
# Where does the time of the patch kernels go?  Timing-only runs of the -DMS_DEBUG_VARIANTS build with parts of
# the consumer loop switched off (results are wrong by construction): bit 1 = no token ring / round barrier,
# bit 2 = no accumulation (shared-memory read-modify-writes), bit 4 = no per-facet compute, bit 8 = the producer
# stages header / records / ids only (no position or seed rows), bit 16 = the epilogue only clears the accumulators.
#   make -C membrane_solver_b200/csrc dbg && bash tools/loop_decomposition.sh > gpurun_out/loop_decomposition.txt
export MS_B200_LIB=$PWD/membrane_solver_b200/libms_b200_dbg.so
for v in ${VARIANTS:-0 1 2 3 4 5 6 7 8 16 24 31}; do
  MS_DEBUG_VARIANT=$v timeout 90 python bench.py --steps 10 --warmup 3 --no-cpu --no-parity --no-full-mesh --strong-facets 0 2>/dev/null |
    python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
k = d['kernels_ms']
print('variant $v  token=%d accumulate=%d compute=%d copies=%d epilogue=%d   pass A %.4f ms  pass B %.4f ms  step %.4f ms' % (not $v & 1, not $v & 2, not $v & 4, not $v & 8, not $v & 16, k['pass_a'], k['pass_b'], d['ms_per_step']))"
done
