#!/bin/bash
# ncu evidence: launch list + one full capture of the patch kernels (1 GPU).
# BENCH_ARGS: extra bench.py arguments; MS_B200_LIB: library variant.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu ${BENCH_ARGS:-}"
if [ "${NCU_LAUNCHES:-1}" = "1" ]; then
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
fi
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_patch -s 6 -c 2 -o gpurun_out/prof -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/ncu_full.log
