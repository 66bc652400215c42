#!/bin/bash
# Full GPU pass of a round: parity tests, smoke, bench (both arms), ncu launch list + full capture.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>> gpurun_out/bench.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.log
bash tools/run_ncu.sh
# timing decomposition of the patch kernels (needs make -C membrane_solver_b200/csrc dbg)
[ -f membrane_solver_b200/libms_b200_dbg.so ] && VARIANTS="0 16 24 26 28 30 22" bash tools/loop_decomposition.sh > gpurun_out/loop_decomposition.txt 2>&1
# multi-GPU (run under gpurun --gpus N):
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N
#   ... tools/multi_halo_check.py        (NCCL vs peer vs fused vs in-kernel transport)
#   ... tools/multi_minimizer_check.py   (minimiser over a partitioned mesh vs one GPU)
