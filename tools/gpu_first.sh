#!/bin/bash
# First GPU pass: quick hang check, parity suite, reference-kernel goldens, variant sweep.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
echo "== quick hang check" 
timeout 120 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "ref_test_shape or bit_exact" > gpurun_out/quick.log 2>&1
echo "quick rc=$?"; tail -5 gpurun_out/quick.log
echo "== full parity"
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
echo "== goldens from the reference kernels"
timeout 300 python tests/golden/make_golden_refkernel.py --out gpurun_out/golden > gpurun_out/golden.log 2>&1
echo "golden rc=$?"; tail -3 gpurun_out/golden.log
echo "== sweep"
timeout 600 python tools/sweep.py --configs M,C1 --variants 0,1,2 --splits 0 --ref > gpurun_out/sweep.log 2>&1
echo "sweep rc=$?"; cat gpurun_out/sweep.log
