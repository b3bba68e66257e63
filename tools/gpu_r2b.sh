#!/bin/bash
# round-2 pass B: parity suite, bench (both arms), reshape_and_cache sweep over kernel modes (experiments build)
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2b_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2b_pytest_gpu.log
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err
echo "bench ref rc=$?"; cut -c1-400 gpurun_out/r2b_bench_ref.json
echo "== reshape_and_cache sweep"
VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so VLLMINI_B200_NO_SHIM=1 timeout 600 python tools/sweep_cache.py --modes 0,1,2,3 --out gpurun_out/r2b_sweep_cache.json > gpurun_out/r2b_sweep_cache.log 2>&1
echo "sweep rc=$?"; tail -3 gpurun_out/r2b_sweep_cache.log
