#!/bin/bash
# round-2 pass C: cluster split-KV -- parity, then sweeps (experiments build): merge mode x forced splits, cache kernels
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2c_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2c_pytest_gpu.log
export VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so VLLMINI_B200_NO_SHIM=1
echo "== split sweep: cluster vs workspace"
timeout 900 python tools/sweep.py --graph --configs S1_1k,S1_8k,S4_4k,S8_2k,S16_512,S32_1k,C1,S128_256,S128_1k,M --splits 1,2,4,8,16 --merge 0,1 --out gpurun_out/r2c_sweep_split.json > gpurun_out/r2c_sweep_split.log 2>&1
echo "sweep rc=$?"; tail -3 gpurun_out/r2c_sweep_split.log
timeout 300 python tools/sweep.py --graph --configs S1_48,S4_64,S1_1k,S1_8k,S4_4k,S8_2k,S16_512,S32_1k,C1,M,MIX --ref --out gpurun_out/r2c_sweep_default.json > gpurun_out/r2c_sweep_default.log 2>&1
echo "default sweep rc=$?"; tail -3 gpurun_out/r2c_sweep_default.log
echo "== reshape_and_cache sweep"
timeout 600 python tools/sweep_cache.py --modes 0,1,2,3 --out gpurun_out/r2c_sweep_cache.json > gpurun_out/r2c_sweep_cache.log 2>&1
echo "cache sweep rc=$?"; tail -3 gpurun_out/r2c_sweep_cache.log
