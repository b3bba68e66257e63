#!/bin/bash
# pass 3m: occupancy-checked split factor -- defaults at large head sizes, full suite, broad synccheck (cluster cap 8)
set -u
mkdir -p gpurun_out
VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so timeout 600 python tools/sweep.py --graph --configs D128_s16_1k,D96_s40_1k,D256_s10_2k,D112_s2_3k,D192_s6_2k,S1_1k,S1_8k,S4_4k,S8_2k,S16_512 --out gpurun_out/r3m_sweep.json > gpurun_out/r3m_sweep.log 2>&1
python - <<'PY'
import json
for r in json.load(open('gpurun_out/r3m_sweep.json')): print(f"{r['config']:14s} {r['us_median']:9.2f} us")
PY
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r3m_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r3m_pytest_gpu.log
export VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so VB200_MAX_CLUSTER=8
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 99 --launch-timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider \
   -k "not split_kv_factor and not small_batch_long_context and not experiments_build and not library and not speculative and not host_threads" > gpurun_out/r3m_synccheck.log 2>&1
echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r3m_synccheck.log | tail -3; echo "divergent/missing: $(grep -c 'Divergent\|Missing wait' gpurun_out/r3m_synccheck.log)"; grep -E "^FAILED" gpurun_out/r3m_synccheck.log | head -3
