#!/bin/bash
# round-2 pass F: new tests, reshape_and_cache variants (256 threads / eager loads), synccheck follow-up
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (decode + parity)"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2f_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2f_pytest_gpu.log
echo "== reshape_and_cache variants"
VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so timeout 600 python tools/sweep_cache.py --check --modes 0,1,2,3,6,7,8,9 --out gpurun_out/r2f_sweep_cache.json > gpurun_out/r2f_sweep_cache.log 2>&1
echo "cache sweep rc=$?"; grep "MISMATCH\|check:" gpurun_out/r2f_sweep_cache.log | head -5
echo "== synccheck: rest of the selection, then the two-stream test alone (concurrent and serialised)"
SEL='(split_kv or v2 or reshape_and_cache or size_extremes or fused_append or grouped) and not two_streams'
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 99 --launch-timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer3_synccheck_rest.log 2>&1
echo "synccheck rest rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer3_synccheck_rest.log | tail -2
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 99 --launch-timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "two_streams" -p no:cacheprovider > gpurun_out/sanitizer3_synccheck_two_streams.log 2>&1
echo "synccheck two_streams concurrent rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer3_synccheck_two_streams.log | tail -2
CUDA_LAUNCH_BLOCKING=1 timeout 600 compute-sanitizer --tool synccheck --error-exitcode 99 --launch-timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "two_streams" -p no:cacheprovider > gpurun_out/sanitizer3_synccheck_two_streams_blocking.log 2>&1
echo "synccheck two_streams serialised rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer3_synccheck_two_streams_blocking.log | tail -2
