#!/bin/bash
# pass 3h: synccheck over everything that launches attention, incl. the randomized tests (speculation with partial-warp
# prefetch loops), on the experiments build with the cluster size capped at 8 through the environment
set -u
mkdir -p gpurun_out
export VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so VB200_MAX_CLUSTER=8
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 99 --launch-timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider \
   -k "not split_kv_factor and not small_batch_long_context and not experiments_build and not library and not speculative and not host_threads" > gpurun_out/r3h_synccheck.log 2>&1
echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r3h_synccheck.log | tail -3; grep -c "Divergent\|Missing wait" gpurun_out/r3h_synccheck.log
