#!/bin/bash
# compute-sanitizer over what round 2 added: split-KV merged through a thread-block cluster (barrier.cluster +
# ld.shared::cluster), paged_attention_v2 (partition records + reduce kernel), the register path of scattered
# reshape_and_cache tiles, concurrent split launches on two streams, the reference host running on the drop-in.
set -u
mkdir -p gpurun_out
SEL='split_kv or v2 or two_streams or reshape_and_cache or size_extremes or fused_append or grouped'
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --launch-timeout 300 \
      python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer3_$tool.log 2>&1
  echo "$tool parity rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer3_$tool.log | tail -3
done
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --launch-timeout 300 \
    python -m pytest tests/test_reference_host_gpu.py -x -q -m gpu -k "kernel_test or scheduler" -p no:cacheprovider > gpurun_out/sanitizer3_host_memcheck.log 2>&1
echo "memcheck reference host rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer3_host_memcheck.log | tail -3
