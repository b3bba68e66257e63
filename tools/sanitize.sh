#!/bin/bash
# compute-sanitizer passes over a representative subset of the GPU parity tests (memcheck, racecheck, synccheck).
set -u
mkdir -p gpurun_out
SEL='ragged_small or bs8_ragged or bs32_ragged or head128_bs16 or with_empty_seq or split_kv or fused_append or bit_exact or flash or copy_blocks'
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --launch-timeout 300 \
      python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -4
done
