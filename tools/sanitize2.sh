#!/bin/bash
# compute-sanitizer passes over what the second session of round 1 added: every head size (partial chunks, q in
# shared memory, 3-warp CTAs), fused append at head 80 / 256, add_layernorm, argmax_rows, the staging ring.
set -u
mkdir -p gpurun_out
SEL='head80 or head96 or head112 or head192 or head256 or fused_append'
SEL2='add_layernorm or argmax_rows or async_step or fused_layernorm'
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 99 --launch-timeout 300 \
      python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer2_$tool.log 2>&1
  echo "$tool parity rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer2_$tool.log | tail -3
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 99 --launch-timeout 300 \
      python -m pytest tests/test_decode_gpu.py -x -q -m gpu -k "$SEL2" -p no:cacheprovider > gpurun_out/sanitizer2d_$tool.log 2>&1
  echo "$tool decode rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer2d_$tool.log | tail -3
done
