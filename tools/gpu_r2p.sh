#!/bin/bash
set -u
SEL='split_kv or v2 or two_streams or reshape_and_cache or size_extremes or fused_append or grouped'
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 99 --launch-timeout 300 \
    python -m pytest tests/test_parity_gpu.py -q -m gpu -k "($SEL) and not two_streams and not small_batch_long_context and not 32768" -p no:cacheprovider > gpurun_out/sanitizer4_synccheck.log 2>&1
echo "synccheck (a) rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer4_synccheck.log | tail -3
