#!/bin/bash
set -u
mkdir -p gpurun_out
export VLLMINI_B200_NO_SHIM=1
for v in at_455c9eb at_74cf157 current; do
  if [ $v = current ]; then LIB=$PWD/vllmini_b200/lib/libvllmini_b200.so; else LIB=$PWD/vllmini_b200/lib/$v/libvllmini_b200.so; fi
  VLLMINI_B200_LIB=$LIB timeout 600 compute-sanitizer --tool synccheck --error-exitcode 99 --launch-timeout 300 python -m pytest "tests/test_parity_gpu.py::test_paged_attention_vs_oracle_and_reference_kernel" -q -m gpu -p no:cacheprovider -k "head112 or head80 or head96 or head256 or head192 or ragged_small" > gpurun_out/r3g_sync_$v.log 2>&1
  echo "$v rc=$? $(grep -c 'Divergent' gpurun_out/r3g_sync_$v.log) divergent; $(grep -o 'pa_v1_kernel<[^>]*>' gpurun_out/r3g_sync_$v.log | sort | uniq -c | tr '\n' ' ')"
  grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r3g_sync_$v.log | tail -2
done
