#!/bin/bash
set -u
mkdir -p gpurun_out
export VLLMINI_B200_NO_SHIM=1
run() { # name lib test
  VLLMINI_B200_LIB=$2 timeout 300 compute-sanitizer --tool synccheck --error-exitcode 99 --launch-timeout 300 python -m pytest "$3" -q -m gpu -p no:cacheprovider > gpurun_out/r3j_$1.log 2>&1
  echo "$1 rc=$? divergent=$(grep -c Divergent gpurun_out/r3j_$1.log) missing=$(grep -c 'Missing wait' gpurun_out/r3j_$1.log) $(grep -E 'passed|failed' gpurun_out/r3j_$1.log | tail -1)"
}
T22="tests/test_parity_gpu.py::test_paged_attention_random_shapes[22]"
run seed22_at_68c8d3d $PWD/vllmini_b200/lib/at_68c8d3d/libvllmini_b200.so "$T22"
run seed22_current    $PWD/vllmini_b200/lib/libvllmini_b200.so "$T22"
run split2048_current $PWD/vllmini_b200/lib/libvllmini_b200.so "tests/test_parity_gpu.py::test_split_kv_factor_does_not_change_the_result[2048]"
run split1280_current $PWD/vllmini_b200/lib/libvllmini_b200.so "tests/test_parity_gpu.py::test_split_kv_factor_does_not_change_the_result[1280]"
# the same case as seed 22 but through the stand-alone script with different hints
for msl in 1024 2048 3072; do
VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/libvllmini_b200.so VB200_MSL=$msl timeout 300 compute-sanitizer --tool synccheck --error-exitcode 99 --launch-timeout 300 python tools/synccheck_seed22.py > gpurun_out/r3j_msl$msl.log 2>&1
echo "msl=$msl rc=$? divergent=$(grep -c Divergent gpurun_out/r3j_msl$msl.log) $(grep -E 'plan|ok' gpurun_out/r3j_msl$msl.log | tr '\n' ' ')"
done
