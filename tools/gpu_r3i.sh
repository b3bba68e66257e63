#!/bin/bash
set -u
mkdir -p gpurun_out
export VLLMINI_B200_NO_SHIM=1
T="tests/test_parity_gpu.py::test_paged_attention_random_shapes[22]"
run() { # name lib extra-env
  env $3 VLLMINI_B200_LIB=$2 timeout 300 compute-sanitizer --tool synccheck --error-exitcode 99 --launch-timeout 300 python -m pytest "$T" -q -m gpu -p no:cacheprovider > gpurun_out/r3i_$1.log 2>&1
  echo "$1 rc=$? divergent=$(grep -c Divergent gpurun_out/r3i_$1.log) $(grep -E 'passed|failed' gpurun_out/r3i_$1.log | tail -1)"
}
run exp_spec_on  $PWD/vllmini_b200/lib/experiments/libvllmini_b200.so VB200_SPEC_PREFETCH=1
run exp_spec_off $PWD/vllmini_b200/lib/experiments/libvllmini_b200.so VB200_SPEC_PREFETCH=0
run default      $PWD/vllmini_b200/lib/libvllmini_b200.so A=1
run at_74cf157   $PWD/vllmini_b200/lib/at_74cf157/libvllmini_b200.so A=1
run at_455c9eb   $PWD/vllmini_b200/lib/at_455c9eb/libvllmini_b200.so A=1
