#!/bin/bash
# round-2 pass H: synccheck follow-up.  Pass G: a one-phase stand-alone kernel is clean for every cluster size and the
# product is clean for 1/2/5/8/9 splits; the reports came from the 11- and 13-split launches.  Here: the stand-alone
# kernel with the product's multi-phase ring pattern, and the product with forced split factors 8..16.
set -u
mkdir -p gpurun_out
echo "== 1. stand-alone kernel, product pattern (mode 1)"
for n in 8 9 10 11 13 16; do
  timeout 120 compute-sanitizer --tool synccheck --error-exitcode 99 tools/bin/synccheck_cluster_repro 1 $n > gpurun_out/r2h_repro_$n.log 2>&1
  rc=$?
  echo "n=$n rc=$rc $(grep -c 'Missing wait' gpurun_out/r2h_repro_$n.log) x 'Missing wait'; blocks z: $(grep -o 'in block ([0-9]*,[0-9]*,[0-9]*)' gpurun_out/r2h_repro_$n.log | sed 's/.*,\([0-9]*\))/\1/' | sort -un | tr '\n' ' ')"
  grep -E "cluster \(|ERROR SUMMARY" gpurun_out/r2h_repro_$n.log
done
tools/bin/synccheck_cluster_repro 1 8 9 10 11 13 16
echo "== 2. product, forced split factor"
export VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so
for n in 8 9 10 11 12 13 16; do
  timeout 300 compute-sanitizer --tool synccheck --error-exitcode 99 --launch-timeout 300 python tools/synccheck_split_cases.py $n 4096 3000 > gpurun_out/r2h_sync_forced_$n.log 2>&1
  rc=$?
  echo "splits=$n rc=$rc $(grep -c 'Missing wait' gpurun_out/r2h_sync_forced_$n.log) x 'Missing wait'; blocks (y,z): $(grep -o 'in block ([0-9]*,[0-9]*,[0-9]*)' gpurun_out/r2h_sync_forced_$n.log | sed 's/in block ([0-9]*,\(.*\))/\1/' | sort -u | tr '\n' ' ')"
  grep -E "plan:|result ok|ERROR SUMMARY|Error" gpurun_out/r2h_sync_forced_$n.log | head -4
done
echo "== 3. product, 11 splits, uniform lengths / one sequence"
for lens in "4096 4096" "4096" "3000"; do
  timeout 300 compute-sanitizer --tool synccheck --error-exitcode 99 --launch-timeout 300 python tools/synccheck_split_cases.py 11 $lens > gpurun_out/r2h_tmp.log 2>&1
  echo "lens=$lens rc=$? $(grep -c 'Missing wait' gpurun_out/r2h_tmp.log) x 'Missing wait'; blocks (y,z): $(grep -o 'in block ([0-9]*,[0-9]*,[0-9]*)' gpurun_out/r2h_tmp.log | sed 's/in block ([0-9]*,\(.*\))/\1/' | sort -u | tr '\n' ' ')"
  grep -E "plan:|result ok|ERROR SUMMARY" gpurun_out/r2h_tmp.log | head -3
done
