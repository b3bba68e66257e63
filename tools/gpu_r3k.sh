#!/bin/bash
set -u
mkdir -p gpurun_out
for n in 1 2 3 4 8; do
  timeout 300 compute-sanitizer --tool synccheck --error-exitcode 99 tools/bin/synccheck_cluster_repro 2 $n > gpurun_out/r3k_repro_multiwave_$n.log 2>&1
  echo "multi-wave n=$n rc=$? divergent=$(grep -c Divergent gpurun_out/r3k_repro_multiwave_$n.log) missing=$(grep -c 'Missing wait' gpurun_out/r3k_repro_multiwave_$n.log)"
  grep -E "cluster \(|ERROR SUMMARY" gpurun_out/r3k_repro_multiwave_$n.log | head -2
done
tools/bin/synccheck_cluster_repro 2 1 2 3 4 8
