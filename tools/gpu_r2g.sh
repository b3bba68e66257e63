#!/bin/bash
# round-2 pass G: is synccheck's "Missing wait" on cluster rank 8 a property of the tool or of the kernel?
#  1. a stand-alone textbook mbarrier + bulk-copy kernel in clusters of 2..16 CTAs under synccheck
#  2. the product's split-KV test per split factor under synccheck (1 / 2 / 5 / 8 splits vs 9)
#  3. metric-shape attention with the block ids spread over a pool 1x / 12x the batch (TLB reach)
set -u
mkdir -p gpurun_out
echo "== 1. synccheck on the stand-alone cluster kernel"
for n in 2 4 8 9 12 16; do
  timeout 120 compute-sanitizer --tool synccheck --error-exitcode 99 tools/bin/synccheck_cluster_repro $n > gpurun_out/r2g_repro_$n.log 2>&1
  rc=$?
  echo "n=$n rc=$rc $(grep -c 'Missing wait' gpurun_out/r2g_repro_$n.log) x 'Missing wait'; blocks: $(grep -o 'in block ([0-9,]*)' gpurun_out/r2g_repro_$n.log | sort -u | tr '\n' ' ')"
  grep -E "cluster \(|ERROR SUMMARY" gpurun_out/r2g_repro_$n.log
done
echo "-- same binary without the sanitizer"
tools/bin/synccheck_cluster_repro 2 4 8 9 12 16
echo "== 2. synccheck on the product per split factor"
for m in 64 512 1280 2048 4608; do
  timeout 300 compute-sanitizer --tool synccheck --error-exitcode 99 --launch-timeout 300 python -m pytest \
     "tests/test_parity_gpu.py::test_split_kv_factor_does_not_change_the_result[$m]" -q -m gpu -p no:cacheprovider > gpurun_out/r2g_sync_split_$m.log 2>&1
  rc=$?
  echo "max_seq_len=$m rc=$rc $(grep -c 'Missing wait' gpurun_out/r2g_sync_split_$m.log) x 'Missing wait'; blocks z: $(grep -o 'in block ([0-9]*,[0-9]*,[0-9]*)' gpurun_out/r2g_sync_split_$m.log | sed 's/.*,\([0-9]*\))/\1/' | sort -un | tr '\n' ' ')"
  grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2g_sync_split_$m.log | tail -2
done
echo "== 3. pool spread"
timeout 300 python tools/sweep.py --graph --configs M,C1 --pool-mult 1 --out gpurun_out/r2g_pool1.json > gpurun_out/r2g_pool1.log 2>&1; tail -3 gpurun_out/r2g_pool1.log
timeout 300 python tools/sweep.py --graph --configs M,C1 --pool-mult 12 --out gpurun_out/r2g_pool12.json > gpurun_out/r2g_pool12.log 2>&1; tail -3 gpurun_out/r2g_pool12.log
echo "== 4. reshape_and_cache: shipped dispatch (512 < T <= 2048 now 16-token tiles on 8 warps) vs reference, bit-exactness"
VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so timeout 600 python tools/sweep_cache.py --check --modes 0,1,2,3 --out gpurun_out/r2g_sweep_cache.json > gpurun_out/r2g_sweep_cache.log 2>&1
echo "cache sweep rc=$?"; grep "MISMATCH\|check:" gpurun_out/r2g_sweep_cache.log | head -5
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "reshape_and_cache" -p no:cacheprovider 2>&1 | tail -2
