#!/bin/bash
# compute-sanitizer, final pass of round 2 (run on the final code): cluster split-KV merge, paged_attention_v2,
# reshape_and_cache (warp / 16-token x 8-warp / 32-token tiles), fused append, grouped-query, size extremes, two
# streams, prefill through the paged path, the reference host on the drop-in.
#
# synccheck and clusters of more than 8 CTAs: the tool reports "Barrier error detected. Missing wait" for mbarriers
# of cluster ranks >= 8 on ANY kernel that reuses an mbarrier over several phases -- including the stand-alone
# textbook kernel tools/synccheck_cluster_repro.cu (mode 1), which is clean at cluster size 8 and flagged at 9..16
# while computing the right result without the tool (profiles/r02f_sanitizer.md).  So synccheck runs with the cluster
# size capped at 8 (experiments build + VB200_MAX_CLUSTER; same kernel code, the cap only changes the launch geometry)
# over the whole parity file, and over the shapes of the tests that assert a larger launch plan; memcheck and
# racecheck run at the shipped cap of 16.
set -u
mkdir -p gpurun_out
SEL='split_kv or v2 or two_streams or reshape_and_cache or size_extremes or fused_append or grouped or random_shapes or speculative or host_threads'
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --launch-timeout 300 \
      python -m pytest tests/test_parity_gpu.py -q -m gpu -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer4_$tool.log 2>&1
  echo "$tool parity rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer4_$tool.log | tail -3
done
echo "== compute-sanitizer --tool synccheck: the whole parity file, cluster size capped at 8 (experiments build + VB200_MAX_CLUSTER)"
# (a) minus the tests that assert a > 8-CTA launch plan or that need the shipped library / a subprocess of their own
VB200_MAX_CLUSTER=8 VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so \
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 99 --launch-timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider \
    -k "not split_kv_factor and not small_batch_long_context and not experiments_build and not library and not speculative and not host_threads" > gpurun_out/sanitizer4_synccheck.log 2>&1
echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer4_synccheck.log | tail -3
echo "== synccheck (b) the shapes of the deselected split tests, cluster size capped at 8"
for lens in "4096 3000" "8192 5000" "7000 8192" "32768" "128 1024 2048"; do
  VB200_MAX_CLUSTER=8 VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so \
  timeout 600 compute-sanitizer --tool synccheck --error-exitcode 99 --launch-timeout 300 \
      python tools/synccheck_split_cases.py 0 $lens > gpurun_out/sanitizer4_synccheck_cap8.tmp 2>&1
  echo "lens=$lens rc=$? $(grep -E 'plan:|result ok|ERROR SUMMARY' gpurun_out/sanitizer4_synccheck_cap8.tmp | tr '\n' ' ')"
  cat gpurun_out/sanitizer4_synccheck_cap8.tmp >> gpurun_out/sanitizer4_synccheck_cap8.log
done
echo "== synccheck (c) the tool on a stand-alone textbook kernel, cluster sizes 8 / 9 / 16"
for n in 8 9 16; do
  timeout 120 compute-sanitizer --tool synccheck --error-exitcode 99 tools/bin/synccheck_cluster_repro 1 $n > gpurun_out/sanitizer4_repro_$n.log 2>&1
  echo "n=$n rc=$? $(grep -c 'Missing wait' gpurun_out/sanitizer4_repro_$n.log) x 'Missing wait', ranks: $(grep -o 'in block ([0-9]*,[0-9]*,[0-9]*)' gpurun_out/sanitizer4_repro_$n.log | sed 's/.*,\([0-9]*\))/\1/' | sort -un | tr '\n' ' ')"
done
tools/bin/synccheck_cluster_repro 1 8 9 16
echo "== memcheck: decode driver (prefill through the paged path) and the reference host"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --launch-timeout 300 \
    python -m pytest tests/test_decode_gpu.py -q -m gpu -k "paged_prefill or prefill" -p no:cacheprovider > gpurun_out/sanitizer4_decode_memcheck.log 2>&1
echo "memcheck decode rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer4_decode_memcheck.log | tail -2
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --launch-timeout 300 \
    python -m pytest tests/test_reference_host_gpu.py -q -m gpu -k "kernel_test or scheduler" -p no:cacheprovider > gpurun_out/sanitizer4_host_memcheck.log 2>&1
echo "memcheck reference host rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer4_host_memcheck.log | tail -2
