#!/bin/bash
# round-2 pass I: where does a lone warp spend a stage?  ncu source-level captures (stall sampling per SASS line) of
#  (a) 1 sequence x 8192 tokens, no split: 12 CTAs, one per SM, 4 warps each -- per-warp latency bound
#  (b) configs[1] 64 x 512: the mid-size shape
set -u
mkdir -p gpurun_out
export VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pa_v1_kernel -s 8 -c 1 -f -o gpurun_out/r2i_s1_8k_nosplit \
   python tools/sweep.py --configs S1_8k --splits 1 --out gpurun_out/r2i_tmp.json > gpurun_out/r2i_ncu_a.log 2>&1
echo "ncu a rc=$?"; tail -2 gpurun_out/r2i_ncu_a.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pa_v1_kernel -s 8 -c 1 -f -o gpurun_out/r2i_c1 \
   python tools/sweep.py --configs C1 --out gpurun_out/r2i_tmp.json > gpurun_out/r2i_ncu_b.log 2>&1
echo "ncu b rc=$?"; tail -2 gpurun_out/r2i_ncu_b.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
