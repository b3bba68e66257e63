#!/bin/bash
# round-2 pass D: parity with the cluster merge, reshape_and_cache kernel modes incl. the rows kernel, CTA timelines
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2d_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2d_pytest_gpu.log
export VLLMINI_B200_NO_SHIM=1
echo "== reshape_and_cache: check + sweep"
VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so timeout 600 python tools/sweep_cache.py --check --modes 0,1,3,4,5 --out gpurun_out/r2d_sweep_cache.json > gpurun_out/r2d_sweep_cache.log 2>&1
echo "cache sweep rc=$?"; grep -c "ok$" gpurun_out/r2d_sweep_cache.log; grep "MISMATCH\|check:" gpurun_out/r2d_sweep_cache.log | head
echo "== CTA timelines: speculative table prefetch on / off"
VB200_TRACE_LIB=$PWD/vllmini_b200/lib/trace/libvllmini_b200.so timeout 300 python tools/trace_timeline.py M C1 S16_512 S1_1k > gpurun_out/r2d_trace_spec.jsonl 2> gpurun_out/r2d_trace_spec.err
echo "trace rc=$?"
VB200_TRACE_LIB=$PWD/vllmini_b200/lib/trace_nospec/libvllmini_b200.so timeout 300 python tools/trace_timeline.py M C1 S16_512 S1_1k > gpurun_out/r2d_trace_nospec.jsonl 2> gpurun_out/r2d_trace_nospec.err
echo "trace rc=$?"; tail -2 gpurun_out/r2d_trace_nospec.err
echo "== A/B sweep spec / nospec (same box)"
for v in trace trace_nospec; do
VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/$v/libvllmini_b200.so timeout 300 python tools/sweep.py --graph --configs C1,S16_512,S1_1k,S1_48,M,MIX --out gpurun_out/r2d_sweep_$v.json > gpurun_out/r2d_sweep_$v.log 2>&1
echo "sweep $v rc=$?"
done
