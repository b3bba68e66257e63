#!/bin/bash
# round-2 pass M: push-based cluster merge (partials pushed into split 0's inbox, remote mbarrier arrive, no exit barrier)
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2m_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2m_pytest_gpu.log
echo "== sweep (defaults), graph timing"
timeout 600 python tools/sweep.py --graph --configs S1_48,S4_64,S1_1k,S1_8k,S4_4k,S8_2k,S16_512,S32_1k,C1,M,MIX,GQA4_d128,GQA8_d128,GQA4_d64,D128 --out gpurun_out/r2m_sweep_default.json > gpurun_out/r2m_sweep_default.log 2>&1
echo "sweep rc=$?"; python - <<'PY'
import json
for r in json.load(open('gpurun_out/r2m_sweep_default.json')):
    print(f"{r['config']:10s} {r['us_median']:8.2f} us  {r['gbps']:8.1f} GB/s")
PY
echo "== trace"
timeout 600 python tools/trace_timeline.py S1_8k S4_4k S1_1k > gpurun_out/r2m_trace.jsonl 2> gpurun_out/r2m_trace.err
python - <<'PY'
import json
for l in open('gpurun_out/r2m_trace.jsonl'):
    r=json.loads(l)
    print(r['config'], 'ctas',r['ctas'],'span',round(r['span_us'],2))
    for k,v in r['phases_us_p10_p50_p90_max'].items(): print('   ',k.ljust(36),v)
PY
echo "== memcheck / racecheck on the split tests"
for tool in memcheck racecheck; do
timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 --launch-timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "split_kv or two_streams or size_extremes or random_shapes" -p no:cacheprovider > gpurun_out/r2m_$tool.log 2>&1
echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2m_$tool.log | tail -2
done
