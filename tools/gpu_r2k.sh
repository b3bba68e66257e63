#!/bin/bash
# round-2 pass K: per-CTA phase timeline of the split launches (where do 1x8192 / 4x4096 / 16x512 spend their time?)
set -u
mkdir -p gpurun_out
timeout 600 python tools/trace_timeline.py S1_8k S4_4k S8_2k S16_512 S1_1k C1 > gpurun_out/r2k_trace.jsonl 2> gpurun_out/r2k_trace.err
echo "trace rc=$?"; tail -3 gpurun_out/r2k_trace.err
python - <<'PY'
import json
for l in open('gpurun_out/r2k_trace.jsonl'):
    r=json.loads(l)
    print(r['config'], 'ctas',r['ctas'],'span',round(r['span_us'],2),'sms',r['sms_used'],'maxconc',r['max_concurrent_ctas'])
    for k,v in r['phases_us_p10_p50_p90_max'].items(): print('   ',k.ljust(36),v)
PY
