#!/bin/bash
set -u
mkdir -p gpurun_out
for c in S16_512 C1 S128_256; do timeout 300 python tools/trace_pdl_gap.py $c; done > gpurun_out/r2y_pdl_gap.jsonl 2> gpurun_out/r2y_err.log
echo rc=$?; tail -3 gpurun_out/r2y_err.log
python - <<'PY'
import json
for l in open('gpurun_out/r2y_pdl_gap.jsonl'):
    r=json.loads(l); print(r['config'],'pdl',r['pdl'])
    for k,v in r.items():
        if k not in ('config','pdl'): print('   ',k,v)
PY
