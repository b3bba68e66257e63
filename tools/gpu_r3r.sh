#!/bin/bash
# final sweep: every shape of tools/sweep.py, new kernel vs the reference's own kernel (graph-timed)
set -u
mkdir -p gpurun_out
CFG=$(python - <<'PY'
import sys; sys.path.insert(0,'tools')
import importlib.util
spec=importlib.util.spec_from_file_location('sweep','tools/sweep.py'); 
src=open('tools/sweep.py').read()
import re
names=re.findall(r'^    "(\w+)": \(', src, re.M)
print(','.join(names))
PY
)
echo "$CFG"
timeout 1500 python tools/sweep.py --graph --ref --configs "$CFG" --out gpurun_out/r3r_sweep_all.json > gpurun_out/r3r_sweep_all.log 2>&1
echo rc=$?
python - <<'PY'
import json,collections
d=collections.defaultdict(dict)
for r in json.load(open('gpurun_out/r3r_sweep_all.json')): d[r['config']][r['impl']]=(r['us_median'],r['gbps'])
print(f"{'config':18s}{'new us':>10s}{'ref us':>10s}{'x':>7s}{'new GB/s':>10s}")
for k,v in d.items():
    n=v.get('new'); r=v.get('ref')
    print(f"{k:18s}{n[0]:10.2f}{(r[0] if r else float('nan')):10.2f}{(r[0]/n[0] if r else float('nan')):7.2f}{n[1]:10.0f}")
PY
