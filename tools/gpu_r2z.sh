#!/bin/bash
set -u
mkdir -p gpurun_out
export VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so
timeout 900 python tools/sweep.py --graph --configs S16_512,S32_1k,C1,S128_256,S8_2k,S4_4k,S1_8k --splits 0,2,3,4,6,8 --out gpurun_out/r2z_sweep_splits.json > gpurun_out/r2z_sweep_splits.log 2>&1
echo rc=$?
python - <<'PY'
import json,collections
d=collections.defaultdict(dict)
for r in json.load(open('gpurun_out/r2z_sweep_splits.json')): d[r['config']][r['splits']]=r['us_median']
print('config     '+''.join(f"{s:>8d}" for s in (0,2,3,4,6,8)))
for k,v in d.items(): print(f"{k:10s} "+''.join(f"{v.get(s,float('nan')):8.2f}" for s in (0,2,3,4,6,8)))
PY
