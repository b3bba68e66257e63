#!/bin/bash
# pass 3l: split factor capped by the resident CTA slots (head sizes >= 96): new default vs the old factor, forced
set -u
mkdir -p gpurun_out
export VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so
timeout 900 python tools/sweep.py --graph --configs D128_s16_1k,D96_s40_1k,D256_s10_2k,D112_s2_3k,D192_s6_2k --splits 0,1,2,3,4,5,6,7,8 --out gpurun_out/r3l_sweep.json > gpurun_out/r3l_sweep.log 2>&1
echo rc=$?
python - <<'PY'
import json,collections
d=collections.defaultdict(dict)
for r in json.load(open('gpurun_out/r3l_sweep.json')): d[r['config']][r['splits']]=r['us_median']
print('config        '+''.join(f"{s:>8d}" for s in range(9)))
for k,v in d.items(): print(f"{k:13s} "+''.join(f"{v.get(s,float('nan')):8.2f}" for s in range(9)))
PY
unset VLLMINI_B200_LIB
echo "== attention parity on the default build"
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
