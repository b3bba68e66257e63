#!/bin/bash
# pass 3c: stage copies issued by one lane in straight-line code (variant "elect") vs one copy per lane
set -u
mkdir -p gpurun_out
CFG=S1_48,S1_1k,S1_8k,S4_4k,S8_2k,S16_512,S32_1k,C1,S128_256,S128_1k,M,MIX,D128,L8k
export VLLMINI_B200_NO_SHIM=1
VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/elect/libvllmini_b200.so timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -k "not experiments_build and not library and not speculative" 2>&1 | tail -2
for rep in 1 2 3; do
  for v in default elect; do
    if [ $v = default ]; then LIB=$PWD/vllmini_b200/lib/libvllmini_b200.so; else LIB=$PWD/vllmini_b200/lib/$v/libvllmini_b200.so; fi
    VLLMINI_B200_LIB=$LIB timeout 600 python tools/sweep.py --graph --configs $CFG --out gpurun_out/r3e_sweep_${v}_$rep.json > /dev/null 2>&1
    echo "$v rep $rep rc=$?"
  done
done
python - <<'PY'
import json
def load(v,r): return {x['config']:x['us_median'] for x in json.load(open(f'gpurun_out/r3e_sweep_{v}_{r}.json'))}
vs=['default','elect']
d={v:[load(v,r) for r in (1,2,3)] for v in vs}
for k in d['default'][0]:
    med={v:sorted(x[k] for x in d[v])[1] for v in vs}
    print(f"{k:10s}"+''.join(f"{med[v]:10.2f}" for v in vs)+f"   {med['elect']/med['default']:.3f}")
PY
