#!/bin/bash
# pass 3a: prologue copies issued in stage order across the warps of a CTA (variant "ordered"), default 4x2 and 4x3 rings
set -u
mkdir -p gpurun_out
CFG=S1_48,S1_1k,S1_8k,S4_4k,S8_2k,S16_512,S32_1k,C1,S128_256,S128_1k,M,MIX,D128
export VLLMINI_B200_NO_SHIM=1
for rep in 1 2 3; do
  for v in experiments ordered; do
    for var in 0 1; do
      VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/$v/libvllmini_b200.so timeout 600 python tools/sweep.py --graph --configs $CFG --variants $var --out gpurun_out/r3a_sweep_${v}_v${var}_$rep.json > /dev/null 2>&1
      echo "$v variant $var rep $rep rc=$?"
    done
  done
done
python - <<'PY'
import json
def load(v,var,r): return {x['config']:x['us_median'] for x in json.load(open(f'gpurun_out/r3a_sweep_{v}_v{var}_{r}.json'))}
keys=[('experiments',0),('ordered',0),('experiments',1),('ordered',1)]
d={k:[load(k[0],k[1],r) for r in (1,2,3)] for k in keys}
print(f"{'config':10s}"+''.join(f"{(k[0][:3]+('4x3' if k[1] else '4x2')):>10s}" for k in keys)+'   ratios vs first')
for c in d[keys[0]][0]:
    med={k:sorted(x[c] for x in d[k])[1] for k in keys}
    print(f"{c:10s}"+''.join(f"{med[k]:10.2f}" for k in keys)+'   '+' '.join(f"{med[k]/med[keys[0]]:.3f}" for k in keys[1:]))
PY
