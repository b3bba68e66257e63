#!/bin/bash
# round-2 pass R: first-wave-only speculative L2 prefetch, 1 / 2 / 8 rings deep, A/B against the default build
set -u
mkdir -p gpurun_out
CFG=S1_48,S4_64,S1_1k,S1_8k,S4_4k,S8_2k,S16_512,S32_1k,C1,S128_256,S128_1k,M,MIX,L1k_h20,D128
for rep in 1 2 3; do
  for v in default spec1 spec2 spec8; do
    if [ $v = default ]; then LIB=$PWD/vllmini_b200/lib/libvllmini_b200.so; else LIB=$PWD/vllmini_b200/lib/$v/libvllmini_b200.so; fi
    VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$LIB timeout 600 python tools/sweep.py --graph --configs $CFG --out gpurun_out/r2r_sweep_${v}_$rep.json > gpurun_out/r2r_sweep_${v}_$rep.log 2>&1
    echo "sweep $v $rep rc=$?"
  done
done
python - <<'PY'
import json
def load(v,r): return {x['config']:x['us_median'] for x in json.load(open(f'gpurun_out/r2r_sweep_{v}_{r}.json'))}
vs=['default','spec1','spec2','spec8']
d={v:[load(v,r) for r in (1,2,3)] for v in vs}
print(f"{'config':10s}"+''.join(f"{v:>10s}" for v in vs)+'   ratios')
for k in d['default'][0]:
    med={v:sorted(x[k] for x in d[v])[1] for v in vs}
    print(f"{k:10s}"+''.join(f"{med[v]:10.2f}" for v in vs)+'   '+' '.join(f"{med[v]/med['default']:.3f}" for v in vs[1:]))
PY
