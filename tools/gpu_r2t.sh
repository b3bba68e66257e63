#!/bin/bash
# round-2 pass T: wave hand-over prefetch A/B (knob bits: 1 = pre-wait speculation, 2 = hand-over)
set -u
mkdir -p gpurun_out
echo "== parity (attention tests) on the default build"
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -k "paged_attention or split or random or fused or full_size or baseline or extremes or v2" 2>&1 | tail -3
export VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so
CFG=S16_512,S32_1k,C1,S128_256,S128_1k,M,MIX,L1k_h20,L4k_b8,D128,D80,D256
for rep in 1 2 3; do
  for f in 0 1 2 3; do
    timeout 600 python tools/sweep.py --graph --configs $CFG --spec $f --out gpurun_out/r2t_sweep_f${f}_$rep.json > /dev/null 2>&1; echo "flags $f rep $rep rc=$?"
  done
done
python - <<'PY'
import json
def load(v,r): return {x['config']:x['us_median'] for x in json.load(open(f'gpurun_out/r2t_sweep_f{v}_{r}.json'))}
d={v:[load(v,r) for r in (1,2,3)] for v in (0,1,2,3)}
print(f"{'config':10s}{'off':>9s}{'spec':>9s}{'handover':>9s}{'both':>9s}   ratios vs off")
for k in d[0][0]:
    m={v:sorted(x[k] for x in d[v])[1] for v in d}
    print(f"{k:10s}"+''.join(f"{m[v]:9.2f}" for v in (0,1,2,3))+'   '+' '.join(f"{m[v]/m[0]:.3f}" for v in (1,2,3)))
PY
