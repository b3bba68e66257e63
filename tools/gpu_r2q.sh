#!/bin/bash
# round-2 pass Q: speculative L2 prefetch of the first ring ahead of griddepcontrol.wait (build variant "spec")
set -u
mkdir -p gpurun_out
echo "== is prefetch.global.L2 a pure hint? (wild addresses)"
tools/bin/prefetch_wild_repro; echo "rc=$?"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 99 tools/bin/prefetch_wild_repro 2>&1 | tail -3
echo "== parity on the spec build"
VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/spec/libvllmini_b200.so timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -k "not experiments_build and not library" 2>&1 | tail -3
echo "== A/B, graph timing, same box, interleaved"
CFG=S1_48,S1_1k,S1_8k,S4_4k,S8_2k,S16_512,S32_1k,C1,S128_1k,M,MIX,L1k_h20,D128
for rep in 1 2; do
  for v in default spec; do
    if [ $v = default ]; then LIB=$PWD/vllmini_b200/lib/libvllmini_b200.so; else LIB=$PWD/vllmini_b200/lib/spec/libvllmini_b200.so; fi
    VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$LIB timeout 600 python tools/sweep.py --graph --configs $CFG --out gpurun_out/r2q_sweep_${v}_$rep.json > gpurun_out/r2q_sweep_${v}_$rep.log 2>&1
    echo "sweep $v $rep rc=$?"
  done
done
python - <<'PY'
import json
def load(v,r): return {x['config']:x['us_median'] for x in json.load(open(f'gpurun_out/r2q_sweep_{v}_{r}.json'))}
d1,s1,d2,s2=load('default',1),load('spec',1),load('default',2),load('spec',2)
print(f"{'config':10s} {'default':>16s} {'spec':>16s}  ratio")
for k in d1: print(f"{k:10s} {d1[k]:7.2f} {d2[k]:7.2f}  {s1[k]:7.2f} {s2[k]:7.2f}  {(s1[k]+s2[k])/(d1[k]+d2[k]):.3f}")
PY
