#!/bin/bash
set -u
mkdir -p gpurun_out
export VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so
for rep in 1 2 3; do for on in 0 1; do
  timeout 600 python tools/bench_spec_ab.py $on > gpurun_out/r2v_bench_spec${on}_$rep.json 2> gpurun_out/r2v_err.log; echo "spec=$on rep=$rep rc=$?"
done; done
python - <<'PY'
import json
def g(on,rep):
    d=json.load(open(f'gpurun_out/r2v_bench_spec{on}_{rep}.json'))
    return dict(step=d['ms_per_step'], fused=d['fused_step']['ms_per_step'], graph=d['graph_step']['ms_per_step'], e2e=d['e2e']['ms_per_step'], full=d['full_model']['ms_per_step'])
for k in ('step','fused','graph','e2e','full'):
    off=sorted(g(0,r)[k] for r in (1,2,3)); on=sorted(g(1,r)[k] for r in (1,2,3))
    print(f"{k:6s} off {off}  on {on}  median ratio {on[1]/off[1]:.4f}")
PY
