#!/bin/bash
# round-2 pass S: speculative first-ring L2 prefetch in the shipped kernel -- tests, bench, A/B with the knob, memcheck
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2s_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2s_pytest_gpu.log
echo "== bench"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r2s_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','decode_tokens_per_s_full_model')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('roofline', d['roofline']['frac'], d['roofline']['mean_launch_us'], 'fused', d['fused_step']['ms_per_step'], 'graph', d['graph_step']['ms_per_step'])
for k,v in d['configs'].items(): print('  ',k, v['new_us'], v.get('ref_us'), v.get('speedup'), v['frac_of_8TBps'])
print('continuous', d['continuous']['value'])
PY
echo "== A/B with the experiments knob (graph timing, 3 interleaved reps)"
export VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=$PWD/vllmini_b200/lib/experiments/libvllmini_b200.so
CFG=S1_48,S4_64,S1_1k,S1_8k,S4_4k,S8_2k,S16_512,S32_1k,C1,S128_256,S128_1k,M,MIX,L1k_h20,L8k,D128,D80,D256
for rep in 1 2 3; do
  timeout 600 python tools/sweep.py --graph --configs $CFG --no-spec --out gpurun_out/r2s_sweep_nospec_$rep.json > /dev/null 2>&1; echo "nospec $rep rc=$?"
  timeout 600 python tools/sweep.py --graph --configs $CFG --out gpurun_out/r2s_sweep_spec_$rep.json > /dev/null 2>&1; echo "spec $rep rc=$?"
done
python - <<'PY'
import json
def load(v,r): return {x['config']:x['us_median'] for x in json.load(open(f'gpurun_out/r2s_sweep_{v}_{r}.json'))}
d={v:[load(v,r) for r in (1,2,3)] for v in ('nospec','spec')}
print(f"{'config':10s}{'off':>10s}{'on':>10s}  ratio")
for k in d['spec'][0]:
    m={v:sorted(x[k] for x in d[v])[1] for v in d}
    print(f"{k:10s}{m['nospec']:10.2f}{m['spec']:10.2f}  {m['spec']/m['nospec']:.3f}")
PY
unset VLLMINI_B200_NO_SHIM VLLMINI_B200_LIB
echo "== memcheck: split / random shapes / stale-index speculation"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --launch-timeout 300 --target-processes all python -m pytest tests/test_parity_gpu.py -q -m gpu -k "split_kv or random_shapes or speculative or fused_append" -p no:cacheprovider > gpurun_out/r2s_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2s_memcheck.log | tail -3
