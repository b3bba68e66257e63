#!/bin/bash
# pass 3f: elect-based copy issue adopted -- full suite, GQA / head-size regression sweep vs the numbers of r02i, racecheck
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r3f_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r3f_pytest_gpu.log
timeout 600 python tools/sweep.py --graph --configs GQA4_d128,GQA8_d128,GQA4_d64,D128,D80,D96,D112,D192,D256,D80_b8,D256_b32,L1k_h20_b8,L1k_h20_b32,L8k_b8,L8k_b32 --out gpurun_out/r3f_sweep_regress.json > gpurun_out/r3f_sweep_regress.log 2>&1
python - <<'PY'
import json
for r in json.load(open('gpurun_out/r3f_sweep_regress.json')): print(f"{r['config']:14s} {r['us_median']:9.2f} us {r['gbps']:8.1f} GB/s")
PY
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3f_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','decode_tokens_per_s_full_model')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
for k,v in d['configs'].items(): print('  ',k, v['new_us'], v.get('ref_us'), v.get('speedup'), v['frac_of_8TBps'])
PY
for tool in racecheck synccheck; do
timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 --launch-timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "(split_kv or v2 or fused_append or grouped or vs_oracle) and not 32768 and not small_batch_long and not two_streams" -p no:cacheprovider > gpurun_out/r3f_$tool.log 2>&1
echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r3f_$tool.log | tail -2
done
