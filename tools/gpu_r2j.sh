#!/bin/bash
# round-2 pass J: cluster merge with lane-parallel remote reads; new parity tests (slot patterns, random shapes)
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2j_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2j_pytest_gpu.log
echo "== sweep (defaults), graph timing"
timeout 600 python tools/sweep.py --graph --configs S1_48,S4_64,S1_1k,S1_8k,S4_4k,S8_2k,S16_512,S32_1k,C1,M,MIX --out gpurun_out/r2j_sweep_default.json > gpurun_out/r2j_sweep_default.log 2>&1
echo "sweep rc=$?"; python - <<'PY'
import json
for r in json.load(open('gpurun_out/r2j_sweep_default.json')):
    print(f"{r['config']:10s} {r['us_median']:8.2f} us  {r['gbps']:8.1f} GB/s")
PY
