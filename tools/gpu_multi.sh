#!/bin/bash
# multi-GPU bench the way the driver launches it: N ranks under torchrun, both arms.  Usage: gpu_multi.sh N [tag]
set -u
N=${1:-2}; TAG=${2:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref_${N}gpu.json 2> gpurun_out/${TAG}_bench_ref_${N}gpu.err
echo "reference arm rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_ref_${N}gpu.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')})
print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
print('clocks',d['clocks']); print('affinity',d['config'].get('cpu_affinity'))
print('full_model',{k:d['full_model'].get(k) for k in ('value','ms_per_step','weight_broadcast_s','replicas_agree_after_broadcast','error')})
print('continuous',{k:(d.get('continuous') or {}).get(k) for k in ('value','ms_per_step','requests','kv_tokens_per_rank_min_max','finished_in_window','still_pending','error')})
PY
