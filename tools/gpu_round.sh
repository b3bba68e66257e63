#!/bin/bash
# Standard GPU pass: parity suite, smoke, bench (both arms), ncu launch list + full capture.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv 2>/dev/null &
SMI=$!
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
kill $SMI 2>/dev/null
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref rc=$?"; cat gpurun_out/bench_ref.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-model --no-extras --no-configs > gpurun_out/ncu_launch_bench.log 2>&1
echo "ncu launches rc=$?"; tail -2 gpurun_out/ncu_launch_bench.log | cut -c1-300
echo "== ncu full capture of the attention kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pa_v1_kernel -s 40 -c 3 -f -o gpurun_out/prof_pa_v1 \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-model --no-extras --no-configs > gpurun_out/ncu_full_bench.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
fi
