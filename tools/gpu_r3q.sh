#!/bin/bash
set -u
mkdir -p gpurun_out
for i in 1 2; do timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r3q_pytest_gpu_$i.log 2>&1; echo "run $i rc=$?"; tail -1 gpurun_out/r3q_pytest_gpu_$i.log; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
