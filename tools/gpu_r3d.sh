#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r3d_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r3d_pytest_gpu.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -k "host_threads" 2>&1 | tail -1; done
