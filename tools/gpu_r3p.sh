#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -k "first_use_inside or host_threads or speculative" 2>&1 | tail -15
