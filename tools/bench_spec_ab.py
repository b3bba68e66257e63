#!/usr/bin/env python
"""A/B of the speculative first-ring prefetch inside bench.py's own legs (hot-path step, fused step, graph step, whole
model): runs bench.py in-process on the experiments build with the knob off / on.

    VLLMINI_B200_NO_SHIM=1 VLLMINI_B200_LIB=vllmini_b200/lib/experiments/libvllmini_b200.so python tools/bench_spec_ab.py 0|1
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vllmini_b200 import _native  # noqa: E402

on = int(sys.argv[1])
_native.set_spec_prefetch(bool(on))
sys.argv = ["bench.py", "--steps", "50", "--warmup", "5", "--no-configs", "--no-cpu-baseline", "--no-continuous"]
import bench  # noqa: E402

sys.exit(bench.main())
