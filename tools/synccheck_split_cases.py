#!/usr/bin/env python
"""One split-KV launch with a forced split factor (= cluster size), checked against the C oracle -- meant to run under
`compute-sanitizer --tool synccheck`, one factor per process (experiments build: VLLMINI_B200_LIB=.../lib/experiments/...).

    python tools/synccheck_split_cases.py <splits> [<seq_len> ...]      splits = 0: the launcher's own factor
    VB200_MAX_CLUSTER=8 python tools/synccheck_split_cases.py 0 8192 5000   cap the cluster size (experiments build)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import assert_close_to_truth  # noqa: E402
from oracle import c_oracle  # noqa: E402
from oracle import paged_oracle as po  # noqa: E402
from vllmini_b200 import _native, ops  # noqa: E402

splits = int(sys.argv[1])
lens = np.array([int(x) for x in sys.argv[2:]] or [4096, 3000], dtype=np.int32)
q, kc, vc, bt, sl = po.make_case(len(lens), 12, 64, 16, lens, seed=21)
truth = c_oracle().paged_attention_v1(q, kc, vc, 0.125, bt, sl, mode="exact").astype(np.float64)
dev = torch.device("cuda:0")
t = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (q, kc, vc, bt, sl)]
out = torch.empty(len(lens), 12, 64, dtype=torch.float16, device=dev)
cap = int(os.environ.get("VB200_MAX_CLUSTER", "16"))
if splits > 0 or cap != 16:
    assert _native.has_experiments(), "forced split factors / cluster caps need the experiments build"
    _native.set_tuning(splits, 0)
    _native.set_merge(0, cap)
msl = int(bt.shape[1]) * 16
print("plan:", _native.plan(len(lens), 12, 12, 64, 16, msl), flush=True)
ops.paged_attention_v1(out, t[0], t[1], t[2], 12, 0.125, t[3], t[4], 16, msl, None, "auto", 1.0, 0, 0, 1, 1, 0)
torch.cuda.synchronize()
assert_close_to_truth(out.cpu().numpy(), truth, what=f"forced splits {splits}")
print("result ok", flush=True)
