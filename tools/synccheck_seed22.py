#!/usr/bin/env python
"""40 sequences of <= 900 tokens x 3 heads x 96, block 32, with a max_seq_len hint taken from the environment
(VB200_MSL): 1024 -> no split; 2048 / 3072 -> 2 / 3 splits of 1024 tokens of which all but the first are EMPTY for every
sequence.  For compute-sanitizer runs (tools/gpu_r3j.sh)."""
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

msl = int(os.environ.get("VB200_MSL", "3072"))
rng = np.random.default_rng(5)
S, H, D, B = 40, 3, 96, 32
lens = rng.integers(0, 901, size=S).astype(np.int32)
lens[3] = 900
q, kc, vc, bt, sl = po.make_case(S, H, D, B, lens, seed=9, max_blocks_per_seq=32)
truth = c_oracle().paged_attention_v1(q, kc, vc, D ** -0.5, bt, sl, mode="exact").astype(np.float64)
dev = torch.device("cuda:0")
t = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (q, kc, vc, bt, sl)]
out = torch.empty(S, H, D, dtype=torch.float16, device=dev)
print("plan:", _native.plan(S, H, H, D, B, msl), flush=True)
ops.paged_attention_v1(out, t[0], t[1], t[2], H, D ** -0.5, t[3], t[4], B, msl, None, "auto", 1.0, 0, 0, 1, 1, 0)
torch.cuda.synchronize()
assert_close_to_truth(out.cpu().numpy(), truth, atol=1.5e-3, what=f"msl {msl}")
print("ok", flush=True)
