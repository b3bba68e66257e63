#!/usr/bin/env python
"""BASELINE configs[2] through the whole stack: GPT-2 small (seeded random weights), 256 rows, mixed contexts
128-2048, block 16, continuous batching by vllmini_b200.scheduler.BatchedScheduler (greedy, CUDA graph step).
Prints tokens/s over a window of decode steps, admissions (prefills) included."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vllmini_b200.decode import BatchedKVState, GPT2Decoder, GPT2Shape, synthetic_gpt2_state_dict  # noqa: E402
from vllmini_b200.scheduler import BatchedScheduler  # noqa: E402

dev = torch.device("cuda:0")
S, B, L, H, D = 256, 16, 12, 12, 64
MAXLEN = 2048
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
churn = (sys.argv[2] != "nochurn") if len(sys.argv) > 2 else True   # nochurn: nobody finishes inside the window
shape = GPT2Shape(n_positions=MAXLEN)
model = GPT2Decoder(shape, dev)
model.load_state_dict(synthetic_gpt2_state_dict(shape, seed=0))
MB = MAXLEN // B
kv = BatchedKVState(L * S * MB, L, H, D, B, S, MB, dev)
rng = np.random.default_rng(0)
sch = BatchedScheduler(model, kv, max_length=MAXLEN, use_graph=True)
t0 = time.perf_counter()
n_req = S + 64 if churn else S
for _ in range(n_req):
    n = int(rng.integers(128, 2041 if churn else 2041 - steps - 8))
    sch.add_sequence(rng.integers(0, shape.vocab_size, size=n).tolist())
torch.cuda.synchronize()
t_prefill = time.perf_counter() - t0
for _ in range(5):
    sch.step()
torch.cuda.synchronize()
tok0, fin0 = sch.tokens_generated, len(sch.finished)
ctx = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t1 = time.perf_counter()
e0.record()
for _ in range(steps):
    ctx.append(float(np.mean([sch.sequence_lengths[s] for s in sch.rows if s is not None])))
    sch.step()
e1.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t1
ms = e0.elapsed_time(e1)
res = {"config": "BASELINE configs[2]: GPT-2 small continuous batching, 256 rows, prompts U[128,2040], max_length 2048, "
                 "block 16, greedy, CUDA-graph step, " + ("admissions (bucket-padded, CUDA-graph prefill) inside the window" if churn else
                                                            "no request finishes inside the window"),
       "steps": steps, "tokens": sch.tokens_generated - tok0, "ms_per_step": ms / steps, "wall_ms_per_step": wall / steps * 1e3,
       "tokens_per_s": (sch.tokens_generated - tok0) / (ms * 1e-3), "mean_ctx": float(np.mean(ctx)),
       "finished_in_window": len(sch.finished) - fin0, "still_pending": len(sch.pending),
       "prefill_s_for_%d_prompts" % S: t_prefill}
print(json.dumps(res))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "continuous_batching%s.json" % ("" if churn else "_nochurn")), "w"),
          indent=1)
