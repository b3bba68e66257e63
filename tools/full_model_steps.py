#!/usr/bin/env python
"""Per-step timing of the whole-model decode loop of bench.py (full_model row): which steps are slow?"""
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

start = int(sys.argv[1]) if len(sys.argv) > 1 else None
args = types.SimpleNamespace(steps=50)
dev = torch.device("cuda:0")

# replicate bench.run_full_model with per-step events
from vllmini_b200.decode import BatchedKVState, GPT2Decoder, GPT2Shape, synthetic_gpt2_state_dict  # noqa: E402
S, H, D, B, CTX, LAYERS = bench.S, bench.H, bench.D, bench.B, bench.CTX, bench.LAYERS
shape = GPT2Shape(n_positions=2048)
model = GPT2Decoder(shape, dev)
model.load_state_dict(synthetic_gpt2_state_dict(shape, seed=0))
MB = 72
kv = BatchedKVState(LAYERS * S * MB, LAYERS, H, D, B, S, MB, dev)
g = torch.Generator(device=dev).manual_seed(99)
kv.key_cache.normal_(generator=g)
kv.value_cache.normal_(generator=g)
steps = 50
start_len = start if start is not None else CTX - steps // 2 - 4
nblk = (start_len + B - 1) // B
perm = torch.randperm(LAYERS * S * MB, generator=g, device=dev).to(torch.int32).cpu().numpy()
kv.h_tables[:, :, :nblk] = perm[: LAYERS * S * nblk].reshape(LAYERS, S, nblk)
kv.free_blocks = type(kv.free_blocks)(int(b) for b in perm[LAYERS * S * nblk:])
kv.h_lens[:] = start_len
kv.active[:] = True
kv.d_tables.copy_(torch.from_numpy(kv.h_tables))
ids = torch.randint(0, shape.vocab_size, (S,), generator=g, device=dev)
kv.begin_decode_step()
model.capture_decode_graph(kv, feed_back=True)
model.g_ids.copy_(ids)
evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 4)]
host = []
for i in range(steps + 3):
    evs[i].record()
    t0 = time.perf_counter()
    kv.begin_decode_step()
    t1 = time.perf_counter()
    model.replay_decode_graph()
    host.append(((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3))
evs[steps + 3].record()
torch.cuda.synchronize()
ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps + 3)]
print("start_len", start_len)
for i, m in enumerate(ms):
    pos = start_len + 1 + i
    print(f"step {i:2d} pos {pos} gpu {m:.3f} ms  host begin {host[i][0]:.3f} replay+copy {host[i][1]:.3f}" + ("  <- growth" if pos % B == 0 else ""))
print("mean of last 50:", sum(ms[3:]) / 50)
