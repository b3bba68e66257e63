#!/usr/bin/env python
"""Kernel-level breakdown of one whole GPT-2 small decode step (bs=256, ctx~1024) driven by
vllmini_b200.decode -- which kernels the time outside the two hot-path ops goes to (B200, torch.profiler/CUPTI)."""
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vllmini_b200.decode import BatchedKVState, GPT2Decoder, GPT2Shape, synthetic_gpt2_state_dict  # noqa: E402

dev = torch.device("cuda:0")
S, H, D, B, CTX, L = bench.S, bench.H, bench.D, bench.B, bench.CTX, bench.LAYERS
shape = GPT2Shape(n_positions=2048)
model = GPT2Decoder(shape, dev)
model.load_state_dict(synthetic_gpt2_state_dict(shape, seed=0))
MB = 72
kv = BatchedKVState(L * S * MB, L, H, D, B, S, MB, dev)
g = torch.Generator(device=dev).manual_seed(1)
kv.key_cache.normal_(generator=g)
kv.value_cache.normal_(generator=g)
nblk = CTX // B
perm = torch.randperm(L * S * MB, generator=g, device=dev).to(torch.int32).cpu().numpy()
kv.h_tables[:, :, :nblk] = perm[: L * S * nblk].reshape(L, S, nblk)
kv.free_blocks = type(kv.free_blocks)(int(b) for b in perm[L * S * nblk:])
kv.h_lens[:] = CTX - 1
kv.active[:] = True
kv.d_tables.copy_(torch.from_numpy(kv.h_tables))
ids = torch.randint(0, shape.vocab_size, (S,), generator=g, device=dev)


def step():
    kv.begin_decode_step()
    pos = torch.from_numpy(kv.positions()).to(dev)
    return model.greedy(model.decode_forward(ids, pos, kv))


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = e.name[:90]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
rows = sorted(agg.items(), key=lambda kv_: -kv_[1][1])
tot = sum(v[1] for _, v in rows) / 3
print(f"kernel time per decode step: {tot:.1f} us")
for k, (n, t) in rows:
    print(f"{t / 3:9.1f} us  x{n // 3:<4d} {k}")
json.dump({k: {"count_per_step": n // 3, "us_per_step": t / 3} for k, (n, t) in rows},
          open(os.path.join(ROOT, "gpurun_out", "model_breakdown.json"), "w"), indent=1)
