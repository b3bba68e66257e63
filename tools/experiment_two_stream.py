#!/usr/bin/env python
"""Experiment: split the decode batch into G groups on G streams inside one CUDA graph, so that the HBM-bound
attention of one group overlaps the latency-bound GEMM / LayerNorm / GELU launches of the others."""
import json
import os
import sys
import types

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
pos = torch.full((S,), CTX - 1, dtype=torch.int64, device=dev)
kv.begin_decode_step()


def group_view(lo, hi):
    v = types.SimpleNamespace()
    v.key_cache, v.value_cache = kv.key_cache, kv.value_cache
    v.block_size, v.max_blocks_per_seq = kv.block_size, kv.max_blocks_per_seq
    v.d_slots = [kv.d_slots[l][lo:hi] for l in range(L)]
    v.d_tables = [kv.d_tables[l][lo:hi] for l in range(L)]
    v.d_lens = kv.d_lens[lo:hi]
    return v


def run(groups, prio):
    bounds = [(i * S // groups, (i + 1) * S // groups) for i in range(groups)]
    views = [group_view(lo, hi) for lo, hi in bounds]
    streams = [torch.cuda.Stream(dev, priority=(-1 if (prio and i % 2) else 0)) for i in range(groups)]
    outs = [None] * groups

    def body():
        cur = torch.cuda.current_stream(dev)
        if groups == 1:
            outs[0] = model.greedy(model.decode_forward(ids, pos, views[0]))
            return
        for st in streams:
            st.wait_stream(cur)
        for i, (lo, hi) in enumerate(bounds):
            with torch.cuda.stream(streams[i]):
                outs[i] = model.greedy(model.decode_forward(ids[lo:hi], pos[lo:hi], views[i]))
        for st in streams:
            cur.wait_stream(st)

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            body()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        body()
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 30, torch.cat(outs)


res = {}
ref = None
for groups, prio in ((1, False), (2, False), (2, True), (3, False), (4, False), (4, True)):
    ms, toks = run(groups, prio)
    if ref is None:
        ref = toks
    res[f"groups={groups} prio={prio}"] = {"ms_per_step": ms, "tokens_per_s": S / ms * 1e3,
                                           "same_tokens": float((toks == ref).float().mean())}
    print(groups, prio, ms, float((toks == ref).float().mean()), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "two_stream.json"), "w"), indent=1)
