#!/usr/bin/env python
"""Where does the decode step go?  CUDA-graph replays of sub-sets of the bench step (B200)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vllmini_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
hp = bench.DecodeHotPath(dev, seed=1)
S, H, D, B, CTX, L = bench.S, bench.H, bench.D, bench.B, bench.CTX, bench.LAYERS


def attn(l):
    q, k, v = hp.views(hp.qkv, l)
    ops.paged_attention_v1(hp.out[l], q, hp.key_cache, hp.value_cache, H, hp.scale, hp.block_tables[l], hp.seq_lens,
                           B, CTX, None, "auto", 1.0, 0, 0, 1, 1, 0)


def resh(l):
    q, k, v = hp.views(hp.qkv, l)
    ops.reshape_and_cache(k, v, hp.key_cache, hp.value_cache, hp.slots[l], "auto", 1.0)


def fused(l):
    q, k, v = hp.views(hp.qkv, l)
    ops.paged_attention_v1_append(hp.out[l], q, k, v, hp.key_cache, hp.value_cache, hp.slots[l], H, hp.scale,
                                  hp.block_tables[l], hp.seq_lens, B, CTX)


neg_slots = torch.full_like(hp.slots, -1)


def fused_nostore(l):
    """fused launch with every slot = -1 (padding): same code path, no row stores."""
    q, k, v = hp.views(hp.qkv, l)
    ops.paged_attention_v1_append(hp.out[l], q, k, v, hp.key_cache, hp.value_cache, neg_slots[l], H, hp.scale,
                                  hp.block_tables[l], hp.seq_lens, B, CTX)


def fused_partial(frac):
    """fused launch where only every `frac`-th sequence appends a row (the others carry slot -1)."""
    sl = hp.slots.clone()
    mask = (torch.arange(S, device=dev) % frac) != 0
    sl[:, mask] = -1

    def fn(l):
        q, k, v = hp.views(hp.qkv, l)
        ops.paged_attention_v1_append(hp.out[l], q, k, v, hp.key_cache, hp.value_cache, sl[l], H, hp.scale,
                                      hp.block_tables[l], hp.seq_lens, B, CTX)
    return fn


def timed(fn, reps=30):
    for l in range(L):
        fn(l)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for l in range(L):
            fn(l)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps / L


res = {
    "attention_only_us_per_layer": timed(attn),
    "reshape_only_us_per_layer": timed(resh),
    "reshape_then_attention_us_per_layer": timed(lambda l: (resh(l), attn(l))),
    "fused_append_us_per_layer": timed(fused),
    "fused_append_no_row_store_us_per_layer": timed(fused_nostore),
    "fused_append_every_2nd_seq_us_per_layer": timed(fused_partial(2)),
    "fused_append_every_8th_seq_us_per_layer": timed(fused_partial(8)),
    "fused_append_every_64th_seq_us_per_layer": timed(fused_partial(64)),
}
print(json.dumps(res))
