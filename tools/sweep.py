#!/usr/bin/env python
"""Kernel-only timing sweep over pipeline variants / split factors (development tool, GPU box).

    python tools/sweep.py [--configs M,C1,...] [--out gpurun_out/sweep.json]

Prints achieved algorithmic GB/s per (config, variant, splits).  Inputs exceed L2 (pool >= 1 GiB
or the workload itself), CUDA-event timing over back-to-back launches rotating distinct pools.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vllmini_b200 import _native, ops  # noqa: E402

CONFIGS = {
    # name: (S, H, D, B, ctx)
    "M": (256, 12, 64, 16, 1024),
    "C1": (64, 12, 64, 16, 512),
    "L8k": (256, 20, 64, 16, 8192),
    "L4k_b8": (256, 20, 64, 8, 4096),
    "L4k_b32": (256, 20, 64, 32, 4096),
    "L1k_h20": (256, 20, 64, 16, 1024),
    "L1k_h20_b8": (256, 20, 64, 8, 1024),
    "L1k_h20_b32": (256, 20, 64, 32, 1024),
    "L8k_b8": (256, 20, 64, 8, 8192),
    "L8k_b32": (256, 20, 64, 32, 8192),
    "MIX": (256, 12, 64, 16, -2048),     # ctx ~ U[128, 2048] (BASELINE configs[2])
    "D128": (256, 8, 128, 16, 1024),
    "D80": (256, 8, 80, 16, 1024),       # the reference launcher's other head sizes
    "D96": (256, 8, 96, 16, 1024),
    "D112": (256, 8, 112, 16, 1024),
    "D192": (256, 4, 192, 16, 1024),
    "D256": (256, 4, 256, 16, 1024),
    "D80_b8": (256, 8, 80, 8, 1024),
    "D256_b32": (256, 4, 256, 32, 1024),
    "GQA4_d128": (256, 8, 128, 16, 1024, 32),   # 32 query heads over 8 kv heads (6th entry = query heads)
    "GQA8_d128": (256, 4, 128, 16, 1024, 32),
    "GQA4_d64": (256, 3, 64, 16, 1024, 12),
    "GQA4_d128_s1_8k": (1, 8, 128, 16, 8192, 32),     # single-stream decode, long context
    "GQA4_d128_s8_4k": (8, 8, 128, 16, 4096, 32),
    "GQA4_d128_s16": (16, 8, 128, 16, 2048, 32),
    "GQA4_d128_s32": (32, 8, 128, 16, 2048, 32),
    "GQA4_d128_s64": (64, 8, 128, 16, 1024, 32),
    "GQA4_d128_s128": (128, 8, 128, 16, 1024, 32),
    "GQA4_d64_s32": (32, 3, 64, 16, 2048, 12),
    "GQA4_d64_s128": (128, 3, 64, 16, 1024, 12),
    "S1_8k": (1, 12, 64, 16, 8192),
    "S1_48": (1, 12, 64, 16, 48),        # vllmini's own serving shape: 1 sequence, <= 64 tokens (server.py:37-50)
    "S4_64": (4, 12, 64, 16, 64),
    "S1_1k": (1, 12, 64, 16, 1024),
    "S4_4k": (4, 12, 64, 16, 4096),
    "S16_512": (16, 12, 64, 16, 512),
    "S32_1k": (32, 12, 64, 16, 1024),
    "S128_256": (128, 12, 64, 16, 256),
    "S128_1k": (128, 12, 64, 16, 1024),
    "S8_2k": (8, 12, 64, 16, 2048),
    # small batches at head sizes that hold 2 / 1 CTAs per SM (split factor capped by the resident slots)
    "D128_s16_1k": (16, 8, 128, 16, 1024),
    "D96_s40_1k": (40, 3, 96, 32, 1024),
    "D256_s10_2k": (10, 10, 256, 16, 2048),
    "D112_s2_3k": (2, 12, 112, 8, 3008, 24),
    "D192_s6_2k": (6, 8, 192, 16, 2048),
}


def algorithmic_bytes(S, H, D, B, lens):
    kv = sum(2 * int(l) * H * D * 2 for l in lens)
    return kv + 2 * S * H * D * 2 + sum(((int(l) + B - 1) // B) * 4 for l in lens) + S * 4


POOL_MULT = 1   # --pool-mult: KV pool this many times larger than the batch needs (block ids spread over it)


def make(S, H, D, B, ctx, seed, dev, mixed=False, QH=None):
    g = torch.Generator(device=dev).manual_seed(seed)
    if mixed:
        lens = torch.randint(128, 2049, (S,), generator=g, device=dev, dtype=torch.int32)
    else:
        lens = torch.full((S,), ctx, dtype=torch.int32, device=dev)
    nblk = (int(lens.max()) + B - 1) // B
    NB = S * nblk * POOL_MULT
    kc = torch.empty(NB, H, D // 8, B, 8, device=dev, dtype=torch.float16).normal_(generator=g)
    vc = torch.empty(NB, H, D, B, device=dev, dtype=torch.float16).normal_(generator=g)
    q = torch.randn(S, QH or H, D, generator=g, device=dev, dtype=torch.float16)
    bt = torch.randperm(NB, generator=g, device=dev)[: S * nblk].to(torch.int32).view(S, nblk).contiguous()
    return q, kc, vc, bt, lens


def time_config(name, variant, splits, iters=20, warm=5, ref=None, graph=False, merge=0, max_cluster=16):
    S, H, D, B, ctx = CONFIGS[name][:5]
    QH = CONFIGS[name][5] if len(CONFIGS[name]) > 5 else None
    mixed = ctx < 0
    ctx = abs(ctx)
    dev = torch.device("cuda:0")
    bytes_one = 2 * S * ctx * H * D * 2
    npools = max(2, min(6, int((1 << 30) // max(bytes_one, 1)) + 1)) if bytes_one < (1 << 30) else 2
    pools = [make(S, H, D, B, ctx, s, dev, mixed=mixed, QH=QH) for s in range(npools)]
    outs = [torch.empty_like(p[0]) for p in pools]
    if splits or variant:          # sweep knobs: needs the experiments build (VLLMINI_B200_LIB=.../lib/experiments/...)
        _native.set_tuning(splits, variant)
    if merge or max_cluster != 16:
        _native.set_merge(merge, max_cluster)
    def call(i):
        q, kc, vc, bt, sl = pools[i % npools]
        if ref is None:
            ops.paged_attention_v1(outs[i % npools], q, kc, vc, H, D ** -0.5, bt, sl, B, ctx, None, "auto", 1.0, 0, 0, 1, 1, 0)
        else:
            ref.paged_attention_v1(outs[i % npools], q, kc, vc, H, D ** -0.5, bt, sl, B, ctx, None, "auto", 1.0, 0, 0, 1, 1, 0)
    for i in range(warm):
        call(i)
    torch.cuda.synchronize()
    if graph:
        # host launch overhead removed: `iters` back-to-back launches captured in one CUDA graph
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(iters):
                call(i)
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / iters)
        ts.sort()
        del g
    else:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for i in range(iters):
            evs[i][0].record()
            call(i)
            evs[i][1].record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
    med = ts[len(ts) // 2]
    ab = sum(algorithmic_bytes(S, H, D, B, p[4].tolist()) for p in pools) / npools
    if splits or variant:
        _native.set_tuning(0, 0)
    if merge or max_cluster != 16:
        _native.set_merge(0, 16)
    del pools, outs
    torch.cuda.empty_cache()
    return dict(config=name, variant=variant, splits=splits, merge="workspace" if merge else "cluster",
                max_cluster=max_cluster, us_median=med, us_min=ts[0], gbps=ab / med / 1e3,
                bytes=ab, impl="ref" if ref is not None else "new", timing="graph" if graph else "events")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="M,C1")
    ap.add_argument("--variants", default="0", help="CTA-shape variants (non-zero: experiments build only)")
    ap.add_argument("--splits", default="0")
    ap.add_argument("--ref", action="store_true", help="also time the reference's own kernel (oracle/_ref)")
    ap.add_argument("--pool-mult", type=int, default=1)
    ap.add_argument("--graph", action="store_true", help="time launches replayed from a CUDA graph (no host overhead)")
    ap.add_argument("--merge", default="0", help="experiments build: 0 cluster merge, 1 workspace merge (comma list)")
    ap.add_argument("--max-cluster", default="16", help="experiments build: cap on the cluster size (comma list)")
    ap.add_argument("--no-spec", action="store_true", help="experiments build: speculative first-ring prefetch off")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    a = ap.parse_args()
    if a.no_spec:
        _native.set_spec_prefetch(False)
    global POOL_MULT
    POOL_MULT = a.pool_mult
    res = []
    ref = None
    if a.ref:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import load_ref_extension
        ref = load_ref_extension()
    for name in a.configs.split(","):
        for v in [int(x) for x in a.variants.split(",")]:
            for sp in [int(x) for x in a.splits.split(",")]:
                for mg in [int(x) for x in a.merge.split(",")]:
                    for mc in [int(x) for x in a.max_cluster.split(",")]:
                        if mg == 0 and sp > mc:
                            continue          # a cluster cannot be larger than its cap
                        r = time_config(name, v, sp, graph=a.graph, merge=mg, max_cluster=mc)
                        res.append(r)
                        print(json.dumps(r), flush=True)
        if ref is not None:
            r = time_config(name, 0, 0, ref=ref, graph=a.graph)
            res.append(r)
            print(json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
