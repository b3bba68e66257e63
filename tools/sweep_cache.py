#!/usr/bin/env python
"""Timing of reshape_and_cache (new vs the reference's own kernel) for decode- and prefill-sized batches.

    python tools/sweep_cache.py [--out gpurun_out/sweep_cache.json]

CUDA-graph replay of 20 launches (no host overhead).  Decode: T = batch tokens to scattered slots;
prefill: T consecutive slots of one sequence.  Algorithmic bytes = 2 * T*H*D*2 read + same written.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from vllmini_b200 import ops  # noqa: E402


def run(fn, T, H, D, B, contiguous_slots, iters=20):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    NB = max((T + B - 1) // B * 4, 4096)
    qkv = torch.randn(T, 3 * H * D, generator=g, device=dev, dtype=torch.float16)
    k, v = qkv[:, H * D:2 * H * D].view(T, H, D), qkv[:, 2 * H * D:].view(T, H, D)
    kc = torch.zeros(NB, H, D // 8, B, 8, dtype=torch.float16, device=dev)
    vc = torch.zeros(NB, H, D, B, dtype=torch.float16, device=dev)
    if contiguous_slots:
        slots = torch.arange(T, device=dev, dtype=torch.int64) + 7 * B
    else:
        slots = (torch.randperm(NB, generator=g, device=dev)[:T] * B + torch.randint(0, B, (T,), generator=g, device=dev)).to(torch.int64)
    for _ in range(3):
        fn(k, v, kc, vc, slots, "auto", 1.0)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(iters):
            fn(k, v, kc, vc, slots, "auto", 1.0)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / iters)
    return sorted(ts)[2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_cache.json"))
    a = ap.parse_args()
    from helpers import load_ref_extension
    ref = load_ref_extension()
    res = []
    for name, T, contiguous in (("decode bs=256", 256, False), ("decode bs=2048", 2048, False),
                                ("prefill 1024 tok", 1024, True), ("prefill 16384 tok", 16384, True),
                                ("prefill 131072 tok", 131072, True)):
        for B in (16,):
            us_new = run(ops.reshape_and_cache, T, 12, 64, B, contiguous)
            us_ref = run(ref.cache_ops.reshape_and_cache, T, 12, 64, B, contiguous) if ref is not None else None
            byts = 4 * T * 12 * 64 * 2
            r = dict(case=name, T=T, block_size=B, us_new=us_new, us_ref=us_ref, gbps_new=byts / us_new / 1e3,
                     gbps_ref=(byts / us_ref / 1e3) if us_ref else None)
            res.append(r)
            print(json.dumps(r), flush=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
