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
    NB = max((T + B - 1) // B * 4, 4096) if contiguous_slots else max(2 * T, 4096)   # scattered: one block per token
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


def check_modes(modes, set_mode):
    """Every kernel mode must write exactly what the reference kernel writes (bit-exact), for scattered, contiguous,
    padded (-1) and ragged batches, head sizes 64..256, strided rows."""
    from helpers import load_ref_extension
    ref = load_ref_extension()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    bad = 0
    for (T, H, D, B, kind) in [(1, 12, 64, 16, "scatter"), (17, 12, 64, 16, "contig"), (256, 12, 64, 16, "scatter"),
                               (300, 3, 80, 8, "contig"), (1000, 2, 256, 32, "scatter"), (1000, 2, 256, 32, "contig"),
                               (777, 5, 128, 16, "mixed"), (2049, 12, 64, 16, "pad"), (4100, 4, 112, 8, "contig")]:
        NB = max(2 * T, 64)
        qkv = torch.randn(T, 3 * H * D, generator=g, device=dev, dtype=torch.float16)
        k, v = qkv[:, H * D:2 * H * D].view(T, H, D), qkv[:, 2 * H * D:].view(T, H, D)
        if kind == "contig":
            slots = torch.arange(T, device=dev, dtype=torch.int64) + 3 * B + 5
        else:
            slots = (torch.randperm(NB, generator=g, device=dev)[:T] * B + torch.randint(0, B, (T,), generator=g, device=dev)).to(torch.int64)
            if kind == "mixed":
                slots[100:400] = torch.arange(300, device=dev, dtype=torch.int64) + (NB + 7) * B
            if kind == "pad":
                slots[::3] = -1
        NBX = NB + 400
        want_k = torch.zeros(NBX, H, D // 8, B, 8, dtype=torch.float16, device=dev)
        want_v = torch.zeros(NBX, H, D, B, dtype=torch.float16, device=dev)
        ref.cache_ops.reshape_and_cache(k, v, want_k, want_v, slots, "auto", 1.0)
        for mode in modes:
            set_mode(mode)
            kc, vc = torch.zeros_like(want_k), torch.zeros_like(want_v)
            ops.reshape_and_cache(k, v, kc, vc, slots, "auto", 1.0)
            torch.cuda.synchronize()
            ok = torch.equal(kc, want_k) and torch.equal(vc, want_v)
            bad += not ok
            print(f"check T={T} H={H} D={D} B={B} {kind} mode={mode}: {'ok' if ok else 'MISMATCH'}", flush=True)
    set_mode(0)
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true", help="bit-exactness of every mode against the reference kernel")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_cache.json"))
    ap.add_argument("--modes", default="0", help="experiments build only: 0 shipped dispatch, 1 warp kernel, "
                                                 "2 16-token tiles, 3 32-token tiles")
    a = ap.parse_args()
    from helpers import load_ref_extension
    from vllmini_b200 import _native
    ref = load_ref_extension()
    modes = [int(m) for m in a.modes.split(",")]
    set_mode = None
    if modes != [0]:
        set_mode = _native.load().vb200_debug_set_cache_kernel       # AttributeError on the shipped library
    if a.check:
        bad = check_modes(modes, set_mode or (lambda m: None))
        print("check:", "all ok" if not bad else f"{bad} MISMATCHES")
        if bad:
            sys.exit(1)
    cases = [(f"{'prefill' if c else 'decode'} T={T}", T, c) for T in (256, 512, 1024, 2048, 4096, 16384)
             for c in (False, True)] + [("prefill T=131072", 131072, True)]
    res = []
    for name, T, contiguous in cases:
        us_ref = run(ref.cache_ops.reshape_and_cache, T, 12, 64, 16, contiguous) if ref is not None else None
        for mode in modes:
            if set_mode:
                set_mode(mode)
            us_new = run(ops.reshape_and_cache, T, 12, 64, 16, contiguous)
            byts = 4 * T * 12 * 64 * 2
            r = dict(case=name, T=T, contiguous=contiguous, mode=mode, us_new=us_new, us_ref=us_ref,
                     speedup=(us_ref / us_new) if us_ref else None, gbps_new=byts / us_new / 1e3)
            res.append(r)
            print(json.dumps(r), flush=True)
    if set_mode:
        set_mode(0)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
