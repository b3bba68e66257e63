#!/usr/bin/env python
"""Micro-benchmark: the four projection GEMMs of a GPT-2 small decode step at M=256 through different torch/cuBLAS
entry points (graph replay over 12 distinct weight sets, us per GEMM)."""
import json
import os

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda:0")
M = 256
shapes = {"c_attn": (768, 2304), "attn.c_proj": (768, 768), "c_fc": (768, 3072), "mlp.c_proj": (3072, 768)}
res = {}


def timed(fn, n=12, reps=20):
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps / n


for name, (K, N) in shapes.items():
    x = torch.randn(M, K, device=dev, dtype=torch.float16)
    W = [torch.randn(N, K, device=dev, dtype=torch.float16) * 0.02 for _ in range(12)]      # nn.Linear layout [out, in]
    Wt = [w.t().contiguous() for w in W]                                                       # [in, out]
    b = [torch.randn(N, device=dev, dtype=torch.float16) for _ in range(12)]
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    r = {}
    r["F.linear(x, W, b)"] = timed(lambda i: F.linear(x, W[i], b[i]))
    r["F.linear(x, W) no bias"] = timed(lambda i: F.linear(x, W[i]))
    r["addmm(b, x, Wt_contig)"] = timed(lambda i: torch.addmm(b[i], x, Wt[i]))
    r["addmm(b, x, W.t())"] = timed(lambda i: torch.addmm(b[i], x, W[i].t()))
    r["matmul(x, Wt_contig)"] = timed(lambda i: torch.matmul(x, Wt[i]))
    r["mm out="] = timed(lambda i: torch.mm(x, Wt[i], out=out))
    res[name] = r
    print(name, json.dumps({k: round(v, 2) for k, v in r.items()}), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_variants.json"), "w"), indent=1)
