#!/usr/bin/env python
"""What does a kernel boundary cost with programmatic dependent launch?  Three launches of one small shape are queued
behind a long launch (so the host is out of the picture) and every CTA records globaltimer stamps (-DVB200_TRACE
build): when does launch B's first CTA become resident, when is its griddepcontrol.wait released relative to launch
A's last CTA end, when does its first data land.

    python tools/trace_pdl_gap.py S16_512 [pdl=1]
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from sweep import CONFIGS, make  # noqa: E402

lib = ctypes.CDLL(os.environ.get("VB200_TRACE_LIB") or
                  os.path.join(ROOT, "vllmini_b200", "lib", "trace", "libvllmini_b200.so"))
vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
lib.vb200_paged_attention_v1_f16.argtypes = [vp] * 6 + [i32] * 7 + [i64] * 3 + [f32, i32, vp]
lib.vb200_debug_set_trace.argtypes = [vp]
lib.vb200_set_pdl.argtypes = [i32]


def launch(pool, cfg, tr, stream):
    S, H, D, B, ctx = cfg
    q, kc, vc, bt, sl = pool
    out = torch.empty_like(q)
    lib.vb200_debug_set_trace(tr)
    rc = lib.vb200_paged_attention_v1_f16(out.data_ptr(), q.data_ptr(), kc.data_ptr(), vc.data_ptr(), bt.data_ptr(),
                                          sl.data_ptr(), S, H, H, D, B, bt.shape[1], ctx, q.stride(0), kc.stride(0),
                                          kc.stride(1), D ** -0.5, 0, stream)
    assert rc == 0
    return out


def run(name, pdl):
    lib.vb200_set_pdl(pdl)
    dev = torch.device("cuda:0")
    cfg = CONFIGS[name][:5]
    S, H, D, B, ctx = cfg
    big_cfg = CONFIGS["M"][:5]
    big = make(*big_cfg, 99, dev)
    kv_bytes = 2 * S * ctx * H * D * 2
    npools = max(4, min(64, -(-(300 << 20) // kv_bytes)))
    pools = [make(S, H, D, B, ctx, s, dev) for s in range(npools)]
    stream = torch.cuda.current_stream(dev).cuda_stream
    nct = S * H * 16
    bufs = [torch.zeros(nct * 8, dtype=torch.int64, device=dev) for _ in range(3)]
    keep = []
    for i in range(npools):
        keep.append(launch(pools[i], cfg, None, stream))
    torch.cuda.synchronize()
    keep.append(launch(big, big_cfg, None, stream))          # ~115 us: the three traced launches queue up behind it
    for j in range(3):
        keep.append(launch(pools[j], cfg, bufs[j].data_ptr(), stream))
    torch.cuda.synchronize()
    ts = []
    for b in bufs:
        t = b.cpu().numpy().reshape(nct, 8).astype(np.int64)
        ts.append(t[t[:, 3] > 0])
    base = ts[0][:, 1].min()
    rep = {"config": name, "pdl": pdl}
    for j, t in enumerate(ts):
        us = lambda c: (t[t[:, c] > 0, c] - base) / 1e3
        rep["ABC"[j]] = {"first_cta_start": round(float(us(1).min()), 2), "last_cta_start": round(float(us(1).max()), 2),
                         "first_wait_released": round(float(us(4).min()), 2),
                         "first_data_p50": round(float(np.median(us(2))), 2),
                         "last_cta_end": round(float(us(3).max()), 2)}
    # launch B in detail: us after its griddepcontrol.wait was first released, p10 / p50 / p90 over its CTAs
    tb = ts[1]
    rel = tb[tb[:, 4] > 0, 4].min()
    def pct(c):
        v = (tb[tb[:, c] > 0, c] - rel) / 1e3
        return [round(float(np.percentile(v, q)), 2) for q in (10, 50, 90)] if len(v) else None
    rep["B after release"] = {"seq_len known": pct(4), "first stage landed": pct(2), "own stages done (warp 0)": pct(5),
                              "all warps done": pct(6), "cta end": pct(3)}
    for a, b in ("AB", "BC"):
        rep[f"{a}.last_end -> {b}.wait_released"] = round(rep[b]["first_wait_released"] - rep[a]["last_cta_end"], 2)
        rep[f"{a}.last_end -> {b}.last_end (period)"] = round(rep[b]["last_cta_end"] - rep[a]["last_cta_end"], 2)
    return rep


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "S16_512"
    for pdl in (1, 0):
        print(json.dumps(run(name, pdl)))
