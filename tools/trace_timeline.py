#!/usr/bin/env python
"""CTA timeline of one paged_attention_v1 launch.  Needs a -DVB200_TRACE build of the library:

    python -m vllmini_b200.build --variant trace --define=-DVB200_TRACE
    python tools/trace_timeline.py M C1            # or VB200_TRACE_LIB=<path to another trace build>

Prints where the launch's wall time goes: ramp-up, steady state, tail; per-CTA prologue and lifetime."""
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
vp, i32, i64, f32, sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
lib.vb200_paged_attention_v1_f16.argtypes = [vp] * 6 + [i32] * 7 + [i64] * 3 + [f32, i32, vp]
lib.vb200_debug_set_trace.argtypes = [vp]
lib.vb200_set_pdl.argtypes = [i32]


def run(name):
    S, H, D, B, ctx = CONFIGS[name]
    mixed = ctx < 0
    ctx = abs(ctx)
    dev = torch.device("cuda:0")
    kv_bytes = 2 * S * ctx * H * D * 2
    npools = max(2, min(64, -(-(300 << 20) // kv_bytes)))     # > 2x L2 in total: the traced launch reads from HBM
    pools = [make(S, H, D, B, ctx, s, dev, mixed=mixed) for s in range(npools)]
    nctas = S * H * 16          # room for a split launch (grid.z <= 16)
    trace = torch.zeros(nctas * 8, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def call(i, tr):
        q, kc, vc, bt, sl = pools[i % npools]
        out = torch.empty_like(q)
        lib.vb200_debug_set_trace(tr)
        rc = lib.vb200_paged_attention_v1_f16(out.data_ptr(), q.data_ptr(), kc.data_ptr(), vc.data_ptr(), bt.data_ptr(),
                                              sl.data_ptr(), S, H, H, D, B, bt.shape[1], ctx, q.stride(0), kc.stride(0),
                                              kc.stride(1), D ** -0.5, 0, stream)
        assert rc == 0
    for i in range(1, npools + 1):
        call(i, None)
    torch.cuda.synchronize()
    call(0, trace.data_ptr())
    torch.cuda.synchronize()
    t = trace.cpu().numpy().reshape(nctas, 8).astype(np.int64)
    t = t[t[:, 3] > 0]          # CTAs that ran (and recorded an end time)
    nctas = len(t)
    t[:, 2] = np.where(t[:, 2] > 0, t[:, 2], t[:, 1])    # a CTA without stages has no first-data time
    smid, t0, t1, t2 = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
    base = t0.min()
    t0, t1, t2 = (t0 - base) / 1e3, (t1 - base) / 1e3, (t2 - base) / 1e3     # us
    span = t2.max()
    life = t2 - t0
    prol = t1 - t0
    # concurrency over time
    ev = np.concatenate([np.stack([t0, np.ones_like(t0)], 1), np.stack([t2, -np.ones_like(t2)], 1)])
    ev = ev[np.argsort(ev[:, 0])]
    conc = np.cumsum(ev[:, 1])
    full = conc.max()
    # time until concurrency first reaches 95 % of max; time after it last drops below 95 %
    hi = np.nonzero(conc >= 0.95 * full)[0]
    ramp_end, tail_start = ev[hi[0], 0], ev[hi[-1] + 1, 0] if hi[-1] + 1 < len(ev) else span
    rep = {
        "config": name, "ctas": int(nctas), "span_us": float(span), "max_concurrent_ctas": int(full),
        "ramp_until_95pct_resident_us": float(ramp_end), "tail_below_95pct_us": float(span - tail_start),
        "first_data_after_start_us": {"p10": float(np.percentile(prol, 10)), "p50": float(np.percentile(prol, 50)),
                                       "p90": float(np.percentile(prol, 90))},
        "cta_lifetime_us": {"p10": float(np.percentile(life, 10)), "p50": float(np.percentile(life, 50)),
                            "p90": float(np.percentile(life, 90)), "max": float(life.max())},
        "first_wave_first_data_us": float(np.percentile(t1[t0 < 1.0], 50)),
        "last_cta_start_us": float(t0.max()), "sms_used": int(len(np.unique(smid))),
        "cta_seconds_total_us": float(life.sum()), "avg_concurrency": float(life.sum() / span),
    }
    # idle SM-slot time in the tail: for each SM, time between its last CTA end and the span end
    last_end = np.array([t2[smid == s].max() for s in np.unique(smid)])
    rep["sm_idle_at_end_us"] = {"mean": float((span - last_end).mean()), "max": float((span - last_end).max())}
    # phases of a CTA (us after the launch's first CTA start; p10 / p50 / p90 / max over the CTAs that have the stamp)
    def pct(x):
        return [round(float(np.percentile(x, q)), 2) for q in (10, 50, 90, 100)] if len(x) else None
    def rel(col, ref=None):
        ok = t[:, col] > 0
        v = (t[ok, col] - base) / 1e3
        return pct(v if ref is None else v - ref[ok])
    rep["phases_us_p10_p50_p90_max"] = {
        "cta_start": pct(t0), "range_known (seq_len read)": rel(4), "first_stage_landed": pct(t1),
        "own_stages_done (warp 0)": rel(5), "all_warps_done": rel(6), "after_first_cluster_barrier": rel(7),
        "cta_end": pct(t2),
        "start->range_known": rel(4, t0), "range_known->first_stage": pct(t1[t[:, 4] > 0] - (t[t[:, 4] > 0, 4] - base) / 1e3),
        "first_stage->all_warps_done": pct((t[t[:, 6] > 0, 6] - base) / 1e3 - t1[t[:, 6] > 0]),
        "all_warps_done->cluster_barrier": rel(7, (t[:, 6] - base) / 1e3) if (t[:, 7] > 0).any() else None,
        "cluster_barrier->end": pct(t2[t[:, 7] > 0] - (t[t[:, 7] > 0, 7] - base) / 1e3) if (t[:, 7] > 0).any() else None,
    }
    return rep


if __name__ == "__main__":
    names = sys.argv[1:] or ["M"]
    for pdl in (0,):
        lib.vb200_set_pdl(pdl)
        for n in names:
            print(json.dumps(run(n)))
