#!/usr/bin/env python
"""Generate tests/golden/refkernel_outputs.npz: outputs of the REFERENCE'S OWN CUDA kernels.

Run on a B200 box (needs a GPU and oracle/_ref/, which oracle/build_ref.py compiles from the
unmodified sources under /root/reference in the authoring container):

    gpurun -- python tests/golden/make_golden_refkernel.py --out gpurun_out/golden
    cp gpurun_out/golden/refkernel_outputs.npz tests/golden/

For every seeded case of tests/helpers.PARITY_CASES it stores the fp16 output of the reference's
`paged_attention_v1` plus, for a reshape_and_cache case, a checksum of the caches the reference's
`cache_ops.reshape_and_cache` produced.  Inputs are regenerated from the seeds by the tests (an
input digest is stored to catch generator drift).  CPU tests then pin the oracle to these numbers
without needing a GPU or /root/reference.
"""
import argparse
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import PARITY_CASES, build_case, load_ref_extension, run_ref_attention, to_cuda  # noqa: E402


def digest(case) -> str:
    h = hashlib.sha256()
    for k in ("q", "kc", "vc", "bt", "sl"):
        h.update(np.ascontiguousarray(case[k]).tobytes())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden"))
    a = ap.parse_args()
    ref = load_ref_extension()
    assert ref is not None, "oracle/_ref is missing: run oracle/build_ref.py where /root/reference exists"
    os.makedirs(a.out, exist_ok=True)
    store = {}
    for c in PARITY_CASES:
        name = c[0]
        case = build_case(name)
        out = run_ref_attention(ref, case, to_cuda(case))
        store[name + "/out"] = out
        store[name + "/digest"] = np.frombuffer(bytes.fromhex(digest(case)), dtype=np.uint8)
        print(name, out.shape, float(np.abs(out.astype(np.float32)).max()))
    # reshape_and_cache through the reference kernel: strided rows, a padding slot
    rng = np.random.default_rng(77)
    T, H, D, B, NB = 37, 12, 64, 16, 6
    qkv = rng.standard_normal((T, 3 * H * D)).astype(np.float16)
    slots = rng.permutation(NB * B)[:T].astype(np.int64)
    slots[5] = -1
    dev = torch.device("cuda:0")
    qkv_t = torch.from_numpy(qkv).to(dev)
    k_t = qkv_t[:, H * D:2 * H * D].view(T, H, D)
    v_t = qkv_t[:, 2 * H * D:].view(T, H, D)
    kc = torch.zeros(NB, H, D // 8, B, 8, dtype=torch.float16, device=dev)
    vc = torch.zeros(NB, H, D, B, dtype=torch.float16, device=dev)
    ref.cache_ops.reshape_and_cache(k_t, v_t, kc, vc, torch.from_numpy(slots).to(dev), "auto", 1.0)
    torch.cuda.synchronize()
    store["reshape/key_cache"] = kc.cpu().numpy()
    store["reshape/value_cache"] = vc.cpu().numpy()
    path = os.path.join(a.out, "refkernel_outputs.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
