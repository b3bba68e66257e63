#!/usr/bin/env python
"""Generate tests/golden/vanilla_*.npz by importing the reference's own Python attention.

Run in the authoring container (needs /root/reference; CPU only):

    python tests/golden/make_golden_vanilla.py

It imports vllmini.model.gpt2 from /root/reference (unmodified; its `from paged_attention_cuda
import ...` resolves to this repo's drop-in, which is never *called* here) and evaluates
`GPT2Attention._vanilla_attention` (gpt2.py:71-78) in fp32 on seeded fp16 inputs.  The fixtures
hold dense q/K/V plus the reference's output; tests scatter K/V into paged caches with the oracle
and check that the oracle's paged attention reproduces these numbers, which pins the oracle to the
reference's CPU-runnable attention.  /root/reference is NOT needed to run the tests.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import vllmini_b200  # noqa: E402

vllmini_b200.install_dropin()
from transformers import GPT2Config  # noqa: E402
from vllmini.model.gpt2 import GPT2Attention  # noqa: E402  (the reference, unmodified)

CASES = [
    # name, H, D, seq_lens, seed
    ("vanilla_h12_d64", 12, 64, (3, 19), 101),        # 3 = the reference kernel test's ctx
    ("vanilla_h4_d64_ragged", 4, 64, (1, 16, 45), 102),
    ("vanilla_h2_d128", 2, 128, (33, 64), 103),
]


def main():
    torch.manual_seed(0)
    for name, H, D, lens, seed in CASES:
        cfg = GPT2Config(n_embd=H * D, n_head=H, n_layer=1)
        attn = GPT2Attention(cfg)
        assert abs(attn.scale - D ** -0.5) < 1e-12
        g = torch.Generator().manual_seed(seed)
        S, Lmax = len(lens), max(lens)
        q = torch.randn(S, H, D, generator=g).to(torch.float16)
        k = torch.zeros(S, Lmax, H, D, dtype=torch.float16)
        v = torch.zeros(S, Lmax, H, D, dtype=torch.float16)
        out = torch.zeros(S, H, D, dtype=torch.float32)
        for s, L in enumerate(lens):
            k[s, :L] = torch.randn(L, H, D, generator=g).to(torch.float16)
            v[s, :L] = torch.randn(L, H, D, generator=g).to(torch.float16)
            qs = q[s].float()[None, :, None, :]                       # [1, H, 1, D]
            ks = k[s, :L].float().permute(1, 0, 2)[None]              # [1, H, L, D]
            vs = v[s, :L].float().permute(1, 0, 2)[None]
            with torch.no_grad():
                out[s] = attn._vanilla_attention(qs, ks, vs, None)[0, :, 0, :]
        path = os.path.join(HERE, name + ".npz")
        np.savez(path, q=q.numpy(), k=k.numpy(), v=v.numpy(), seq_lens=np.asarray(lens, dtype=np.int32),
                 out=out.numpy(), scale=np.float64(attn.scale))
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
