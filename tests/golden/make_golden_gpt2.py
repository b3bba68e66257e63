#!/usr/bin/env python
"""Generate tests/golden/gpt2_greedy_*.npz with the REFERENCE'S OWN GPT-2 model on CPU.

Run in the authoring container (needs /root/reference; CPU only, ~1 minute):

    python tests/golden/make_golden_gpt2.py

BASELINE.json configs[0]: "GPT-2 small, 1 prompt, 8-token greedy decode via torch eager attention on
CPU (correctness reference, no paged kernel)".  This script imports vllmini.model.gpt2.
GPT2LMHeadModel from /root/reference unmodified, loads the seeded synthetic weights of
vllmini_b200.decode.synthetic_gpt2_state_dict (no network for the HF checkpoint), and decodes
greedily by re-running the model's eager (prefill) attention path over the growing sequence -- the
paged kernel is never involved.  The model's unconditional `cache_ops.reshape_and_cache` call
(gpt2.py:44) is pointed at the numpy oracle so it can run on CPU tensors.

The fixture stores the prompt, the 8 greedy tokens, and for every step the top-64 logits (values,
ids) plus the full-row mean/std, so tests can compare the GPU paged path (fp16) against it with a
tolerance and check token equality wherever the top-2 margin exceeds that tolerance.
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
import paged_attention_cuda  # noqa: E402
from oracle import paged_oracle as po  # noqa: E402
from transformers import GPT2Config  # noqa: E402
from vllmini_b200.decode import GPT2Shape, synthetic_gpt2_state_dict  # noqa: E402


def _cpu_reshape_and_cache(key, value, key_cache, value_cache, slot_mapping, kv_cache_dtype, kv_scale):
    po.reshape_and_cache(key.detach().numpy(), value.detach().numpy(), key_cache.numpy(), value_cache.numpy(),
                         slot_mapping.numpy())


paged_attention_cuda.cache_ops.reshape_and_cache = _cpu_reshape_and_cache
from vllmini.model.gpt2 import GPT2LMHeadModel  # noqa: E402  (the reference, unmodified)

CASES = [
    ("gpt2_greedy_tiny", GPT2Shape(n_layer=2, n_head=4, n_embd=256, vocab_size=503, n_positions=128), 11, 13, 8),
    ("gpt2_greedy_small", GPT2Shape(), 12, 16, 8),          # GPT-2 small: configs[0]
]


def main():
    for name, shape, seed, prompt_len, new_tokens in CASES:
        cfg = GPT2Config(n_layer=shape.n_layer, n_head=shape.n_head, n_embd=shape.n_embd, vocab_size=shape.vocab_size,
                         n_positions=shape.n_positions, layer_norm_epsilon=shape.layer_norm_epsilon)
        model = GPT2LMHeadModel(cfg).float().eval()
        sd = synthetic_gpt2_state_dict(shape, seed=seed)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        rng = np.random.default_rng(seed)
        prompt = rng.integers(0, shape.vocab_size, size=prompt_len).tolist()
        H, D, B = shape.n_head, shape.head_size, 16
        ids = list(prompt)
        toks, top_vals, top_ids, stats = [], [], [], []
        for step in range(new_tokens):
            T = len(ids)
            nblk = (T + B - 1) // B
            key_cache = torch.zeros(shape.n_layer * nblk, H, D // 8, B, 8)
            value_cache = torch.zeros(shape.n_layer * nblk, H, D, B)
            slot_mappings = [torch.arange(T, dtype=torch.long) + l * nblk * B for l in range(shape.n_layer)]
            mask = torch.triu(torch.full((T, T), float("-inf")), diagonal=1)[None, None].expand(1, H, T, T)
            with torch.no_grad():
                logits, _ = model(input_ids=torch.tensor([ids]), position_ids=torch.arange(T), attention_mask=mask,
                                  use_cache=True, is_prefill=True, key_cache=key_cache, value_cache=value_cache,
                                  slot_mappings=slot_mappings, block_tables=None)
            row = logits[0, -1].double()
            tv, ti = torch.topk(row, 64)
            toks.append(int(ti[0]))
            top_vals.append(tv.numpy())
            top_ids.append(ti.numpy())
            stats.append([float(row.mean()), float(row.std())])
            ids.append(int(ti[0]))
            if step == 0:   # the caches the reference model filled must hold the prompt's K rows (layout pin)
                assert float(key_cache.abs().sum()) > 0 and float(value_cache.abs().sum()) > 0
        path = os.path.join(HERE, name + ".npz")
        np.savez(path, prompt=np.asarray(prompt, dtype=np.int64), tokens=np.asarray(toks, dtype=np.int64),
                 top_vals=np.asarray(top_vals), top_ids=np.asarray(top_ids), stats=np.asarray(stats),
                 shape=np.asarray([shape.n_layer, shape.n_head, shape.n_embd, shape.vocab_size, shape.n_positions]),
                 seed=np.int64(seed))
        margins = [float(v[0] - v[1]) for v in top_vals]
        print("wrote", path, os.path.getsize(path), "bytes; tokens", toks, "top-2 margins", np.round(margins, 4))


if __name__ == "__main__":
    main()
