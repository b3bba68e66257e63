"""Torch-eager CPU attention over paged K/V -- port of the reference's own CPU path.

TEST INFRASTRUCTURE ONLY (checker / reported CPU baseline; never imported by
vllmini_b200/).

The reference has exactly one CPU-runnable attention: `GPT2Attention.
_vanilla_attention` (vllmini/model/gpt2.py:71-78):

    attn_weights = torch.matmul(q, k.transpose(-1, -2)) * self.scale
    attn_weights = attn_weights + attention_mask        # if given
    attn_weights = softmax(attn_weights, dim=-1)
    return torch.matmul(attn_weights, v)

BASELINE.json's north_star asks for "a torch-eager CPU attention over the same
(q, key_cache, value_cache, block_tables, seq_lens) inputs ... timed on the
box's host cores".  This module gathers K/V through the block tables into
dense [S, H, Lmax, D] tensors (padding masked with -inf, which is what the
reference's additive `attention_mask` does in prefill, gpt2.py:73-74) and then
applies the four lines above verbatim in fp32 (CPU fp16 matmul is emulated and
would understate the baseline).
"""
from __future__ import annotations

import torch


def gather_dense(key_cache: torch.Tensor, value_cache: torch.Tensor, block_tables: torch.Tensor,
                 seq_lens: torch.Tensor):
    """K, V as dense fp32 [S, H, Lmax, D] + additive mask [S, 1, 1, Lmax].
    Address formulas: attention_kernels.cu:272-290 (K), :402-412 (V)."""
    NB, H, DX, B, x = key_cache.shape
    D = DX * x
    S = block_tables.shape[0]
    lmax = int(seq_lens.max().item()) if S else 0
    nblk = (lmax + B - 1) // B
    bt = block_tables[:, :nblk].to(torch.long).clamp_min(0)          # [S, nblk]
    k = key_cache[bt]                                                 # [S, nblk, H, D/x, B, x]
    k = k.permute(0, 2, 1, 4, 3, 5).reshape(S, H, nblk * B, D)[:, :, :lmax]
    v = value_cache[bt]                                               # [S, nblk, H, D, B]
    v = v.permute(0, 2, 1, 4, 3).reshape(S, H, nblk * B, D)[:, :, :lmax]
    pos = torch.arange(lmax)[None, :]
    mask = torch.where(pos < seq_lens.to(torch.long)[:, None], 0.0, float("-inf")).to(torch.float32)
    # NaN-safe: garbage in padded slots must not poison 0 * NaN in P @ V
    valid = (pos < seq_lens.to(torch.long)[:, None])[:, None, :, None]
    v = torch.where(valid, v.float(), torch.zeros((), dtype=torch.float32))
    k = torch.where(valid, k.float(), torch.zeros((), dtype=torch.float32))
    return k, v, mask[:, None, None, :]


def vanilla_attention(q, k, v, attention_mask, scale):
    """gpt2.py:71-78, verbatim semantics."""
    attn_weights = torch.matmul(q, k.transpose(-1, -2)) * scale
    if attention_mask is not None:
        attn_weights = attn_weights + attention_mask
    attn_weights = torch.nn.functional.softmax(attn_weights, dim=-1)
    return torch.matmul(attn_weights, v)


def paged_attention_eager_cpu(query, key_cache, value_cache, block_tables, seq_lens, scale):
    """query [S, H, D] (any float dtype, CPU) -> fp32 [S, H, D].
    Sequences with seq_len == 0 produce zeros (the paged kernel visits no block)."""
    k, v, mask = gather_dense(key_cache, value_cache, block_tables, seq_lens)
    q = query.float()[:, :, None, :]                                  # [S, H, 1, D]
    out = vanilla_attention(q, k, v, mask, scale)[:, :, 0, :]
    empty = (seq_lens <= 0)[:, None, None]
    return torch.where(empty, torch.zeros((), dtype=out.dtype), out)
