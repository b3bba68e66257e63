"""Drop-in replacement for the reference's `paged_attention_cuda` extension package.

vllmini imports exactly this surface (vllmini/model/gpt2.py:5, tests/kernels/paged_attention.py:4):

    from paged_attention_cuda import paged_attention_v1, cache_ops
    cache_ops.reshape_and_cache(k, v, key_cache, value_cache, slot_mapping, "auto", 1.0)
    paged_attention_v1(out, q, key_cache, value_cache, num_heads, scale, block_table, seq_lens,
                       block_size, max_seq_len, None, "auto", 1.0, 0, 0, 1, 1, 0)

and the reference package exports `__all__ = ['paged_attention_v1', 'paged_attention_v2', 'cache_ops']`
(paged_attention_ext/paged_attention_cuda/__init__.py:4-8).  Put this directory's parent on sys.path
(`vllmini_b200.install_dropin()` does it) or pip-install it in place of the reference extension and
vllmini's kv_cache / block_manager / scheduler / model run unmodified on the sm_100a kernels.
"""
import types as _types

import torch  # noqa: F401  (the reference extension also needs torch loaded first)

from vllmini_b200 import ops as _ops

paged_attention_v1 = _ops.paged_attention_v1
paged_attention_v2 = _ops.paged_attention_v2

# pybind submodule `cache_ops` (paged_attention_cuda.cpp:56-61)
cache_ops = _types.ModuleType(__name__ + ".cache_ops", "Cache operations")
cache_ops.swap_blocks = _ops.swap_blocks
cache_ops.copy_blocks = _ops.copy_blocks
cache_ops.reshape_and_cache = _ops.reshape_and_cache
cache_ops.reshape_and_cache_flash = _ops.reshape_and_cache_flash
cache_ops.convert_fp8 = _ops.convert_fp8

__all__ = ["paged_attention_v1", "paged_attention_v2", "cache_ops"]
