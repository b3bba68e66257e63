"""Host-side mirror of the reference's operator interface for the paged-attention decode path.

Same names, positional argument order, in-place semantics and error behaviour as the reference's
pybind11 module `paged_attention_cuda` (paged_attention_ext/paged_attention_cuda/
paged_attention_cuda.cpp:7-25 and :51-61, cache_kernels.h:5-22), implemented as a thin layer over
the C ABI in include/vllmini_b200.h: this file only extracts pointers, sizes, strides and the
current CUDA stream from torch tensors.  PyTorch is plumbing here (device memory + streams).

Like the reference, every call is asynchronous on torch's current stream of the tensors' device,
allocates nothing (there is no workspace: split-KV merges inside a thread-block cluster), keeps no
state, and returns None.
"""
from __future__ import annotations

import os as _os
from typing import List, Optional

import torch

from . import _native

_ERR_NAMES = {-1: "CUDA error", -2: "Unsupported block size", -3: "Unsupported head size",
              -4: "invalid argument", -5: "unsupported"}

def _check(rc: int) -> None:
    if rc != 0:
        # reference: TORCH_CHECK -> c10::Error -> Python RuntimeError
        raise RuntimeError(f"vllmini_b200: {_native.last_error() or _ERR_NAMES.get(rc, 'error')} (code {rc})")


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        # reference: "CUDAGuardImpl initialized with non-CUDA DeviceType: cpu" 
        raise RuntimeError(f"vllmini_b200: {name} must be a CUDA tensor, got device {t.device}")


def paged_attention_v1(
    out: torch.Tensor,
    query: torch.Tensor,
    key_cache: torch.Tensor,
    value_cache: torch.Tensor,
    num_kv_heads: int,
    scale: float,
    block_tables: torch.Tensor,
    seq_lens: torch.Tensor,
    block_size: int,
    max_seq_len: int,
    alibi_slopes: Optional[torch.Tensor],
    kv_cache_dtype: str,
    kv_scale: float,
    tp_rank: int = 0,
    blocksparse_local_blocks: int = 0,
    blocksparse_vert_stride: int = 1,
    blocksparse_block_size: int = 1,
    blocksparse_head_sliding_step: int = 0,
) -> None:
    """Decode attention over the paged KV cache; writes `out` in place.

    Reference: attention_kernels.cu:805-826 (entry), :693-767 (launcher).  Shapes:
    out/query [num_seqs, num_heads, head_size] (query may have any row stride, out is written as
    contiguous [num_seqs, num_heads, head_size] whatever its nominal rank -- the reference does not
    check out's rank either, tests/kernels/paged_attention.py:114); key_cache
    [NB, KVH, D/8, B, 8]; value_cache [NB, KVH, D, B]; block_tables int32 [S, MB]; seq_lens int32 [S].
    """
    _require_cuda(query, "query")
    if query.dtype is not torch.float16:
        # reference dispatch also has float/bfloat16 arms (quant_utils.cuh:529-566); vllmini is
        # fp16-only (kv_cache.py:13-14, scheduler.py:13) and so is this build
        raise RuntimeError(f"vllmini_b200: unsupported input type of kv cache: {query.dtype}")
    if kv_cache_dtype != "auto":
        raise RuntimeError(f"vllmini_b200: Unsupported data type of kv cache: {kv_cache_dtype}")
    if alibi_slopes is not None:
        raise RuntimeError("vllmini_b200: alibi_slopes is not supported (vllmini passes None, gpt2.py:105)")
    if blocksparse_vert_stride > 1:
        raise RuntimeError("vllmini_b200: block-sparse attention is not supported "
                           "(vllmini passes blocksparse_vert_stride=1, gpt2.py:110)")
    for t, name in ((out, "out"), (key_cache, "key_cache"), (value_cache, "value_cache"),
                    (block_tables, "block_tables"), (seq_lens, "seq_lens")):
        _require_cuda(t, name)
    if block_tables.dtype is not torch.int32:
        raise RuntimeError(f"vllmini_b200: expected scalar type Int but found {block_tables.dtype} (block_tables)")
    if seq_lens.dtype is not torch.int32:
        raise RuntimeError(f"vllmini_b200: expected scalar type Int but found {seq_lens.dtype} (seq_lens)")
    if key_cache.dtype is not torch.float16 or value_cache.dtype is not torch.float16 or out.dtype is not torch.float16:
        raise RuntimeError("vllmini_b200: out, key_cache and value_cache must be float16")
    if query.dim() != 3 or query.stride(2) != 1 or query.stride(1) != query.size(2):
        raise RuntimeError("vllmini_b200: query must be [num_seqs, num_heads, head_size] with contiguous heads")
    if not out.is_contiguous() or not block_tables.is_contiguous() or not seq_lens.is_contiguous():
        raise RuntimeError("vllmini_b200: out, block_tables and seq_lens must be contiguous")
    num_seqs, num_heads, head_size = query.shape
    if out.numel() != query.numel():
        raise RuntimeError("vllmini_b200: out must have num_seqs*num_heads*head_size elements")
    max_blocks = block_tables.size(1) if block_tables.dim() == 2 else 0
    dev = query.device
    dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
    lib = _native.load()
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = lib.vb200_paged_attention_v1_f16(
        out.data_ptr(), query.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr(),
        block_tables.data_ptr(), seq_lens.data_ptr(),
        num_seqs, num_heads, int(num_kv_heads), head_size, int(block_size), max_blocks, int(max_seq_len),
        query.stride(0), key_cache.stride(0), key_cache.stride(1), float(scale), dev_index, stream)
    _check(rc)


def reshape_and_cache(
    key: torch.Tensor,
    value: torch.Tensor,
    key_cache: torch.Tensor,
    value_cache: torch.Tensor,
    slot_mapping: torch.Tensor,
    kv_cache_dtype: str,
    kv_scale: float,
) -> None:
    """Scatter key/value rows [T, H, D] into the paged caches at slot_mapping (int64), in place.
    Reference: cache_kernels.cu:256-281 (entry), :152-207 (kernel)."""
    _require_cuda(key, "key")
    if kv_cache_dtype != "auto":
        raise RuntimeError(f"vllmini_b200: Unsupported data type of kv cache: {kv_cache_dtype}")
    if key.dtype is not torch.float16:
        raise RuntimeError(f"vllmini_b200: unsupported input type of kv cache: {key.dtype}")
    for t, name in ((value, "value"), (key_cache, "key_cache"), (value_cache, "value_cache"),
                    (slot_mapping, "slot_mapping")):
        _require_cuda(t, name)
    if slot_mapping.dtype is not torch.int64:
        raise RuntimeError(f"vllmini_b200: expected scalar type Long but found {slot_mapping.dtype} (slot_mapping)")
    if value.dtype is not torch.float16 or key_cache.dtype is not torch.float16 or value_cache.dtype is not torch.float16:
        raise RuntimeError("vllmini_b200: value, key_cache and value_cache must be float16")
    if key.dim() != 3 or key.stride(2) != 1 or key.stride(1) != key.size(2) \
            or value.stride(2) != 1 or value.stride(1) != value.size(2):
        raise RuntimeError("vllmini_b200: key/value must be [num_tokens, num_heads, head_size] with contiguous heads")
    if key_cache.dim() != 5 or not key_cache.is_contiguous() or not value_cache.is_contiguous():
        raise RuntimeError("vllmini_b200: key_cache must be contiguous [NB, H, D/x, B, x]; value_cache [NB, H, D, B]")
    if not slot_mapping.is_contiguous():
        raise RuntimeError("vllmini_b200: slot_mapping must be contiguous")
    num_tokens, num_heads, head_size = key.shape
    block_size, x = key_cache.size(3), key_cache.size(4)
    dev = key.device
    dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
    lib = _native.load()
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = lib.vb200_reshape_and_cache_f16(
        key.data_ptr(), value.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr(), slot_mapping.data_ptr(),
        num_tokens, num_heads, head_size, block_size, x, key.stride(0), value.stride(0), dev_index, stream)
    _check(rc)


def paged_attention_v1_append(
    out: torch.Tensor,
    query: torch.Tensor,
    key: torch.Tensor,
    value: torch.Tensor,
    key_cache: torch.Tensor,
    value_cache: torch.Tensor,
    slot_mapping: torch.Tensor,
    num_kv_heads: int,
    scale: float,
    block_tables: torch.Tensor,
    seq_lens: torch.Tensor,
    block_size: int,
    max_seq_len: int,
) -> None:
    """Extension (no reference counterpart): one launch that does exactly
    `reshape_and_cache(key, value, key_cache, value_cache, slot_mapping, "auto", 1.0)` followed by
    `paged_attention_v1(out, query, ...)` for a decode step -- key/value [num_seqs, num_kv_heads, head_size]
    hold ONE new token per sequence, located at position >= seq_lens[s] - 1 of its sequence."""
    for t, name in ((out, "out"), (query, "query"), (key, "key"), (value, "value"), (key_cache, "key_cache"),
                    (value_cache, "value_cache"), (slot_mapping, "slot_mapping"), (block_tables, "block_tables"),
                    (seq_lens, "seq_lens")):
        _require_cuda(t, name)
    if any(t.dtype is not torch.float16 for t in (out, query, key, value, key_cache, value_cache)):
        raise RuntimeError("vllmini_b200: out, query, key, value and the caches must be float16")
    if block_tables.dtype is not torch.int32 or seq_lens.dtype is not torch.int32 or slot_mapping.dtype is not torch.int64:
        raise RuntimeError("vllmini_b200: block_tables/seq_lens must be int32 and slot_mapping int64")
    num_seqs, num_heads, head_size = query.shape
    if tuple(key.shape) != (num_seqs, int(num_kv_heads), head_size) or tuple(value.shape) != tuple(key.shape):
        raise RuntimeError("vllmini_b200: key/value must be [num_seqs, num_kv_heads, head_size] (one token per sequence)")
    for t in (query, key, value):
        if t.stride(2) != 1 or t.stride(1) != head_size:
            raise RuntimeError("vllmini_b200: query/key/value must have contiguous heads")
    if not (out.is_contiguous() and block_tables.is_contiguous() and seq_lens.is_contiguous()
            and slot_mapping.is_contiguous() and key_cache.is_contiguous() and value_cache.is_contiguous()):
        raise RuntimeError("vllmini_b200: out, block_tables, seq_lens, slot_mapping and the caches must be contiguous")
    dev = query.device
    dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
    lib = _native.load()
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = lib.vb200_paged_attention_v1_append_f16(
        out.data_ptr(), query.data_ptr(), key.data_ptr(), value.data_ptr(), key_cache.data_ptr(),
        value_cache.data_ptr(), slot_mapping.data_ptr(), block_tables.data_ptr(), seq_lens.data_ptr(),
        num_seqs, num_heads, int(num_kv_heads), head_size, int(block_size),
        block_tables.size(1) if block_tables.dim() == 2 else 0, int(max_seq_len),
        query.stride(0), key.stride(0), value.stride(0), key_cache.stride(0), key_cache.stride(1), float(scale),
        dev_index, stream)
    _check(rc)


def add_layernorm(a: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
                  x_out: torch.Tensor, ln_out: torch.Tensor, a_rows: Optional[torch.Tensor] = None,
                  x_rows: Optional[torch.Tensor] = None) -> None:
    """Decode-driver helper (no reference counterpart; vllmini does this with eager torch ops, gpt2.py:150-160):
    x_out = fp16(a[a_rows] + x[x_rows]); ln_out = LayerNorm(x_out) * gamma + beta, one launch.  `x_out` may be `x`
    itself (in place) when x_rows is None.  a, x: [*, hidden] fp16 with contiguous rows; x_out, ln_out
    [num_rows, hidden] contiguous; a_rows / x_rows: optional int64 row indices (the embedding lookup)."""
    for t, name in ((a, "a"), (x, "x"), (gamma, "gamma"), (beta, "beta"), (x_out, "x_out"), (ln_out, "ln_out")):
        _require_cuda(t, name)
        if t.dtype is not torch.float16:
            raise RuntimeError(f"vllmini_b200: {name} must be float16")
    hidden = x_out.shape[-1]
    num_rows = x_out.numel() // hidden
    if a.dim() != 2 or x.dim() != 2 or a.stride(1) != 1 or x.stride(1) != 1 or a.shape[1] != hidden or x.shape[1] != hidden:
        raise RuntimeError("vllmini_b200: a and x must be [rows, hidden] with contiguous rows")
    if not (x_out.is_contiguous() and ln_out.is_contiguous() and gamma.is_contiguous() and beta.is_contiguous()) \
            or ln_out.numel() != x_out.numel() or gamma.numel() != hidden or beta.numel() != hidden:
        raise RuntimeError("vllmini_b200: x_out/ln_out must be contiguous [num_rows, hidden], gamma/beta [hidden]")
    for rows, src, name in ((a_rows, a, "a_rows"), (x_rows, x, "x_rows")):
        if rows is None:
            if src.shape[0] < num_rows:
                raise RuntimeError(f"vllmini_b200: fewer source rows than num_rows and no {name}")
        else:
            _require_cuda(rows, name)
            if rows.dtype is not torch.int64 or not rows.is_contiguous() or rows.numel() != num_rows:
                raise RuntimeError(f"vllmini_b200: {name} must be a contiguous int64 tensor of num_rows entries")
    dev = x_out.device
    dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
    rc = _native.load().vb200_add_layernorm_f16(
        x_out.data_ptr(), ln_out.data_ptr(), a.data_ptr(), a_rows.data_ptr() if a_rows is not None else None,
        a.shape[0], a.stride(0), x.data_ptr(), x_rows.data_ptr() if x_rows is not None else None, x.shape[0],
        x.stride(0),
        gamma.data_ptr(), beta.data_ptr(), num_rows, hidden, float(eps), dev_index,
        torch.cuda.current_stream(dev).cuda_stream)
    _check(rc)


def argmax_rows(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Decode-driver helper: torch.argmax(x, dim=-1) for fp16 logits [rows, n] (rows may be strided), one launch."""
    _require_cuda(x, "x")
    if x.dtype is not torch.float16 or x.dim() != 2 or x.stride(1) != 1:
        raise RuntimeError("vllmini_b200: x must be a float16 [rows, n] tensor with contiguous rows")
    if out is None:
        out = torch.empty(x.shape[0], dtype=torch.int64, device=x.device)
    elif out.dtype is not torch.int64 or not out.is_contiguous() or out.numel() != x.shape[0] or out.device != x.device:
        raise RuntimeError("vllmini_b200: out must be a contiguous int64 tensor of `rows` entries on x's device")
    dev = x.device
    dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
    _check(_native.load().vb200_argmax_rows_f16(out.data_ptr(), x.data_ptr(), x.stride(0), x.shape[0], x.shape[1],
                                                dev_index, torch.cuda.current_stream(dev).cuda_stream))
    return out


# ---- binding selection -------------------------------------------------------------------------------
# The two functions above bind the C ABI through ctypes (no build dependency on torch).  When the compiled
# torch shim (csrc/torch_shim.cpp -> _torch_shim*.so, `python -m vllmini_b200.build`) is present it replaces
# them: same validation, same C ABI calls, same kernels, ~5x less Python overhead per call (the reference's
# own binding is a pybind11 module too).  VLLMINI_B200_NO_SHIM=1 keeps the ctypes path.
paged_attention_v1_ctypes = paged_attention_v1
reshape_and_cache_ctypes = reshape_and_cache
BINDING = "ctypes"
if not _os.environ.get("VLLMINI_B200_NO_SHIM"):
    try:
        _native.load()                       # fail loudly first if the CUDA library itself is missing
        from . import _torch_shim as _shim   # noqa: E402
        paged_attention_v1 = _shim.paged_attention_v1
        reshape_and_cache = _shim.reshape_and_cache
        BINDING = "torch_shim"
    except ImportError:
        pass


# ---- the rest of the reference's exported surface -------------------------------------------------
# Present so that the drop-in covers what the reference package exports (paged_attention_cuda.cpp:51-61).
# vllmini calls none of these; they are control-plane
# operations, so they stay on the ctypes binding.
def paged_attention_v2(
    out: torch.Tensor,
    exp_sums: torch.Tensor,
    max_logits: torch.Tensor,
    tmp_out: torch.Tensor,
    query: torch.Tensor,
    key_cache: torch.Tensor,
    value_cache: torch.Tensor,
    num_kv_heads: int,
    scale: float,
    block_tables: torch.Tensor,
    seq_lens: torch.Tensor,
    block_size: int,
    max_seq_len: int,
    alibi_slopes: Optional[torch.Tensor],
    kv_cache_dtype: str,
    kv_scale: float,
    tp_rank: int = 0,
    blocksparse_local_blocks: int = 0,
    blocksparse_vert_stride: int = 1,
    blocksparse_block_size: int = 1,
    blocksparse_head_sliding_step: int = 0,
) -> None:
    """The reference's partitioned entry (attention_kernels.cu:966-990; launcher :844-931).  The context is cut
    into the reference's fixed 512-token partitions; every partition that holds a token of its sequence leaves its
    record -- max logit, sum of exp(logit - max), normalised partial output -- in the CALLER's scratch tensors
    exp_sums / max_logits [S, H, P] (fp32) and tmp_out [S, H, P, D] (fp16), P = ceil(max_seq_len / 512), exactly
    where the reference's partition kernel puts them (:349-357, :484-495), and a second launch reduces the records
    into `out` (:564-669).  No hidden workspace is involved."""
    _require_cuda(query, "query")
    if query.dtype is not torch.float16:
        raise RuntimeError(f"vllmini_b200: unsupported input type of kv cache: {query.dtype}")
    if kv_cache_dtype != "auto":
        raise RuntimeError(f"vllmini_b200: Unsupported data type of kv cache: {kv_cache_dtype}")
    if alibi_slopes is not None:
        raise RuntimeError("vllmini_b200: alibi_slopes is not supported (vllmini passes None, gpt2.py:105)")
    if blocksparse_vert_stride > 1:
        raise RuntimeError("vllmini_b200: block-sparse attention is not supported")
    for t, name in ((out, "out"), (exp_sums, "exp_sums"), (max_logits, "max_logits"), (tmp_out, "tmp_out"),
                    (key_cache, "key_cache"), (value_cache, "value_cache"), (block_tables, "block_tables"),
                    (seq_lens, "seq_lens")):
        _require_cuda(t, name)
    if block_tables.dtype is not torch.int32:
        raise RuntimeError(f"vllmini_b200: expected scalar type Int but found {block_tables.dtype} (block_tables)")
    if seq_lens.dtype is not torch.int32:
        raise RuntimeError(f"vllmini_b200: expected scalar type Int but found {seq_lens.dtype} (seq_lens)")
    if any(t.dtype is not torch.float16 for t in (out, tmp_out, key_cache, value_cache)):
        raise RuntimeError("vllmini_b200: out, tmp_out, key_cache and value_cache must be float16")
    if query.dim() != 3 or query.stride(2) != 1 or query.stride(1) != query.size(2):
        raise RuntimeError("vllmini_b200: query must be [num_seqs, num_heads, head_size] with contiguous heads")
    S, H, D = query.shape
    if int(max_seq_len) <= 0:
        raise RuntimeError("vllmini_b200: max_seq_len must be positive")
    P = (int(max_seq_len) + 511) // 512
    if tuple(exp_sums.shape) != (S, H, P) or tuple(max_logits.shape) != (S, H, P):
        raise RuntimeError(f"vllmini_b200: exp_sums/max_logits must be [{S}, {H}, {P}]")
    if tuple(tmp_out.shape) != (S, H, P, D):
        raise RuntimeError(f"vllmini_b200: tmp_out must be [{S}, {H}, {P}, {D}]")
    if exp_sums.dtype is not torch.float32 or max_logits.dtype is not torch.float32:
        raise RuntimeError("vllmini_b200: exp_sums and max_logits must be float32")
    if not (out.is_contiguous() and exp_sums.is_contiguous() and max_logits.is_contiguous() and tmp_out.is_contiguous()
            and block_tables.is_contiguous() and seq_lens.is_contiguous()):
        raise RuntimeError("vllmini_b200: out, the scratch tensors, block_tables and seq_lens must be contiguous")
    if out.numel() != query.numel():
        raise RuntimeError("vllmini_b200: out must have num_seqs*num_heads*head_size elements")
    dev = query.device
    dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
    rc = _native.load().vb200_paged_attention_v2_f16(
        out.data_ptr(), exp_sums.data_ptr(), max_logits.data_ptr(), tmp_out.data_ptr(), query.data_ptr(),
        key_cache.data_ptr(), value_cache.data_ptr(), block_tables.data_ptr(), seq_lens.data_ptr(),
        S, H, int(num_kv_heads), D, int(block_size), block_tables.size(1) if block_tables.dim() == 2 else 0,
        int(max_seq_len), query.stride(0), key_cache.stride(0), key_cache.stride(1), float(scale), dev_index,
        torch.cuda.current_stream(dev).cuda_stream)
    _check(rc)


def swap_blocks(src: torch.Tensor, dst: torch.Tensor, block_mapping: torch.Tensor) -> None:
    """Copy whole cache blocks between two cache tensors (GPU<->GPU, GPU->CPU, CPU->GPU) for every
    (src_block, dst_block) row of the CPU int64 `block_mapping`.  Reference: cache_kernels.cu:24-63."""
    if src.is_cuda and dst.is_cuda:
        if src.device != dst.device:
            raise RuntimeError("vllmini_b200: src and dst must be on the same GPU")
        kind, dev = 0, src.device
    elif src.is_cuda and not dst.is_cuda:
        kind, dev = 1, src.device
    elif not src.is_cuda and dst.is_cuda:
        kind, dev = 2, dst.device
    else:
        raise RuntimeError("vllmini_b200: Invalid device combination")
    if block_mapping.is_cuda:
        raise RuntimeError("vllmini_b200: block_mapping must be on CPU")
    if not src.is_contiguous() or not dst.is_contiguous():
        raise RuntimeError("vllmini_b200: src and dst must be contiguous")
    bm = block_mapping.to(torch.int64).contiguous()
    block_bytes = src.element_size() * src[0].numel()
    dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
    lib = _native.load()
    rc = lib.vb200_swap_blocks(src.data_ptr(), dst.data_ptr(), bm.data_ptr(), bm.size(0), block_bytes, kind,
                               dev_index, torch.cuda.current_stream(dev).cuda_stream)
    _check(rc)


def copy_blocks(key_caches: List[torch.Tensor], value_caches: List[torch.Tensor],
                block_mapping: torch.Tensor) -> None:
    """For every layer, copy block src over block dst in that layer's key and value cache, for every
    (src, dst) row of the CUDA int64 `block_mapping` [num_pairs, 2].  Reference: cache_kernels.cu:67-148."""
    num_layers = len(key_caches)
    if num_layers != len(value_caches):
        raise RuntimeError("vllmini_b200: key_caches and value_caches must have the same length")
    if num_layers == 0:
        return
    dev = key_caches[0].device
    if not dev.type == "cuda":
        raise RuntimeError("vllmini_b200: caches must be CUDA tensors")
    if not block_mapping.is_cuda or block_mapping.dtype is not torch.int64:
        raise RuntimeError("vllmini_b200: block_mapping must be a CUDA int64 tensor")
    bm = block_mapping.contiguous()
    ptrs = [t.data_ptr() for t in key_caches] + [t.data_ptr() for t in value_caches]
    # like the reference, the pointer table is staged through a small H2D copy (cache_kernels.cu:121-130)
    table = torch.tensor(ptrs, dtype=torch.int64).to(dev, non_blocking=False)
    block_bytes = key_caches[0].element_size() * key_caches[0][0].numel()
    aligned = int(all(p % 16 == 0 for p in ptrs))
    dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
    lib = _native.load()
    rc = lib.vb200_copy_blocks(table.data_ptr(), table.data_ptr() + 8 * num_layers, bm.data_ptr(), num_layers,
                               bm.size(0), block_bytes, aligned, dev_index,
                               torch.cuda.current_stream(dev).cuda_stream)
    _check(rc)


def reshape_and_cache_flash(key: torch.Tensor, value: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor,
                            slot_mapping: torch.Tensor, kv_cache_dtype: str) -> None:
    """Scatter key/value rows [T, H, D] into caches laid out [num_blocks, block_size, H, D].
    Reference: cache_kernels.cu:283-317 (entry), :209-240 (kernel)."""
    _require_cuda(key, "key")
    if kv_cache_dtype != "auto":
        raise RuntimeError(f"vllmini_b200: Unsupported data type of kv cache: {kv_cache_dtype}")
    for t, name in ((value, "value"), (k_cache, "k_cache"), (v_cache, "v_cache"), (slot_mapping, "slot_mapping")):
        _require_cuda(t, name)
    if key.dtype is not torch.float16 or value.dtype is not torch.float16 or k_cache.dtype is not torch.float16 \
            or v_cache.dtype is not torch.float16:
        raise RuntimeError("vllmini_b200: key, value and caches must be float16")
    if slot_mapping.dtype is not torch.int64:
        raise RuntimeError(f"vllmini_b200: expected scalar type Long but found {slot_mapping.dtype} (slot_mapping)")
    if k_cache.stride(0) != v_cache.stride(0):
        raise RuntimeError("vllmini_b200: k_cache.stride(0) must equal v_cache.stride(0)")   # cache_kernels.cu:300
    T, H, D = key.shape
    if key.stride(2) != 1 or key.stride(1) != D or value.stride(2) != 1 or value.stride(1) != D:
        raise RuntimeError("vllmini_b200: key/value must have contiguous heads")
    if tuple(k_cache.shape[2:]) != (H, D) or k_cache.stride(3) != 1 or k_cache.stride(2) != D \
            or k_cache.stride(1) != H * D:
        raise RuntimeError("vllmini_b200: k_cache must be [num_blocks, block_size, num_heads, head_size]")
    dev = key.device
    dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
    lib = _native.load()
    rc = lib.vb200_reshape_and_cache_flash_f16(key.data_ptr(), value.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(),
                                               slot_mapping.contiguous().data_ptr(), T, H, D, k_cache.size(1),
                                               k_cache.stride(0), key.stride(0), value.stride(0), dev_index,
                                               torch.cuda.current_stream(dev).cuda_stream)
    _check(rc)


def convert_fp8(*args, **kwargs) -> None:
    raise NotImplementedError("cache_ops.convert_fp8 (cache_kernels.cu:334-392) is test-only and fp8 is disabled "
                              "in the reference build (no ENABLE_FP8: its kernel body is assert(false))")
