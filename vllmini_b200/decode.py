"""Batched decode driver for the paged-attention hot path.

The reference steps ONE sequence per forward pass with dozens of GPU->CPU syncs in its block
bookkeeping (vllmini/scheduler.py:55-108, block_manager.py:28-63, kv_cache.py:56-73), so its host code
never exercises the kernels at batch > 1.  This module is the harness-side driver that does: it keeps
the reference's data model -- one KV pool for all layers, FIFO free list, per-(sequence, layer) block
tables padded with -1, int64 slot mappings, the cache layouts of kv_cache.py:13-14 -- but advances all
sequences together, with the bookkeeping vectorised on the host (numpy) and no device->host sync on
the step path.

  BatchedKVState   batched counterpart of KVCache + BlockManager.decode_step
  GPT2Decoder      a GPT-2 host model with the reference GPT2LMHeadModel's parameter names
                   (state_dict-compatible with vllmini/model/gpt2.py:243-331): GEMMs / LayerNorm /
                   GELU are torch (cuBLAS) calls -- plumbing -- and the two hot ops go through
                   vllmini_b200.ops.  Optional CUDA-graph capture of the whole decode step.

Nothing here is required to use the drop-in with an unmodified vllmini; it exists so that the
BASELINE configs with 256+ concurrent sequences can be driven end to end.
"""
from __future__ import annotations

import collections
import math
import weakref
from dataclasses import dataclass
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

from . import ops


@dataclass
class GPT2Shape:
    """GPT-2 small by default (transformers.GPT2Config() defaults)."""
    n_layer: int = 12
    n_head: int = 12
    n_embd: int = 768
    vocab_size: int = 50257
    n_positions: int = 1024
    layer_norm_epsilon: float = 1e-5

    @property
    def head_size(self) -> int:
        return self.n_embd // self.n_head


# ---------------------------------------------------------------------------------------------
# KV bookkeeping
# ---------------------------------------------------------------------------------------------
class BatchedKVState:
    """Paged KV pool + block tables for up to `max_seqs` concurrent sequences, all layers.

    Layout contract is the reference's (kv_cache.py:13-14, 30-35): key_cache [NB, H, D/8, B, 8],
    value_cache [NB, H, D, B]; one pool shared by all layers, a block belongs to one (sequence,
    layer); table rows are int32 padded with -1; free list is FIFO (kv_cache.py:25-26, 60, 83).
    Unlike the reference's allocate_for_prefill (one block per layer whatever the prompt length,
    kv_cache.py:25 -- prompts longer than a block spill into the neighbour's block), prefill here
    allocates ceil(prompt_len / B) blocks per layer.
    """

    def __init__(self, num_blocks: int, num_layers: int, num_heads: int, head_size: int, block_size: int,
                 max_seqs: int, max_blocks_per_seq: int, device: torch.device, dtype=torch.float16):
        assert head_size % 8 == 0
        self.num_blocks, self.num_layers, self.num_heads = num_blocks, num_layers, num_heads
        self.head_size, self.block_size = head_size, block_size
        self.max_seqs, self.max_blocks_per_seq = max_seqs, max_blocks_per_seq
        self.device = device
        self.key_cache = torch.zeros(num_blocks, num_heads, head_size // 8, block_size, 8, dtype=dtype, device=device)
        self.value_cache = torch.zeros(num_blocks, num_heads, head_size, block_size, dtype=dtype, device=device)
        self.free_blocks = collections.deque(range(num_blocks))
        # host mirrors (authoritative) and device copies read by the kernels
        self.h_tables = np.full((num_layers, max_seqs, max_blocks_per_seq), -1, dtype=np.int32)
        self.h_lens = np.zeros(max_seqs, dtype=np.int32)
        self.active = np.zeros(max_seqs, dtype=bool)
        self.d_tables = torch.full((num_layers, max_seqs, max_blocks_per_seq), -1, dtype=torch.int32, device=device)
        # Per-step host->device staging: ONE packed buffer [slots int64 [L,S] | positions int64 [S] | flat table
        # indices of newly appended blocks int64 [L,S] | lens int32 [S] | new block ids int32 [L,S]] and one async
        # copy per step.  d_slots / d_pos / d_lens ARE views of the device copy (no unpacking kernels); only a
        # step that appends blocks launches one scatter into d_tables.  On the host a ring of pinned buffers, each
        # guarded by an event recorded after its copy, lets the host prepare steps ahead of the GPU: nothing on
        # the step path synchronises the host with the device.
        L, S = num_layers, max_seqs
        self._o_pos = L * S * 8
        self._o_grow = self._o_pos + S * 8
        self._o_lens = self._o_grow + L * S * 8
        self._o_blocks = self._o_lens + S * 4
        self._stage_bytes = self._o_blocks + L * S * 4
        self._ring = []
        for _ in range(4):
            buf = torch.zeros(self._stage_bytes, dtype=torch.uint8)
            if device.type == "cuda":
                buf = buf.pin_memory()
            self._ring.append([buf, torch.cuda.Event() if device.type == "cuda" else None, False])
        self._ring_i = 0
        ds = self._d_stage = torch.zeros(self._stage_bytes, dtype=torch.uint8, device=device)
        self.d_slots = ds[: self._o_pos].view(torch.int64).view(L, S)               # slot_mapping per layer
        self.d_slots.fill_(-1)
        self.d_pos = ds[self._o_pos: self._o_grow].view(torch.int64)                # position ids of the new tokens
        self._d_grow = ds[self._o_grow: self._o_lens].view(torch.int64)             # [L*S] flat indices into d_tables
        self.d_lens = ds[self._o_lens: self._o_blocks].view(torch.int32)            # seq_lens passed to attention
        self._d_blocks = ds[self._o_blocks:].view(torch.int32)                      # [L*S] new block ids
        # first use of the scatter kernel pays a one-off module load: do it now, not inside a decode step
        self.d_tables.view(-1).scatter_(0, torch.zeros(1, dtype=torch.int64, device=device),
                                        torch.full((1,), -1, dtype=torch.int32, device=device))

    # -- allocation ---------------------------------------------------------------------------
    def _take(self, n: int) -> List[int]:
        if len(self.free_blocks) < n:
            raise RuntimeError("No free blocks available")            # kv_cache.py:22-23, 57-58
        return [self.free_blocks.popleft() for _ in range(n)]

    def add_sequence(self, seq: int, prompt_len: int) -> List[torch.Tensor]:
        """Allocate blocks for a prompt; returns the per-layer int64 slot mappings [prompt_len]
        (reference: kv_cache.py:21-37, slot = block * B + offset)."""
        assert not self.active[seq] and 0 < prompt_len <= self.max_blocks_per_seq * self.block_size
        B = self.block_size
        nblk = (prompt_len + B - 1) // B
        pos = np.arange(prompt_len)
        if len(self.free_blocks) < nblk * self.num_layers:
            raise RuntimeError("No free blocks available")            # kv_cache.py:22-23
        for l in range(self.num_layers):                              # layer by layer, FIFO (kv_cache.py:25-26)
            self.h_tables[l, seq, :nblk] = np.asarray(self._take(nblk), dtype=np.int32)
        slots = self.h_tables[:, seq, pos // B].astype(np.int64) * B + pos % B      # [L, prompt_len]
        self.h_lens[seq] = prompt_len
        self.active[seq] = True
        self.d_tables[:, seq].copy_(torch.from_numpy(self.h_tables[:, seq]), non_blocking=True)
        d_slots = torch.from_numpy(np.ascontiguousarray(slots)).to(self.device)   # one upload for all layers
        return [d_slots[l] for l in range(self.num_layers)]

    def free(self, seq: int) -> None:
        """Return a sequence's blocks to the BACK of the free list (kv_cache.py:81-86)."""
        if not self.active[seq]:
            return
        used = self.h_tables[:, seq][self.h_tables[:, seq] >= 0]
        self.free_blocks.extend(int(b) for b in used)
        self.h_tables[:, seq] = -1
        self.h_lens[seq] = 0
        self.active[seq] = False
        self.d_tables[:, seq].fill_(-1)

    # -- one decode step for every active sequence ----------------------------------------------
    def begin_decode_step(self, attend_to_new_token: bool = True) -> None:
        """Reserve the slot of the token being decoded for every active sequence and publish
        `d_slots` [L, S], `d_tables`, `d_lens` for the kernels.  Batched counterpart of
        BlockManager.decode_step (block_manager.py:28-63): when a sequence's last block is full a
        new block is appended per layer (kv_cache.py:56-73).  Inactive rows get slot -1 (skipped by
        reshape_and_cache, cache_kernels.cu:166-169) and seq_len 0 (zeros from attention).

        attend_to_new_token=False reproduces the reference scheduler's off-by-one (it passes
        seq_lens = L after writing slot L, scheduler.py:96-101, so the new token does not attend to
        itself); True attends over L + 1 tokens like ordinary causal decoding."""
        B = self.block_size
        L, S = self.num_layers, self.max_seqs
        act = np.nonzero(self.active)[0]
        pos = self.h_lens[act].astype(np.int64)                    # position of the new token
        if (pos >= self.max_blocks_per_seq * B).any():
            raise RuntimeError("sequence exceeds max_blocks_per_seq * block_size")
        # next staging buffer of the ring; wait (host side) only if the GPU is a whole ring behind
        buf, ev, used = self._ring[self._ring_i]
        if used and ev is not None:
            ev.synchronize()
        hb = buf.numpy()
        h_slots = hb[: self._o_pos].view(np.int64).reshape(L, S)
        h_pos = hb[self._o_pos: self._o_grow].view(np.int64)
        h_grow = hb[self._o_grow: self._o_lens].view(np.int64)
        h_lens = hb[self._o_lens: self._o_blocks].view(np.int32)
        h_blocks = hb[self._o_blocks:].view(np.int32)

        need = act[pos % B == 0]                                    # last block full (or empty): append one
        n_new = int(need.size) * L
        if n_new:
            idx = (self.h_lens[need] // B).astype(np.int64)
            new_blocks = np.asarray(self._take(n_new), dtype=np.int32).reshape(L, need.size)
            self.h_tables[:, need, idx] = new_blocks
            MB = self.max_blocks_per_seq
            flat = (np.arange(L, dtype=np.int64)[:, None] * S + need[None, :]) * MB + idx[None, :]
            h_grow[:n_new] = flat.reshape(-1)
            h_blocks[:n_new] = new_blocks.reshape(-1)
        h_slots[:] = -1
        blk = self.h_tables[:, act, pos // B].astype(np.int64)      # [L, n_act]
        h_slots[:, act] = blk * B + pos % B
        h_lens[:] = 0
        h_lens[act] = self.h_lens[act] + (1 if attend_to_new_token else 0)
        h_pos[:] = 0
        h_pos[act] = pos

        # one H2D copy publishes d_slots / d_pos / d_lens; appended blocks are scattered into d_tables
        self._d_stage.copy_(buf, non_blocking=True)
        if ev is not None:
            ev.record()
        self._ring[self._ring_i][2] = True
        self._ring_i = (self._ring_i + 1) % len(self._ring)
        if n_new:
            self.d_tables.view(-1).scatter_(0, self._d_grow[:n_new], self._d_blocks[:n_new])
        self.h_lens[act] += 1

    def positions(self) -> np.ndarray:
        """Position id of the token being decoded in the step begun last (0 for inactive rows)."""
        p = np.zeros(self.max_seqs, dtype=np.int64)
        p[self.active] = self.h_lens[self.active] - 1
        return p


# ---------------------------------------------------------------------------------------------
# model
# ---------------------------------------------------------------------------------------------
def synthetic_gpt2_state_dict(shape: GPT2Shape, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights under the reference model's parameter names (no network: the HF
    checkpoint gpt2.py:276-331 loads cannot be downloaded).  CPU generator => identical on every box.
    Scale 0.02 like GPT-2's initialiser, LayerNorm weight ~ 1."""
    g = torch.Generator().manual_seed(seed)
    E, V, P = shape.n_embd, shape.vocab_size, shape.n_positions

    def rnd(*s, std=0.02):
        return (torch.randn(*s, generator=g) * std).to(dtype)

    sd = {"transformer.wte.weight": rnd(V, E), "transformer.wpe.weight": rnd(P, E, std=0.01),
          "transformer.ln_f.weight": (1 + rnd(E)).to(dtype), "transformer.ln_f.bias": rnd(E)}
    for i in range(shape.n_layer):
        p = f"transformer.h.{i}."
        sd.update({
            p + "ln_1.weight": (1 + rnd(E)).to(dtype), p + "ln_1.bias": rnd(E),
            p + "attn.c_attn.weight": rnd(3 * E, E), p + "attn.c_attn.bias": rnd(3 * E),
            p + "attn.c_proj.weight": rnd(E, E, std=0.02 / math.sqrt(2 * shape.n_layer)), p + "attn.c_proj.bias": rnd(E),
            p + "ln_2.weight": (1 + rnd(E)).to(dtype), p + "ln_2.bias": rnd(E),
            p + "mlp.c_fc.weight": rnd(4 * E, E), p + "mlp.c_fc.bias": rnd(4 * E),
            p + "mlp.c_proj.weight": rnd(E, 4 * E, std=0.02 / math.sqrt(2 * shape.n_layer)), p + "mlp.c_proj.bias": rnd(E),
        })
    sd["lm_head.weight"] = sd["transformer.wte.weight"]                 # tied (gpt2.py:247)
    return sd


class GPT2Decoder:
    """GPT-2 forward for serving: prefill (causal attention over the prompt, like gpt2.py:46-59, 71-78) and
    batched single-token decode through the paged-attention hot path.

    Parameter names and shapes are the reference GPT2LMHeadModel's (nn.Linear [out, in] weights,
    exact-erf GELU as gpt2.py:123), so `load_state_dict` accepts a reference state_dict."""

    def __init__(self, shape: GPT2Shape, device: torch.device, dtype=torch.float16, fused_append: bool = True,
                 fused_layernorm: bool = True, prefill_attention: str = "sdpa"):
        if prefill_attention not in ("sdpa", "paged"):
            raise ValueError("prefill_attention must be 'sdpa' or 'paged'")
        self.shape, self.device, self.dtype = shape, device, dtype
        # prefill attention: "sdpa" = a library kernel over the prompt's own q/k/v (what the reference does with
        # eager matmuls, gpt2.py:46-59, 71-78); "paged" = through the paged path: after
        # reshape_and_cache has scattered the prompt, paged_attention_v1 is called with one ROW per prompt token
        # -- every row shares the sequence's block table and row t has seq_len t + 1 -- which is causal attention
        # over the paged cache, with exactly the arithmetic of the decode steps that follow
        self.prefill_attention = prefill_attention
        # decode: one fused launch per layer (scatter + attention) instead of the reference's two calls
        self.fused_append = fused_append
        # decode: residual add + LayerNorm (and the embedding sum + ln_1) as one launch each instead of two
        # eager torch kernels; False keeps plain torch ops (the test reference for the fused path)
        self.fused_layernorm = fused_layernorm and dtype is torch.float16
        self.w: Dict[str, torch.Tensor] = {}
        self.scale = shape.head_size ** -0.5
        self._graph = None
        self._prefill_graphs: Dict = {}

    # -- weights -----------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        self.w = {k: v.detach().to(device=self.device, dtype=self.dtype).contiguous() for k, v in sd.items()
                  if k != "lm_head.weight"}
        self.w["lm_head.weight"] = self.w["transformer.wte.weight"]
        self._lm_head_padded = None
        self._refresh_derived()
        self._graph = None
        self._prefill_graphs = {}

    def _refresh_derived(self) -> None:
        """Decode-time lm_head operand: the tied embedding padded to a multiple of 64 rows.  vocab 50257 is odd, so
        the [S, E] x [E, V] GEMM has an unaligned leading dimension and cuBLAS falls back to a legacy kernel
        (182 us vs 23 us at S=256); the zero rows only add logits that are sliced away again.  Rebuilt in place
        whenever the weights change (load, broadcast) so that a captured graph keeps a valid pointer."""
        V, E = self.w["lm_head.weight"].shape
        if self._lm_head_padded is None or self._lm_head_padded.shape != ((V + 63) // 64 * 64, E):
            self._lm_head_padded = torch.zeros(((V + 63) // 64 * 64, E), dtype=self.dtype, device=self.device)
        self._lm_head_padded[:V].copy_(self.w["lm_head.weight"])

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return self.w

    def broadcast_weights(self, src: int = 0) -> None:
        """The path's only collective: replicate the weights from rank `src` (NCCL over NVLink)."""
        from . import sharding
        sharding.broadcast_state_dict(self.w, src=src)
        self._refresh_derived()

    # -- building blocks -----------------------------------------------------------------------
    def _ln(self, x, prefix):
        return F.layer_norm(x, (self.shape.n_embd,), self.w[prefix + ".weight"], self.w[prefix + ".bias"],
                            self.shape.layer_norm_epsilon)

    def _mlp(self, x, p):
        h = F.linear(x, self.w[p + "mlp.c_fc.weight"], self.w[p + "mlp.c_fc.bias"])
        return F.linear(F.gelu(h), self.w[p + "mlp.c_proj.weight"], self.w[p + "mlp.c_proj.bias"])

    # -- prefill: one sequence, causal attention over the prompt + reshape_and_cache -----------------------
    def _prefill_forward(self, input_ids: torch.Tensor, slot_maps, kv: BatchedKVState, last: torch.Tensor,
                         tables: torch.Tensor = None):
        """input_ids [T], slot_maps[layer] [T] (-1 = padding), `last` int64 [1] (index of the last real token)
        -> logits of that position [vocab].  `tables` [L, MB]: the sequence's block-table rows (paged prefill)."""
        s = self.shape
        T = input_ids.numel()
        H, D = s.n_head, s.head_size
        x = self.w["transformer.wte.weight"][input_ids] + self.w["transformer.wpe.weight"][:T]
        paged = self.prefill_attention == "paged"
        if paged:
            # row t attends over tokens 0..t; padding rows (t > last) attend over nothing and yield zeros
            pos = torch.arange(T, device=self.device)
            row_lens = torch.where(pos <= last, pos + 1, torch.zeros_like(pos)).to(torch.int32)
            max_seq_len = kv.max_blocks_per_seq * kv.block_size
        for i in range(s.n_layer):
            p = f"transformer.h.{i}."
            qkv = F.linear(self._ln(x, p + "ln_1"), self.w[p + "attn.c_attn.weight"], self.w[p + "attn.c_attn.bias"])
            q, k, v = (t.view(T, H, D) for t in qkv.split(s.n_embd, dim=-1))      # strided views, stride 3*E
            ops.reshape_and_cache(k, v, kv.key_cache, kv.value_cache, slot_maps[i], "auto", 1.0)
            if paged:
                # prefill THROUGH the paged path: T rows share one block table; K/V come from the cache just written
                o = torch.empty((T, H, D), dtype=self.dtype, device=self.device)
                ops.paged_attention_v1(o, q, kv.key_cache, kv.value_cache, H, self.scale,
                                       tables[i].expand(T, -1).contiguous(), row_lens, kv.block_size, max_seq_len,
                                       None, "auto", 1.0, 0, 0, 1, 1, 0)
            else:
                # causal attention over the prompt (gpt2.py:46-59, 71-78): a library kernel.  The memory-efficient
                # backend is pinned: cuDNN's builds a plan per prompt length (45-95 ms each on B200).
                with sdpa_kernel([SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH]):
                    o = F.scaled_dot_product_attention(q.transpose(0, 1)[None], k.transpose(0, 1)[None],
                                                       v.transpose(0, 1)[None], is_causal=True, scale=self.scale)[0]
                o = o.transpose(0, 1)
            o = o.reshape(T, s.n_embd)
            x = x + F.linear(o, self.w[p + "attn.c_proj.weight"], self.w[p + "attn.c_proj.bias"])
            x = x + self._mlp(self._ln(x, p + "ln_2"), p)
        x = self._ln(x.index_select(0, last), "transformer.ln_f")
        return F.linear(x, self.w["lm_head.weight"])[0]

    @torch.no_grad()
    def prefill(self, input_ids: torch.Tensor, kv: BatchedKVState, seq: int, graph_bucket: int = 0) -> torch.Tensor:
        """input_ids [T] -> logits of the last position [vocab].  K/V of all T tokens are scattered
        into the pool through the hot path's reshape_and_cache (gpt2.py:44).

        graph_bucket > 0: the prompt is padded to a multiple of `graph_bucket` tokens (padding carries slot -1, is
        never attended by real tokens under the causal mask and is never stored) and the forward pass is replayed
        from a CUDA graph captured once per padded length -- prefill is ~200 small launches, i.e. host-bound when
        issued eagerly.  The returned tensor is then a static buffer, valid until the next graphed prefill."""
        T = input_ids.numel()
        slot_maps = kv.add_sequence(seq, T)
        if graph_bucket <= 0 or self.device.type != "cuda":
            last = torch.full((1,), T - 1, dtype=torch.int64, device=self.device)
            return self._prefill_forward(input_ids, slot_maps, kv, last, kv.d_tables[:, seq])
        Tb = min((T + graph_bucket - 1) // graph_bucket * graph_bucket, self.shape.n_positions)
        key = (Tb, id(kv))
        st = self._prefill_graphs.get(key)
        if st is not None and st["kv"]() is not kv:      # the id was recycled by a different (newer) pool
            st = None
        if st is None:
            L = self.shape.n_layer
            st = {"kv": weakref.ref(kv),
                  "ids": torch.zeros(Tb, dtype=torch.int64, device=self.device),
                  "slots": torch.full((L, Tb), -1, dtype=torch.int64, device=self.device),
                  "tables": torch.full((L, kv.max_blocks_per_seq), -1, dtype=torch.int32, device=self.device),
                  "last": torch.zeros(1, dtype=torch.int64, device=self.device)}
            rows = [st["slots"][l] for l in range(L)]
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                              # warm-up outside capture (all slots -1)
                self._prefill_forward(st["ids"], rows, kv, st["last"], st["tables"])
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            st["graph"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(st["graph"]):
                st["logits"] = self._prefill_forward(st["ids"], rows, kv, st["last"], st["tables"])
            self._prefill_graphs[key] = st
        st["ids"].zero_()
        st["ids"][:T].copy_(input_ids)
        st["slots"].fill_(-1)
        st["slots"][:, :T].copy_(torch.stack(slot_maps))
        st["last"].fill_(T - 1)
        st["tables"].copy_(kv.d_tables[:, seq])
        st["graph"].replay()
        return st["logits"]

    # -- decode: all sequences, one token each ----------------------------------------------------
    @torch.no_grad()
    def decode_forward(self, input_ids: torch.Tensor, position_ids: torch.Tensor, kv: BatchedKVState) -> torch.Tensor:
        """input_ids, position_ids [S] (device) -> logits [S, vocab].  Uses kv.d_slots / d_tables /
        d_lens as published by kv.begin_decode_step().  Per layer: reshape_and_cache then
        paged_attention_v1 (gpt2.py:44, 62)."""
        s = self.shape
        S = input_ids.numel()
        if self.fused_layernorm:
            return self._decode_forward_fused(input_ids, position_ids, kv)
        x = self.w["transformer.wte.weight"][input_ids] + self.w["transformer.wpe.weight"][position_ids]
        for i in range(s.n_layer):
            p = f"transformer.h.{i}."
            qkv = F.linear(self._ln(x, p + "ln_1"), self.w[p + "attn.c_attn.weight"], self.w[p + "attn.c_attn.bias"])
            o = self._decode_attention(qkv, kv, i)
            x = x + F.linear(o.view(S, s.n_embd), self.w[p + "attn.c_proj.weight"], self.w[p + "attn.c_proj.bias"])
            x = x + self._mlp(self._ln(x, p + "ln_2"), p)
        return self._lm_head(self._ln(x, "transformer.ln_f"))

    def _decode_attention(self, qkv: torch.Tensor, kv: BatchedKVState, layer: int) -> torch.Tensor:
        """The hot path of one layer: reshape_and_cache + paged_attention_v1 (gpt2.py:44, 62), or their fusion."""
        s = self.shape
        S = qkv.shape[0]
        H, D, B = s.n_head, s.head_size, kv.block_size
        max_seq_len = kv.max_blocks_per_seq * B                       # capacity, as scheduler.py:97
        q, k, v = (t.view(S, H, D) for t in qkv.split(s.n_embd, dim=-1))
        o = torch.empty((S, H, D), dtype=self.dtype, device=self.device)
        if self.fused_append:
            ops.paged_attention_v1_append(o, q, k, v, kv.key_cache, kv.value_cache, kv.d_slots[layer], H, self.scale,
                                          kv.d_tables[layer], kv.d_lens, B, max_seq_len)
        else:
            ops.reshape_and_cache(k, v, kv.key_cache, kv.value_cache, kv.d_slots[layer], "auto", 1.0)
            ops.paged_attention_v1(o, q, kv.key_cache, kv.value_cache, H, self.scale, kv.d_tables[layer], kv.d_lens,
                                   B, max_seq_len, None, "auto", 1.0, 0, 0, 1, 1, 0)
        return o

    def _lm_head(self, h: torch.Tensor) -> torch.Tensor:
        return F.linear(h, self._lm_head_padded)[:, : self.shape.vocab_size]

    def _decode_forward_fused(self, input_ids: torch.Tensor, position_ids: torch.Tensor, kv: BatchedKVState):
        """Same arithmetic with every `x = x + y; h = LayerNorm(x)` pair (and the embedding sum + ln_1) done by
        ops.add_layernorm: 8 launches per layer instead of 10."""
        s, w = self.shape, self.w
        S, E, eps = input_ids.numel(), s.n_embd, s.layer_norm_epsilon
        x = torch.empty((S, E), dtype=self.dtype, device=self.device)
        h = torch.empty((S, E), dtype=self.dtype, device=self.device)
        ops.add_layernorm(w["transformer.wte.weight"], w["transformer.wpe.weight"], w["transformer.h.0.ln_1.weight"],
                          w["transformer.h.0.ln_1.bias"], eps, x, h, a_rows=input_ids, x_rows=position_ids)
        for i in range(s.n_layer):
            p = f"transformer.h.{i}."
            nxt = f"transformer.h.{i + 1}.ln_1" if i + 1 < s.n_layer else "transformer.ln_f"
            qkv = F.linear(h, w[p + "attn.c_attn.weight"], w[p + "attn.c_attn.bias"])
            o = self._decode_attention(qkv, kv, i)
            a = F.linear(o.view(S, E), w[p + "attn.c_proj.weight"], w[p + "attn.c_proj.bias"])
            ops.add_layernorm(a, x, w[p + "ln_2.weight"], w[p + "ln_2.bias"], eps, x, h)
            a = F.linear(F.gelu(F.linear(h, w[p + "mlp.c_fc.weight"], w[p + "mlp.c_fc.bias"])),
                         w[p + "mlp.c_proj.weight"], w[p + "mlp.c_proj.bias"])
            ops.add_layernorm(a, x, w[nxt + ".weight"], w[nxt + ".bias"], eps, x, h)
        return self._lm_head(h)

    def greedy(self, logits: torch.Tensor) -> torch.Tensor:
        """argmax over the vocabulary (one launch of the library's row-argmax when the fused helpers are on)."""
        if self.fused_layernorm and logits.dtype is torch.float16:
            return ops.argmax_rows(logits)
        return torch.argmax(logits, dim=-1)

    # -- CUDA-graph capture of the decode step -----------------------------------------------------
    def capture_decode_graph(self, kv: BatchedKVState, feed_back: bool = False) -> None:
        """Capture decode_forward + greedy argmax for kv.max_seqs rows into one CUDA graph.  Inputs
        live in static buffers (`g_ids`, and kv.d_pos / d_slots / d_tables / d_lens, which
        begin_decode_step updates in place; `g_pos` is kv.d_pos); outputs in `g_logits`, `g_next`.
        feed_back=True also copies `g_next` into `g_ids` at the end of the graph (free-running greedy decode)."""
        S = kv.max_seqs
        self.g_ids = torch.zeros(S, dtype=torch.int64, device=self.device)
        self.g_pos = kv.d_pos
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):                                   # warm-up outside capture
            for _ in range(2):
                self.g_logits = self.decode_forward(self.g_ids, self.g_pos, kv)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self.g_logits = self.decode_forward(self.g_ids, self.g_pos, kv)
            self.g_next = self.greedy(self.g_logits)
            if feed_back:                 # the sampled ids become the next step's input inside the graph
                self.g_ids.copy_(self.g_next)

    def replay_decode_graph(self) -> torch.Tensor:
        self._graph.replay()
        return self.g_next


# ---------------------------------------------------------------------------------------------
# greedy generation loop (harness)
# ---------------------------------------------------------------------------------------------
@torch.no_grad()
def greedy_generate(model: GPT2Decoder, kv: BatchedKVState, prompts: Sequence[Sequence[int]], max_new_tokens: int,
                    use_graph: bool = False) -> List[List[int]]:
    """Prefill every prompt, then decode all sequences together, greedy argmax.  Returns the new
    token ids per prompt.  The only device->host transfer is the final read of the ids."""
    dev = model.device
    n = len(prompts)
    assert n <= kv.max_seqs
    first = torch.zeros(kv.max_seqs, dtype=torch.int64, device=dev)
    for i, pr in enumerate(prompts):
        logits = model.prefill(torch.tensor(list(pr), dtype=torch.int64, device=dev), kv, i)
        first[i] = torch.argmax(logits)
    out = [first.clone()]
    cur = first
    if use_graph and max_new_tokens > 1:
        model.capture_decode_graph(kv)
    for _ in range(max_new_tokens - 1):
        kv.begin_decode_step(attend_to_new_token=True)              # publishes kv.d_pos (= model.g_pos) too
        if use_graph:
            model.g_ids.copy_(cur)
            cur = model.replay_decode_graph().clone()
        else:
            cur = model.greedy(model.decode_forward(cur, kv.d_pos, kv))
        out.append(cur)
    ids = torch.stack(out, dim=1)[:n].cpu().tolist()
    return ids
