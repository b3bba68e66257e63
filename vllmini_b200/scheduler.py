"""Continuous-batching scheduler over the batched decode driver.

Batched counterpart of the reference's `Scheduler` (vllmini/scheduler.py:10-150).  The reference keeps a
priority queue of sequence ids and advances ONE sequence per model call (scheduler.py:55-108); this class keeps
the same life cycle per request --

    add_sequence(input_ids)   prefill, keep the last logits                      (scheduler.py:22-53)
    every step                sample from the last logits, append the token, run the model on it,
                              length += 1                                         (scheduler.py:75-101)
    stop                      token == eos or length >= max_length; blocks freed  (scheduler.py:103-108, 133-139)
    out of blocks             the most recently arrived sequence is dropped       (scheduler.py:117-131)

-- but advances every active request in ONE batched step through `GPT2Decoder` / `BatchedKVState`, admits waiting
requests into rows as they free up (continuous batching) and overlaps the host bookkeeping of step t with the
GPU work of step t: the sampled ids are copied to pinned memory asynchronously and only read after the model
step has been enqueued.

`sampling="top_k"` is the reference's sampler (temperature 1.0, top-k 50, multinomial, scheduler.py:141-149) done
for all rows at once; `"greedy"` is argmax.  `reference_off_by_one=True` passes seq_lens without the new token,
as scheduler.py:96-101 does.
"""
from __future__ import annotations

import collections
import itertools
import time
from typing import Deque, Dict, List, Optional, Sequence, Tuple

import torch

from .decode import BatchedKVState


class BatchedScheduler:
    def __init__(self, model, kv: BatchedKVState, max_length: int, eos_token_id: Optional[int] = None,
                 sampling: str = "greedy", top_k: int = 50, temperature: float = 1.0, seed: int = 0,
                 use_graph: bool = False, reference_off_by_one: bool = False, prefill_bucket: int = 128):
        if sampling not in ("greedy", "top_k"):
            raise ValueError("sampling must be 'greedy' or 'top_k'")
        self.model, self.kv = model, kv
        # a sequence can never outgrow its block-table row, nor the model's position table (wpe has n_positions rows;
        # the fused embedding launch traps on a position beyond it instead of reading out of bounds)
        self.max_length = min(int(max_length), kv.max_blocks_per_seq * kv.block_size)
        n_positions = getattr(getattr(model, "shape", None), "n_positions", None)
        if n_positions is not None:
            self.max_length = min(self.max_length, int(n_positions))
        self.eos_token_id = eos_token_id
        self.sampling, self.top_k, self.temperature = sampling, int(top_k), float(temperature)
        self.use_graph = use_graph
        self.prefill_bucket = int(prefill_bucket) if use_graph else 0     # graphed prefill, padded to this multiple
        self.attend_to_new_token = not reference_off_by_one
        self.device = kv.device
        S = kv.max_seqs
        # reference-named bookkeeping (scheduler.py:17-21)
        self.active_sequences: Dict[int, float] = {}          # seq_id -> arrival time
        self.sequence_lengths: Dict[int, int] = {}
        self.sequences: Dict[int, List[int]] = {}             # prompt + generated tokens, kept after completion
        self.finished: Dict[int, str] = {}                    # seq_id -> "eos" | "max_length" | "evicted"
        self.rows: List[Optional[int]] = [None] * S           # row -> seq_id
        self.row_of: Dict[int, int] = {}
        self.pending: Deque[Tuple[int, List[int]]] = collections.deque()
        self._ids = itertools.count()
        self._gen = torch.Generator(device=self.device).manual_seed(seed)
        self._logits: Optional[torch.Tensor] = None           # [S, vocab] last logits of every row (device)
        self._cur = torch.zeros(S, dtype=torch.int64, device=self.device)
        self._pin = torch.zeros(S, dtype=torch.int64)
        self._ev = None
        if self.device.type == "cuda":
            self._pin = self._pin.pin_memory()
            self._ev = torch.cuda.Event()
        self._graph_ready = False
        self.steps = 0
        self.tokens_generated = 0

    # -- admission --------------------------------------------------------------------------------
    def add_sequence(self, input_ids) -> int:
        """Queue a prompt ([1, T] / [T] tensor or a list of ids); it is prefilled as soon as a row and enough
        blocks are free (immediately if possible, like the reference).  Returns the sequence id."""
        ids = input_ids.reshape(-1).tolist() if isinstance(input_ids, torch.Tensor) else [int(t) for t in input_ids]
        if not ids:
            raise ValueError("empty prompt")
        seq_id = next(self._ids)
        self.active_sequences[seq_id] = time.time()
        self.sequences[seq_id] = list(ids)
        self.sequence_lengths[seq_id] = len(ids)
        self.pending.append((seq_id, ids))
        self._admit()
        return seq_id

    def _blocks_needed(self, prompt_len: int) -> int:
        """Blocks an admission must find free: the prompt's, the block its FIRST decoded token may open (prompt
        length a multiple of the block size) and the blocks the rows already running will open in the coming step
        -- otherwise the step right after a prefill could run out of blocks and the eviction rule (newest sequence
        goes, scheduler.py:117-131) would throw away the prompt that was just prefilled."""
        kv = self.kv
        B = kv.block_size
        own = ((prompt_len + 1 + B - 1) // B) * kv.num_layers
        growing = int(((kv.h_lens[kv.active] % B) == 0).sum()) * kv.num_layers
        return own + growing

    def _admit(self) -> None:
        kv = self.kv
        while self.pending:
            seq_id, ids = self.pending[0]
            if len(ids) > min(kv.max_blocks_per_seq * kv.block_size, self.max_length):
                self.pending.popleft()           # can never run: longer than a table row / the position table
                self._retire(seq_id, "evicted")
                continue
            try:
                row = self.rows.index(None)
            except ValueError:
                return
            if len(kv.free_blocks) < self._blocks_needed(len(ids)):
                return
            self.pending.popleft()
            self.rows[row] = seq_id
            self.row_of[seq_id] = row
            t_ids = torch.tensor(ids, dtype=torch.int64, device=self.device)
            if self.prefill_bucket:
                logits = self.model.prefill(t_ids, kv, row, graph_bucket=self.prefill_bucket)
            else:
                logits = self.model.prefill(t_ids, kv, row)
            if self._logits is None:
                self._logits = torch.zeros((kv.max_seqs, logits.shape[-1]), dtype=logits.dtype, device=self.device)
            self._logits[row].copy_(logits)

    # -- sampling ----------------------------------------------------------------------------------
    def _sample(self, logits: torch.Tensor) -> torch.Tensor:
        if self.sampling == "greedy":
            return self.model.greedy(logits) if hasattr(self.model, "greedy") else torch.argmax(logits, dim=-1)
        scaled = logits.float() / self.temperature                       # scheduler.py:141-149, all rows at once
        vals, idx = torch.topk(scaled, min(self.top_k, scaled.shape[-1]), dim=-1)
        probs = torch.softmax(vals, dim=-1)
        pick = torch.multinomial(probs, num_samples=1, generator=self._gen)
        return idx.gather(-1, pick).squeeze(-1)

    # -- one batched decode step -----------------------------------------------------------------------
    def step(self) -> int:
        """Advance every active row by one token.  Returns the number of rows that were active."""
        # a sequence that already reached max_length generates nothing more (scheduler.py:70-73)
        for seq_id in [s for s in self.rows if s is not None]:
            if self.sequence_lengths[seq_id] >= self.max_length:
                self._retire(seq_id, "max_length")
        self._admit()
        active_rows = [r for r, s in enumerate(self.rows) if s is not None]
        if not active_rows:
            if self.pending:        # nothing is running and the oldest waiting prompt still does not fit: it never will
                self._retire(self.pending.popleft()[0], "evicted")
            return 0
        kv, model = self.kv, self.model
        # 1. sample the token each row consumes in this step; start its copy to the host
        self._cur.copy_(self._sample(self._logits))
        self._pin.copy_(self._cur, non_blocking=True)
        if self._ev is not None:
            self._ev.record()
        # 2. reserve the slots; out of blocks -> drop the newest sequence and retry (scheduler.py:117-131)
        while True:
            try:
                kv.begin_decode_step(attend_to_new_token=self.attend_to_new_token)
                break
            except RuntimeError as e:
                if "No free blocks" not in str(e) or len(active_rows) <= 1:
                    raise
                newest = max((self.rows[r] for r in active_rows), key=self.active_sequences.get)
                active_rows.remove(self.row_of[newest])
                self._retire(newest, "evicted")
        # 3. enqueue the model step for all rows
        if self.use_graph:
            if not self._graph_ready:
                model.capture_decode_graph(kv)
                self._graph_ready = True
            model.g_ids.copy_(self._cur)
            model.replay_decode_graph()
            self._logits = model.g_logits
        else:
            self._logits = model.decode_forward(self._cur, kv.d_pos, kv)
        # 4. host bookkeeping of this step's tokens while the GPU runs the step
        if self._ev is not None:
            self._ev.synchronize()
        toks = self._pin.numpy()
        for r in active_rows:
            seq_id = self.rows[r]
            tok = int(toks[r])
            self.sequences[seq_id].append(tok)
            self.sequence_lengths[seq_id] += 1
            self.tokens_generated += 1
            if self.eos_token_id is not None and tok == self.eos_token_id:
                self._retire(seq_id, "eos")
            elif self.sequence_lengths[seq_id] >= self.max_length:
                self._retire(seq_id, "max_length")
        self.steps += 1
        return len(active_rows)

    def _retire(self, seq_id: int, why: str) -> None:
        """remove_sequence_from_processing (scheduler.py:133-139): free the blocks, forget the state; the token
        list stays in `sequences` until remove_completed_sequence."""
        row = self.row_of.pop(seq_id, None)
        if row is not None:
            self.kv.free(row)
            self.rows[row] = None
        self.active_sequences.pop(seq_id, None)
        self.sequence_lengths.pop(seq_id, None)
        self.finished[seq_id] = why

    def remove_completed_sequence(self, seq_id: int) -> None:                 # scheduler.py:141-143
        self.sequences.pop(seq_id, None)

    def run(self, max_steps: Optional[int] = None) -> int:
        """Step until nothing is active or waiting (scheduler.py:55).  Returns the number of steps taken."""
        n = 0
        while (self.pending or any(s is not None for s in self.rows)) and (max_steps is None or n < max_steps):
            self.step()
            n += 1
        return n

    def generated(self, seq_id: int, prompt_len: int) -> Sequence[int]:
        return self.sequences[seq_id][prompt_len:]
