"""Request-level sharding of the decode path across GPUs (one process per GPU).

Each sequence's attention touches only its own KV blocks (reference kernel: grid.y = sequence,
attention_kernels.cu:111), so the path partitions by request with NO data-path collective: every rank
owns a private KV pool, block tables and scheduler state for its shard.  torch.distributed is used for
exactly two things: broadcasting the (replicated) model weights from rank 0 at init, and reducing
timings / gathering results on the host side.  Works with the `nccl` backend on GPUs (NVLink/NVSwitch)
and with `gloo` on CPU (tests).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_requests: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of `num_requests` for `rank` (sizes differ by <= 1)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size: {rank}/{world_size}")
    base, rem = divmod(num_requests, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_requests(requests: Sequence, world_size: int, rank: int) -> List:
    b, e = shard_bounds(len(requests), world_size, rank)
    return list(requests[b:e])


def balanced_shards_by_length(seq_lens: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-first greedy assignment of request indices to ranks so that the summed context
    (= KV bytes the attention kernel must read) is balanced; used for mixed-length batches."""
    order = sorted(range(len(seq_lens)), key=lambda i: -int(seq_lens[i]))
    loads = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda j: (loads[j], len(shards[j])))
        shards[r].append(i)
        loads[r] += int(seq_lens[i])
    for s in shards:
        s.sort()
    return shards


def broadcast_state_dict(state: Dict[str, torch.Tensor], src: int = 0, bucket_bytes: int = 256 << 20) -> None:
    """In-place broadcast of every tensor of `state` from rank `src` (weights are replicated).

    Tensors are packed per dtype into flat buckets so GPT-2 small's ~250 MB of fp16 weights go out
    in one or two NCCL broadcasts over NVLink instead of ~150 small ones.  No-op without an
    initialised process group or with world_size 1."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    by_key: Dict[Tuple[torch.dtype, torch.device], List[torch.Tensor]] = {}
    seen = set()
    for name in sorted(state):
        t = state[name]
        if t.data_ptr() in seen:       # tied weights (lm_head.weight is wte.weight, gpt2.py:247)
            continue
        seen.add(t.data_ptr())
        by_key.setdefault((t.dtype, t.device), []).append(t)
    for (_dtype, _device), tensors in by_key.items():
        bucket: List[torch.Tensor] = []
        size = 0
        for t in tensors + [None]:
            if t is not None and (size + t.numel() * t.element_size() <= bucket_bytes or not bucket):
                bucket.append(t)
                size += t.numel() * t.element_size()
                continue
            flat = torch.cat([b.reshape(-1) for b in bucket])
            dist.broadcast(flat, src=src)
            off = 0
            for b in bucket:
                n = b.numel()
                b.copy_(flat[off:off + n].view_as(b))
                off += n
            bucket, size = ([t], t.numel() * t.element_size()) if t is not None else ([], 0)


def replicas_agree(state: Dict[str, torch.Tensor]) -> bool:
    """True when every rank holds the same weights: per-tensor fp64 checksums (sum and sum of squares) are
    all_reduced with MAX and MIN and must coincide.  One small collective, outside any timed region; trivially True
    without a process group."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return True
    names = sorted(state)
    dev = state[names[0]].device
    sums = torch.stack([torch.stack([state[n].double().sum(), state[n].double().square().sum()]) for n in names]).reshape(-1)
    sums = torch.nan_to_num(sums).to(dev)
    hi, lo = sums.clone(), sums.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return bool(torch.equal(hi, lo))


def reduce_sum(values: Sequence[float], device: torch.device) -> List[float]:
    """Sum over ranks of a few host scalars (token counts)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def reduce_max(values: Sequence[float], device: torch.device) -> List[float]:
    """Max over ranks of a few host scalars (timings): the job is as slow as its slowest rank."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def gather_token_ids(local_ids: torch.Tensor, counts: Sequence[int]) -> torch.Tensor:
    """Concatenate each rank's generated token ids [n_r, T] on every rank (outside the decode path)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_ids
    T = local_ids.shape[1]
    m = max(int(c) for c in counts)                     # all_gather needs equal shapes: pad to the largest shard
    padded = torch.zeros((m, T), dtype=local_ids.dtype, device=local_ids.device)
    padded[: local_ids.shape[0]] = local_ids
    bufs = [torch.empty_like(padded) for _ in counts]
    dist.all_gather(bufs, padded)
    return torch.cat([b[: int(c)] for b, c in zip(bufs, counts)], dim=0)
