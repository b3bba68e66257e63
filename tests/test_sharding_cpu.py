"""CPU tests of the multi-GPU plumbing: request sharding, weight broadcast and max-over-ranks timing
with a real 2-process `gloo` group (world_size 2)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vllmini_b200 import sharding


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 256, 2048, 2049):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.shard_bounds(2048, 8, 3) == (768, 1024)       # BASELINE configs[4]: 256 per GPU
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def test_balanced_shards_by_length():
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(128, 2049, (256,), generator=g).tolist()
    shards = sharding.balanced_shards_by_length(lens, 8)
    assert sorted(i for s in shards for i in s) == list(range(256))
    loads = [sum(lens[i] for i in s) for s in shards]
    assert (max(loads) - min(loads)) / (sum(loads) / 8) < 0.02    # within 2 % of the mean load


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                      # ranks start with DIFFERENT weights
        wte = torch.randn(50, 16)
        state = {"transformer.wte.weight": wte, "lm_head.weight": wte,               # tied
                 "h.0.ln.weight": torch.randn(16), "h.0.attn.w": torch.randn(16, 48).half(),
                 "h.0.attn.b": torch.randn(48).half(), "step": torch.tensor([rank], dtype=torch.int64)}
        assert not sharding.replicas_agree(state)            # ... and the checksum all_reduce sees it
        sharding.broadcast_state_dict(state, src=0, bucket_bytes=1024)   # small buckets: exercise the splitting
        assert sharding.replicas_agree(state)                # one all_reduce: every rank holds rank 0's weights now
        assert sharding.reduce_sum([1.0 + rank, 2.0], torch.device("cpu")) == [3.0, 4.0]
        torch.manual_seed(100)
        wte0 = torch.randn(50, 16)
        assert torch.equal(state["transformer.wte.weight"], wte0)
        assert state["lm_head.weight"].data_ptr() == state["transformer.wte.weight"].data_ptr()
        assert torch.equal(state["h.0.ln.weight"], torch.randn(16))
        assert torch.equal(state["h.0.attn.w"], torch.randn(16, 48).half())
        assert int(state["step"]) == 0
        # the decode driver's weight broadcast also refreshes what it derived from the weights (padded lm_head)
        from vllmini_b200.decode import GPT2Decoder, GPT2Shape, synthetic_gpt2_state_dict
        shape = GPT2Shape(n_layer=1, n_head=2, n_embd=16, vocab_size=37, n_positions=8)
        model = GPT2Decoder(shape, torch.device("cpu"), dtype=torch.float32)
        model.load_state_dict(synthetic_gpt2_state_dict(shape, seed=rank))
        model.broadcast_weights(src=0)
        wte_rank0 = synthetic_gpt2_state_dict(shape, seed=0)["transformer.wte.weight"]
        assert torch.equal(model.w["transformer.wte.weight"], wte_rank0) and sharding.replicas_agree(model.state_dict())
        assert model._lm_head_padded.shape == (64, 16) and torch.equal(model._lm_head_padded[:37], wte_rank0)
        assert not model._lm_head_padded[37:].any()
        # request sharding + host-side gather of results
        reqs = list(range(11))
        mine = sharding.shard_requests(reqs, world, rank)
        counts = [len(sharding.shard_requests(reqs, world, r)) for r in range(world)]
        ids = torch.tensor([[r * 10, r * 10 + 1] for r in mine], dtype=torch.int64)
        allids = sharding.gather_token_ids(ids, counts)
        assert allids[:, 0].tolist() == [r * 10 for r in reqs]
        # timing: max over ranks
        got = sharding.reduce_max([1.0 + rank, 5.0 - rank], torch.device("cpu"))
        assert got == [float(world), 5.0]
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_broadcast_shard_and_reduce(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
