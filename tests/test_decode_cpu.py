"""CPU tests of the batched decode driver's host logic (block bookkeeping), device='cpu'."""
import pytest
import torch

from vllmini_b200.decode import BatchedKVState, GPT2Shape, synthetic_gpt2_state_dict


def make_state(num_blocks=64, L=3, S=4, MB=5, B=4):
    return BatchedKVState(num_blocks, L, num_heads=2, head_size=8, block_size=B, max_seqs=S,
                          max_blocks_per_seq=MB, device=torch.device("cpu"))


def test_prefill_allocation_matches_reference_conventions():
    kv = make_state()
    maps = kv.add_sequence(1, prompt_len=6)                 # 2 blocks per layer, FIFO from the front
    assert [m.dtype for m in maps] == [torch.int64] * 3
    assert kv.h_tables[0, 1].tolist() == [0, 1, -1, -1, -1]     # row padded with -1 (kv_cache.py:31)
    assert kv.h_tables[1, 1].tolist() == [2, 3, -1, -1, -1]     # next layer: next blocks (kv_cache.py:25-26)
    assert maps[0].tolist() == [0, 1, 2, 3, 4, 5]               # slot = block*B + offset (kv_cache.py:35)
    assert maps[2].tolist() == [16, 17, 18, 19, 20, 21]
    assert kv.h_lens[1] == 6 and list(kv.free_blocks)[:2] == [6, 7]
    assert torch.equal(kv.d_tables[:, 1], torch.from_numpy(kv.h_tables[:, 1]))


def test_decode_step_slots_block_append_and_lens():
    kv = make_state()
    kv.add_sequence(0, 3)       # blocks 0,1,2 (one per layer), 3 of 4 slots used
    kv.add_sequence(2, 4)       # blocks 3,4,5, last block exactly full
    kv.begin_decode_step(attend_to_new_token=True)
    # seq 0 writes into offset 3 of its block; seq 2 needs a fresh block per layer (6,7,8)
    assert kv.d_slots[:, 0].tolist() == [0 * 4 + 3, 1 * 4 + 3, 2 * 4 + 3]
    assert kv.h_tables[:, 2, 1].tolist() == [6, 7, 8]
    assert kv.d_slots[:, 2].tolist() == [24, 28, 32]
    assert kv.d_slots[:, 1].tolist() == [-1, -1, -1] and kv.d_slots[:, 3].tolist() == [-1, -1, -1]   # inactive rows
    assert kv.d_lens.tolist() == [4, 0, 5, 0]
    assert kv.positions().tolist() == [3, 0, 4, 0]
    assert kv.d_pos.tolist() == [3, 0, 4, 0]                  # device copy of the position ids (graph input)
    assert torch.equal(kv.d_tables, torch.from_numpy(kv.h_tables))
    # the reference scheduler's off-by-one (scheduler.py:96-101): seq_lens excludes the new token
    kv.begin_decode_step(attend_to_new_token=False)
    assert kv.d_lens.tolist() == [4, 0, 5, 0] and kv.h_lens.tolist() == [5, 0, 6, 0]
    assert kv.h_tables[:, 0, 1].tolist() == [9, 10, 11]       # seq 0 crossed its block boundary now


def test_slots_never_collide_over_many_steps_and_free_recycles_fifo():
    kv = make_state(num_blocks=200, L=2, S=4, MB=10, B=4)
    for s, n in enumerate((1, 4, 7, 9)):
        kv.add_sequence(s, n)
    seen = set()
    for _ in range(20):
        kv.begin_decode_step()
        slots = kv.d_slots.numpy()
        flat = slots[slots >= 0].tolist()
        assert len(set(flat)) == len(flat) and not (set(flat) & seen)
        seen |= set(flat)
        # every slot lies in a block its (layer, seq) owns
        for l in range(2):
            for s in range(4):
                assert slots[l, s] // 4 in kv.h_tables[l, s]
    used = sorted(int(b) for b in kv.h_tables[kv.h_tables >= 0])
    assert len(used) == len(set(used))                       # a block belongs to one (sequence, layer)
    before = list(kv.free_blocks)
    owned = [int(b) for b in kv.h_tables[:, 1][kv.h_tables[:, 1] >= 0]]
    kv.free(1)
    assert list(kv.free_blocks) == before + owned            # appended at the back (kv_cache.py:83)
    assert not kv.active[1] and (kv.h_tables[:, 1] == -1).all() and (kv.d_tables[:, 1] == -1).all()


def test_exhaustion_and_capacity_errors():
    kv = make_state(num_blocks=6, L=3, S=2, MB=2, B=4)
    kv.add_sequence(0, 4)
    with pytest.raises(RuntimeError, match="No free blocks"):
        kv.add_sequence(1, 5)                                 # needs 6, only 3 left
    kv = make_state(num_blocks=60, L=3, S=1, MB=2, B=4)
    kv.add_sequence(0, 8)
    with pytest.raises(RuntimeError, match="max_blocks_per_seq"):
        kv.begin_decode_step()


def test_synthetic_weights_are_reproducible_and_reference_named():
    shape = GPT2Shape(n_layer=2, n_head=4, n_embd=64, vocab_size=101, n_positions=32)
    a, b = synthetic_gpt2_state_dict(shape, seed=3), synthetic_gpt2_state_dict(shape, seed=3)
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    assert a["lm_head.weight"] is a["transformer.wte.weight"]                      # tied (gpt2.py:247)
    assert a["transformer.h.1.attn.c_attn.weight"].shape == (192, 64)             # nn.Linear [out, in]
    assert a["transformer.h.0.mlp.c_fc.weight"].shape == (256, 64)
    assert not torch.equal(a["transformer.wte.weight"], synthetic_gpt2_state_dict(shape, seed=4)["transformer.wte.weight"])
