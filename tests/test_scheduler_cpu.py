"""CPU tests of the continuous-batching scheduler's host logic (vllmini_b200/scheduler.py) with a stub model:
admission, per-request stop conditions, row reuse, eviction on block exhaustion -- the life cycle of the
reference's Scheduler (vllmini/scheduler.py:22-139)."""
import numpy as np
import torch

from vllmini_b200.decode import BatchedKVState
from vllmini_b200.scheduler import BatchedScheduler

V = 64


class StubModel:
    """logits are one-hot at (7 * token + position) % V: the greedy continuation of any prompt is computable by
    hand and depends only on that request's own tokens (so batching must not change it)."""

    def __init__(self):
        self.prefills = []

    @staticmethod
    def _next(tok, pos):
        return (7 * int(tok) + int(pos)) % V

    def prefill(self, ids, kv, row):
        kv.add_sequence(row, ids.numel())
        self.prefills.append(row)
        out = torch.zeros(V)
        out[self._next(ids[-1], ids.numel() - 1)] = 1.0
        return out

    def decode_forward(self, ids, pos, kv):
        out = torch.zeros(ids.numel(), V)
        for r in range(ids.numel()):
            if kv.active[r]:
                out[r, self._next(ids[r], pos[r])] = 1.0
        return out


def expected(prompt, max_length, eos=None):
    seq = list(prompt)
    while len(seq) < max_length:
        tok = StubModel._next(seq[-1], len(seq) - 1)
        seq.append(tok)
        if eos is not None and tok == eos:
            break
    return seq


def make(S=3, L=2, MB=6, B=4, num_blocks=None):
    return BatchedKVState(num_blocks or L * S * MB, L, num_heads=1, head_size=8, block_size=B, max_seqs=S,
                          max_blocks_per_seq=MB, device=torch.device("cpu"))


def test_more_requests_than_rows_each_matches_its_solo_continuation():
    kv = make(S=3)
    sch = BatchedScheduler(StubModel(), kv, max_length=20)
    rng = np.random.default_rng(0)
    prompts = [rng.integers(0, V, size=n).tolist() for n in (1, 4, 5, 9, 2, 7, 3)]
    ids = [sch.add_sequence(torch.tensor([p])) for p in prompts]            # [1, T] like the reference
    assert len(sch.pending) == 4 and sch.rows.count(None) == 0                # three admitted at once
    steps = sch.run()
    assert steps > 0 and not sch.pending and sch.rows.count(None) == 3 and not sch.active_sequences
    for i, p in zip(ids, prompts):
        assert sch.sequences[i] == expected(p, 20) and sch.finished[i] == "max_length"
    assert len(kv.free_blocks) == kv.num_blocks                                # every block came back
    assert sch.tokens_generated == sum(20 - len(p) for p in prompts)
    sch.remove_completed_sequence(ids[0])
    assert ids[0] not in sch.sequences


def test_eos_stops_a_request_and_frees_its_row_for_the_next_one():
    p0, p1, p2 = [3], [10, 11], [5, 6, 7]
    eos = expected(p0, 12)[4]                                  # a token request 0 will produce at its 4th step
    kv = make(S=2)
    sch = BatchedScheduler(StubModel(), kv, max_length=12, eos_token_id=eos)
    a, b, c = (sch.add_sequence(p) for p in (p0, p1, p2))
    sch.run()
    for i, p in ((a, p0), (b, p1), (c, p2)):
        want = expected(p, 12, eos)
        assert sch.sequences[i] == want
        assert sch.finished[i] == ("eos" if want[-1] == eos and len(want) < 12 else "max_length")
    assert sch.finished[a] == "eos" and sch.sequences[a][-1] == eos            # the eos token is kept (scheduler.py:84)


def test_prompt_at_max_length_generates_nothing_and_oversized_prompt_is_evicted():
    kv = make(S=2, MB=3, B=4)                                   # capacity 12 tokens per sequence
    sch = BatchedScheduler(StubModel(), kv, max_length=50)      # clamped to the table capacity
    assert sch.max_length == 12
    a = sch.add_sequence(list(range(12)))
    b = sch.add_sequence(list(range(13)))
    c = sch.add_sequence([1, 2])
    sch.run()
    assert sch.sequences[a] == list(range(12)) and sch.finished[a] == "max_length"
    assert sch.finished[b] == "evicted" and sch.sequences[b] == list(range(13))
    assert sch.sequences[c] == expected([1, 2], 12)


def test_block_exhaustion_drops_the_newest_sequence_and_the_rest_finish():
    # 2 layers x block 4: each request needs 2 blocks per 4 tokens (6 to reach 12 tokens); the pool holds 14 blocks
    kv = make(S=3, L=2, MB=6, B=4, num_blocks=14)
    sch = BatchedScheduler(StubModel(), kv, max_length=12)
    a = sch.add_sequence([1, 2, 3])            # 2 blocks
    b = sch.add_sequence([4, 5, 6])            # 2 blocks
    c = sch.add_sequence([7, 8, 9])            # 2 blocks -> 6 used, 12 at 5+ tokens; 9+ tokens would need 18
    sch.run()
    assert sch.finished[c] == "evicted"        # the most recently arrived one (scheduler.py:117-131)
    assert sch.sequences[a] == expected([1, 2, 3], 12) and sch.sequences[b] == expected([4, 5, 6], 12)
    assert len(kv.free_blocks) == 14


def test_top_k_sampling_is_seeded_and_stays_inside_the_top_k():
    class FlatModel(StubModel):
        def decode_forward(self, ids, pos, kv):
            out = torch.zeros(ids.numel(), V)
            out[:, 10:15] = 5.0                 # five equally likely candidates, everything else far below
            return out

        def prefill(self, ids, kv, row):
            kv.add_sequence(row, ids.numel())
            out = torch.zeros(V)
            out[10:15] = 5.0
            return out

    outs = []
    for _ in range(2):
        sch = BatchedScheduler(FlatModel(), make(S=2), max_length=16, sampling="top_k", top_k=5, seed=123)
        i = sch.add_sequence([1])
        j = sch.add_sequence([2, 3])
        sch.run()
        outs.append((sch.sequences[i], sch.sequences[j]))
        assert all(10 <= t < 15 for t in sch.sequences[i][1:] + sch.sequences[j][2:])
    assert outs[0] == outs[1]                   # same seed, same draws
    assert len(set(outs[0][0][1:])) > 1         # and it does sample


def test_admission_reserves_the_first_decode_block_so_a_fresh_prefill_is_never_evicted():
    """a prompt whose length is a multiple of the block size needs a new block per layer for its first
    decoded token.  With only the prompt's blocks free, admitting it would make the very next step run out of
    blocks and the eviction rule would drop the request that was just prefilled; it must wait in `pending`."""
    B, L = 4, 2
    kv = make(S=2, L=L, MB=6, B=B, num_blocks=2 * L + 2 * L + 1)          # 9 blocks
    sch = BatchedScheduler(StubModel(), kv, max_length=10)
    a = sch.add_sequence(list(range(1, 8)))       # 7 tokens: 2 blocks/layer = 4 blocks; 5 free
    assert a in sch.row_of and len(kv.free_blocks) == 5
    # 8 tokens = 2 full blocks/layer = 4 blocks, and its first decoded token opens a third (2 more): 6 > 5 free
    b = sch.add_sequence(list(range(10, 18)))
    assert b not in sch.row_of and len(sch.pending) == 1, "admitted without room for its first decode step"
    sch.run()
    assert sch.finished[a] == "max_length" and sch.finished[b] == "max_length"      # b ran once a's blocks came back
    assert sch.sequences[b] == expected(list(range(10, 18)), 10)
    assert "evicted" not in sch.finished.values()


def test_max_length_is_clamped_to_the_models_position_table():
    """positions index wpe, which has n_positions rows; the scheduler must never run past it."""
    class Shaped(StubModel):
        class shape:
            n_positions = 9
    kv = make(S=1, MB=6, B=4)                     # table capacity 24 tokens
    sch = BatchedScheduler(Shaped(), kv, max_length=20)
    assert sch.max_length == 9
    s = sch.add_sequence([1, 2, 3])
    sch.run()
    assert len(sch.sequences[s]) == 9 and sch.finished[s] == "max_length"
    too_long = sch.add_sequence(list(range(10)))  # longer than the position table: can never run
    sch.run()
    assert sch.finished[too_long] == "evicted"
