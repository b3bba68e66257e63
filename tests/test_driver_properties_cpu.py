"""Property tests (hypothesis) of the batched driver's host logic on CPU: block accounting of BatchedKVState under
random add / step / free interleavings, and the scheduler's per-request results under random workloads."""
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from test_scheduler_cpu import StubModel, expected
from vllmini_b200.decode import BatchedKVState
from vllmini_b200.scheduler import BatchedScheduler

L, S, MB, B = 2, 5, 6, 4


def check_invariants(kv, shadow_lens):
    owned = kv.h_tables[kv.h_tables >= 0].tolist()
    assert len(owned) == len(set(owned)), "a block belongs to one (sequence, layer)"
    assert not (set(owned) & set(kv.free_blocks)) and len(owned) + len(kv.free_blocks) == kv.num_blocks
    for s_ in range(S):
        if kv.active[s_]:
            n = int(kv.h_lens[s_])
            assert n == shadow_lens[s_]
            nblk = (n + B - 1) // B
            assert (kv.h_tables[:, s_, :nblk] >= 0).all() and (kv.h_tables[:, s_, nblk:] == -1).all()
        else:
            assert (kv.h_tables[:, s_] == -1).all() and kv.h_lens[s_] == 0
    assert torch.equal(kv.d_tables, torch.from_numpy(kv.h_tables))


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.sampled_from(["add", "step", "free"]), st.integers(0, S - 1), st.integers(1, MB * B - 1)),
                min_size=1, max_size=60))
def test_block_accounting_under_random_interleavings(ops_list):
    kv = BatchedKVState(L * S * MB, L, 1, 8, B, S, MB, torch.device("cpu"))
    lens = {}
    for op, row, n in ops_list:
        if op == "add" and not kv.active[row]:
            kv.add_sequence(row, n)
            lens[row] = n
        elif op == "free":
            kv.free(row)
            lens.pop(row, None)
        elif op == "step" and kv.active.any():
            full = [r for r in lens if lens[r] >= MB * B]
            for r in full:                               # at capacity: the caller retires it first
                kv.free(r)
                lens.pop(r)
            if not kv.active.any():
                continue
            kv.begin_decode_step()
            slots = kv.d_slots.numpy()
            for r in range(S):
                if r in lens:
                    pos = lens[r]
                    assert (slots[:, r] // B == kv.h_tables[:, r, pos // B]).all() and (slots[:, r] % B == pos % B).all()
                    assert kv.d_lens[r] == pos + 1 and kv.d_pos[r] == pos
                    lens[r] += 1
                else:
                    assert (slots[:, r] == -1).all() and kv.d_lens[r] == 0
        check_invariants(kv, lens)


@settings(max_examples=40, deadline=None)
@given(st.lists(st.lists(st.integers(0, 63), min_size=1, max_size=10), min_size=1, max_size=9),
       st.integers(2, 4), st.integers(11, 20), st.one_of(st.none(), st.integers(0, 63)))
def test_scheduler_results_are_independent_of_batching(prompts, rows, max_length, eos):
    kv = BatchedKVState(L * rows * MB, L, 1, 8, B, rows, MB, torch.device("cpu"))
    sch = BatchedScheduler(StubModel(), kv, max_length=max_length, eos_token_id=eos)
    ids = [sch.add_sequence(p) for p in prompts]
    sch.run()
    for i, p in zip(ids, prompts):
        assert sch.sequences[i] == expected(p, max_length, eos)
    assert len(kv.free_blocks) == kv.num_blocks and not sch.active_sequences and not sch.pending
