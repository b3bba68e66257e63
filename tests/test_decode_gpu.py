"""GPU tests of the batched decode driver against goldens produced by the reference's own GPT-2
model on CPU (tests/golden/make_golden_gpt2.py; BASELINE configs[0])."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR
from vllmini_b200.decode import BatchedKVState, GPT2Decoder, GPT2Shape, greedy_generate, synthetic_gpt2_state_dict

pytestmark = pytest.mark.gpu


def _load(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    L, H, E, V, P = (int(x) for x in g["shape"])
    return g, GPT2Shape(n_layer=L, n_head=H, n_embd=E, vocab_size=V, n_positions=P)


def _run(shape, seed, prompts, new_tokens, use_graph, max_seqs=None, dtype=torch.float16, fused=True):
    dev = torch.device("cuda:0")
    model = GPT2Decoder(shape, dev, dtype=dtype, fused_append=fused, fused_layernorm=fused)
    model.load_state_dict(synthetic_gpt2_state_dict(shape, seed=seed))
    B, MB = 16, 4
    S = max_seqs or len(prompts)
    kv = BatchedKVState(num_blocks=shape.n_layer * S * MB + 8, num_layers=shape.n_layer, num_heads=shape.n_head,
                        head_size=shape.head_size, block_size=B, max_seqs=S, max_blocks_per_seq=MB, device=dev,
                        dtype=dtype)
    # harness variant that also records the logits of row 0 at every step
    rows = []
    first = torch.zeros(S, dtype=torch.int64, device=dev)
    for i, pr in enumerate(prompts):
        lg = model.prefill(torch.tensor(list(pr), dtype=torch.int64, device=dev), kv, i)
        first[i] = torch.argmax(lg)
        if i == 0:
            rows.append(lg.float().cpu().numpy())
    cur = first
    toks = [int(cur[0])]
    if use_graph:
        model.capture_decode_graph(kv)
    for _ in range(new_tokens - 1):
        kv.begin_decode_step(attend_to_new_token=True)
        pos = torch.from_numpy(kv.positions()).to(dev)
        if use_graph:
            model.g_ids.copy_(cur)
            model.g_pos.copy_(pos)
            cur = model.replay_decode_graph().clone()
            lg = model.g_logits
        else:
            lg = model.decode_forward(cur, pos, kv)
            cur = torch.argmax(lg, dim=-1)
        rows.append(lg[0].float().cpu().numpy())
        toks.append(int(cur[0]))
    return toks, rows


@pytest.mark.parametrize("name,use_graph,batched", [("gpt2_greedy_tiny", False, False), ("gpt2_greedy_tiny", True, True),
                                                    ("gpt2_greedy_small", False, True), ("gpt2_greedy_small", True, False)])
def test_greedy_decode_matches_reference_model_golden(name, use_graph, batched):
    g, shape = _load(name)
    prompt = g["prompt"].tolist()
    rng = np.random.default_rng(5)
    prompts = [prompt]
    if batched:      # other sequences of different lengths in the same batch must not change row 0
        prompts += [rng.integers(0, shape.vocab_size, size=n).tolist() for n in (3, 16, 17, 30)]
    toks, rows = _run(shape, int(g["seed"]), prompts, len(g["tokens"]), use_graph, max_seqs=len(prompts) + 1,
                      fused=use_graph)      # graph runs use the fused launches (append+attend, add+LayerNorm),
    #                                         eager runs the reference's two calls and plain torch ops
    tol = 0.06       # fp16 weights/activations vs the reference model in fp32 on CPU
    for step, row in enumerate(rows):
        want_v, want_i = g["top_vals"][step], g["top_ids"][step]
        got_v = row[want_i]
        assert np.abs(got_v - want_v).max() < tol, f"step {step}: logits differ by {np.abs(got_v - want_v).max():.4f}"
        margin = want_v[0] - want_v[1]
        if margin > 2 * tol:
            assert toks[step] == int(g["tokens"][step]), f"step {step}: token {toks[step]} != {int(g['tokens'][step])}"


def test_decode_path_equals_prefill_path_in_fp32_weights_rounding():
    """Self-consistency on the GPU: the logits of a token decoded through the paged path equal the
    logits obtained by prefilling the whole sequence eagerly (same fp16 weights)."""
    g, shape = _load("gpt2_greedy_tiny")
    dev = torch.device("cuda:0")
    model = GPT2Decoder(shape, dev)
    model.load_state_dict(synthetic_gpt2_state_dict(shape, seed=int(g["seed"])))
    rng = np.random.default_rng(1)
    seq = rng.integers(0, shape.vocab_size, size=41).tolist()

    def fresh_kv():
        return BatchedKVState(shape.n_layer * 8, shape.n_layer, shape.n_head, shape.head_size, 16, 1, 4, dev)

    full = model.prefill(torch.tensor(seq, device=dev), fresh_kv(), 0)
    kv = fresh_kv()
    model.prefill(torch.tensor(seq[:-1], device=dev), kv, 0)
    kv.begin_decode_step()
    lg = model.decode_forward(torch.tensor(seq[-1:], device=dev), torch.from_numpy(kv.positions()).to(dev), kv)[0]
    assert (lg.float() - full.float()).abs().max().item() < 0.03


def test_greedy_generate_harness_and_block_growth():
    """Generate past several block boundaries for a ragged batch; graph and eager paths agree."""
    shape = GPT2Shape(n_layer=2, n_head=4, n_embd=256, vocab_size=503, n_positions=128)
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(2)
    prompts = [rng.integers(0, 503, size=n).tolist() for n in (1, 15, 16, 33)]
    outs = []
    for use_graph in (False, True):
        model = GPT2Decoder(shape, dev)
        model.load_state_dict(synthetic_gpt2_state_dict(shape, seed=7))
        kv = BatchedKVState(2 * 4 * 6 + 4, 2, 4, 64, 16, 4, 6, dev)
        outs.append(greedy_generate(model, kv, prompts, max_new_tokens=40, use_graph=use_graph))
        assert kv.h_lens.tolist() == [1 + 39, 15 + 39, 16 + 39, 33 + 39]
    agree = np.mean([a == b for ra, rb in zip(*outs) for a, b in zip(ra, rb)])
    assert agree > 0.9          # identical arithmetic; near-ties may flip under different split factors


@pytest.mark.parametrize("rows,hidden", [(1, 64), (5, 768), (256, 768), (33, 1024), (7, 1600), (3, 2048)])
@pytest.mark.parametrize("gather", [False, True])
def test_add_layernorm_matches_torch_fp32_reference(rows, hidden, gather):
    """ops.add_layernorm against the same arithmetic in plain torch fp32: the residual sum is rounded to fp16
    first (it is stored in fp16), LayerNorm statistics in fp32.  Tolerance: 1 fp16 ulp of the output (|y| <~ 8)
    for ln_out; x_out bit-exact."""
    from vllmini_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(rows * 10007 + hidden)
    gamma = (1 + 0.1 * torch.randn(hidden, generator=g, device=dev)).half()
    beta = (0.1 * torch.randn(hidden, generator=g, device=dev)).half()
    if gather:          # the embedding step: a = table[ids], x = table2[pos]; row stride wider than hidden
        table = torch.randn(97, hidden + 8, generator=g, device=dev).half()[:, :hidden]
        table2 = torch.randn(50, hidden, generator=g, device=dev).half()
        ia = torch.randint(0, 97, (rows,), generator=g, device=dev)
        ix = torch.randint(0, 50, (rows,), generator=g, device=dev)
        a_src, x_src, a_eff, x_eff = table, table2, table[ia], table2[ix]
        x_out = torch.empty(rows, hidden, dtype=torch.float16, device=dev)
    else:
        a_src = a_eff = torch.randn(rows, hidden, generator=g, device=dev).half()
        x_src = torch.randn(rows, hidden, generator=g, device=dev).half() * 3
        x_eff = x_src.clone()
        ia = ix = None
        x_out = x_src                                                # in place
    ln_out = torch.empty(rows, hidden, dtype=torch.float16, device=dev)
    want_x = (a_eff.float() + x_eff.float()).half()
    want_ln = torch.nn.functional.layer_norm(want_x.float(), (hidden,), gamma.float(), beta.float(), 1e-5)
    ops.add_layernorm(a_src, x_src, gamma, beta, 1e-5, x_out, ln_out, a_rows=ia, x_rows=ix)
    torch.cuda.synchronize()
    assert torch.equal(x_out, want_x)
    err = (ln_out.float() - want_ln).abs()
    assert bool((err <= 2.0 ** -10 * want_ln.abs() + 1e-3).all()), f"max err {err.max().item():.3e}"


def test_fused_layernorm_decode_matches_plain_torch_decode():
    """The decoder with fused add+LayerNorm launches (and the padded lm_head) gives the same logits as the one
    built from plain torch ops, on the same fp16 weights."""
    g, shape = _load("gpt2_greedy_small")
    dev = torch.device("cuda:0")
    sd = synthetic_gpt2_state_dict(shape, seed=3)
    rng = np.random.default_rng(4)
    prompts = [rng.integers(0, shape.vocab_size, size=n).tolist() for n in (5, 16, 23)]
    logits = []
    for fused in (False, True):
        model = GPT2Decoder(shape, dev, fused_layernorm=fused)
        model.load_state_dict(sd)
        kv = BatchedKVState(shape.n_layer * 3 * 4, shape.n_layer, shape.n_head, shape.head_size, 16, 3, 4, dev)
        for i, pr in enumerate(prompts):
            model.prefill(torch.tensor(pr, dtype=torch.int64, device=dev), kv, i)
        kv.begin_decode_step()
        ids = torch.tensor([11, 22, 33], dtype=torch.int64, device=dev)
        lg = model.decode_forward(ids, torch.from_numpy(kv.positions()).to(dev), kv)
        assert lg.shape == (3, shape.vocab_size)
        logits.append(lg.float())
    assert (logits[0] - logits[1]).abs().max().item() < 0.02


def test_argmax_rows_matches_torch():
    from vllmini_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(9)
    for rows, n, pad in ((1, 7, 0), (3, 503, 1), (256, 50257, 47), (5, 4096, 0)):
        x = torch.randn(rows, n + pad, generator=g, device=dev).half()[:, :n]
        x[0, n // 2] = x[0, n - 1] = 60000.0                  # tie: first index wins
        assert torch.equal(ops.argmax_rows(x), torch.argmax(x, dim=-1)), (rows, n)
    x = torch.randn(4, 100, generator=g, device=dev).half()
    x[2, 17] = x[2, 60] = float("nan")
    x[1, :] = float("-inf")
    x[3, :50] = float("-inf")
    assert torch.equal(ops.argmax_rows(x), torch.argmax(x, dim=-1))


def test_async_step_publishing_ring_never_tears():
    """begin_decode_step stages slots / lens / positions / new blocks through a ring of pinned buffers with no
    host-device sync; running far ahead of the GPU must still give every step its own values.  A CPU twin of
    the same state machine supplies the expectations."""
    dev = torch.device("cuda:0")
    L, S, MB, B = 3, 37, 12, 4

    def make(device):
        kv = BatchedKVState(L * S * MB, L, 2, 8, B, S, MB, device)
        rng = np.random.default_rng(0)
        for s_ in range(S):
            if s_ % 5 != 4:
                kv.add_sequence(s_, int(rng.integers(1, 9)))
        return kv

    gpu, cpu = make(dev), make(torch.device("cpu"))
    slow = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    snaps, want = [], []
    for step in range(30):
        slow.normal_()                                   # keep the GPU busy so that the host runs ahead
        gpu.begin_decode_step()
        cpu.begin_decode_step()
        snaps.append((gpu.d_slots.clone(), gpu.d_lens.clone(), gpu.d_pos.clone(), gpu.d_tables.clone()))
        want.append((cpu.d_slots.clone(), cpu.d_lens.clone(), cpu.d_pos.clone(), cpu.d_tables.clone()))
    torch.cuda.synchronize()
    for step, (got, exp) in enumerate(zip(snaps, want)):
        for g_, e_ in zip(got, exp):
            assert torch.equal(g_.cpu(), e_), f"step {step}"


def test_graphed_bucket_padded_prefill_equals_eager_prefill():
    """prefill(graph_bucket=16) pads the prompt, replays a CUDA graph per padded length, and must produce the same
    last-position logits and the same cache contents as the eager prefill."""
    g, shape = _load("gpt2_greedy_tiny")
    dev = torch.device("cuda:0")
    model = GPT2Decoder(shape, dev)
    model.load_state_dict(synthetic_gpt2_state_dict(shape, seed=5))
    rng = np.random.default_rng(8)

    def fresh():
        return BatchedKVState(shape.n_layer * 2 * 4, shape.n_layer, shape.n_head, shape.head_size, 16, 2, 4, dev)

    kv_e, kv_g = fresh(), fresh()
    for row, T in ((0, 37), (1, 45)):                    # both pad to 48: the second replay reuses the graph
        ids = torch.tensor(rng.integers(0, shape.vocab_size, size=T), device=dev)
        le = model.prefill(ids, kv_e, row)
        lg = model.prefill(ids, kv_g, row, graph_bucket=16).clone()
        assert (le.float() - lg.float()).abs().max().item() < 2e-2
        assert int(le.argmax()) == int(lg.argmax())
    assert len(model._prefill_graphs) == 1
    torch.cuda.synchronize()
    assert torch.equal(kv_e.key_cache, kv_g.key_cache) and torch.equal(kv_e.value_cache, kv_g.value_cache)
    assert np.array_equal(kv_e.h_tables, kv_g.h_tables)


@pytest.mark.parametrize("name,graph_bucket", [("gpt2_greedy_tiny", 0), ("gpt2_greedy_tiny", 16), ("gpt2_greedy_small", 0)])
def test_prefill_through_the_paged_path(name, graph_bucket):
    """prefill_attention="paged" runs the prompt's causal attention through
    paged_attention_v1 itself -- one row per prompt token, all rows sharing the sequence's block table, row t with
    seq_len t + 1, K/V read back from the cache reshape_and_cache just filled.  It must agree with the library-kernel
    prefill (same caches bit for bit in layer 0, last-position logits within fp16 noise, same greedy token), with
    the reference model's golden logits, and the decode steps that follow must be unaffected."""
    g, shape = _load(name)
    dev = torch.device("cuda:0")
    sd = synthetic_gpt2_state_dict(shape, seed=int(g["seed"]))
    prompt = g["prompt"].tolist()

    def run(mode):
        model = GPT2Decoder(shape, dev, prefill_attention=mode)
        model.load_state_dict(sd)
        kv = BatchedKVState(shape.n_layer * 2 * 4, shape.n_layer, shape.n_head, shape.head_size, 16, 2, 4, dev)
        other = np.random.default_rng(1).integers(0, shape.vocab_size, size=21).tolist()
        model.prefill(torch.tensor(other, dtype=torch.int64, device=dev), kv, 0, graph_bucket=graph_bucket)
        lg = model.prefill(torch.tensor(prompt, dtype=torch.int64, device=dev), kv, 1, graph_bucket=graph_bucket).clone()
        toks = [int(lg.argmax())]
        cur = torch.zeros(2, dtype=torch.int64, device=dev)
        cur[1] = toks[0]
        for _ in range(3):
            kv.begin_decode_step(attend_to_new_token=True)
            cur = model.greedy(model.decode_forward(cur, kv.d_pos, kv))
            toks.append(int(cur[1]))
        torch.cuda.synchronize()
        return lg.float().cpu().numpy(), toks, kv

    lg_s, toks_s, kv_s = run("sdpa")
    lg_p, toks_p, kv_p = run("paged")
    assert np.abs(lg_s - lg_p).max() < 2e-2
    want_v, want_i = g["top_vals"][0], g["top_ids"][0]               # the reference model's own prefill logits
    assert np.abs(lg_p[want_i] - want_v).max() < 0.06
    if want_v[0] - want_v[1] > 0.12:
        assert toks_p[0] == int(g["tokens"][0]) == toks_s[0]
    blocks0 = torch.tensor(sorted(int(b) for b in kv_s.h_tables[0][kv_s.h_tables[0] >= 0]), device=dev)
    assert np.array_equal(kv_s.h_tables, kv_p.h_tables)
    assert torch.equal(kv_s.key_cache[blocks0], kv_p.key_cache[blocks0])      # layer 0: embeddings only
    assert (kv_s.key_cache.float() - kv_p.key_cache.float()).abs().max().item() < 5e-2
