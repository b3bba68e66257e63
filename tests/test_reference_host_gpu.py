"""The UNMODIFIED vllmini host running on the drop-in, on the GPU, against the same host on the reference's own
kernels (north star: "drops in behind the existing paged_attention_ext Python binding so vllmini.kv_cache,
block_manager and scheduler are untouched").

The host files -- vllmini/model/gpt2.py (call sites :81, :94; import :5), kv_cache.py, block_manager.py,
scheduler.py, helpers/generate_triangular_mask.py and the reference's kernel test tests/kernels/paged_attention.py
-- are byte-for-byte copies staged by oracle/stage_ref_host.py into the git-ignored oracle/_ref/vllmini_host/
(they travel to the GPU box like oracle/_ref/*.so); their SHA-256 digests are checked against the committed
oracle/ref_host_manifest.json before anything runs.  The ONLY thing that differs between two runs is what
`import paged_attention_cuda` resolves to while the host package is imported:

    "ref"    oracle/_ref/paged_attention_cuda*.so   the reference's kernels compiled for sm_100 (build_ref.py)
    "new"    vllmini_b200/dropin/paged_attention_cuda   this project (C ABI -> sm_100a kernels)
    "mixed"  the reference's paged_attention_v1 + this project's cache_ops (isolates reshape_and_cache inside the
             real stack: identical attention arithmetic => caches and logits must be BIT-identical to "ref")

Tensors reaching the kernels are the ones the reference's own KVCache / BlockManager.decode_step /
GPT2Attention.forward build: [1, max_blocks] int32 tables per (sequence, layer), [1] int64 slot tensors, q/k/v views
of the fused qkv projection (stride(0) = 768 at T == 1), max_seq_len = max_blocks_per_seq * block_size, and the
scheduler's seq_lens = L convention (the new token does not attend to itself, scheduler.py:96-101).
"""
import contextlib
import importlib
import importlib.util
import io
import os
import sys
import types
import unittest

import numpy as np
import pytest
import torch

from helpers import ROOT, load_ref_extension
from oracle import stage_ref_host

pytestmark = pytest.mark.gpu

HOST_MODULES = ("vllmini.model.gpt2", "vllmini.model.helpers.generate_triangular_mask", "vllmini.kv_cache",
                "vllmini.block_manager", "vllmini.scheduler")


def _require_staged():
    if not stage_ref_host.verify_staged():
        if stage_ref_host.reference_available():
            stage_ref_host.stage()
        assert stage_ref_host.verify_staged(), (
            "oracle/_ref/vllmini_host is missing or differs from oracle/ref_host_manifest.json; "
            "run `python oracle/stage_ref_host.py` where /root/reference exists")
    return stage_ref_host.staged_dir()


def _purge():
    for k in [k for k in sys.modules if k == "vllmini" or k.startswith("vllmini.") or k == "paged_attention_cuda"
              or k.startswith("paged_attention_cuda.")]:
        del sys.modules[k]


def _binding(kind, ref_ext):
    """The module object `import paged_attention_cuda` will resolve to."""
    import vllmini_b200
    _purge()
    vllmini_b200.install_dropin()
    new = importlib.import_module("paged_attention_cuda")          # vllmini_b200/dropin/paged_attention_cuda
    assert os.path.dirname(new.__file__).startswith(os.path.join(ROOT, "vllmini_b200", "dropin"))
    if kind == "new":
        return new
    if kind == "ref":
        return ref_ext
    m = types.ModuleType("paged_attention_cuda")
    if kind == "mixed":              # reference attention, this project's cache ops
        m.paged_attention_v1, m.cache_ops = ref_ext.paged_attention_v1, new.cache_ops
    else:
        raise KeyError(kind)
    return m


def import_host(kind, ref_ext):
    """Import the staged, unmodified host package with `paged_attention_cuda` resolving to `kind`; returns the
    module objects.  sys.modules is purged before and after, so every call yields fresh, independently bound
    modules (gpt2.py binds the two functions at import time, gpt2.py:5)."""
    staged = _require_staged()
    binding = _binding(kind, ref_ext)
    _purge()
    sys.modules["paged_attention_cuda"] = binding
    sys.path.insert(0, staged)
    try:
        mods = {name.rsplit(".", 1)[-1]: importlib.import_module(name) for name in HOST_MODULES}
        assert all(os.path.abspath(m.__file__).startswith(staged) for m in mods.values())
        assert mods["gpt2"].paged_attention_v1 is binding.paged_attention_v1
        assert mods["gpt2"].cache_ops is binding.cache_ops
    finally:
        sys.path.remove(staged)
        _purge()
    return mods


@pytest.fixture(scope="module")
def ref_ext():
    ext = load_ref_extension()
    if ext is None:
        pytest.skip("oracle/_ref (the reference's kernels for sm_100) is not built")
    return ext


@pytest.fixture(autouse=True)
def _quiet_and_clean():
    yield
    _purge()


def _silenced(fn, *a, **kw):
    """The reference host print()s on every step (scheduler.py:57-69, block_manager.py:49)."""
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


# ---------------------------------------------------------------------------------------------------------
# 1. the reference's own kernel test, unmodified (SURVEY.md section 7 "minimum slice")
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["new", "ref"])
def test_reference_kernel_test_runs_unmodified(kind, ref_ext):
    """vllmini/tests/kernels/paged_attention.py:84-139 executed as it is (unittest), with paged_attention_cuda ->
    the drop-in ("new"); and -> the reference's own kernels ("ref") to show the test itself is sound on B200."""
    from vllmini_b200 import _native
    staged = _require_staged()
    binding = _binding(kind, ref_ext)
    _purge()
    sys.modules["paged_attention_cuda"] = binding
    path = os.path.join(staged, "vllmini", "tests", "kernels", "paged_attention.py")
    spec = importlib.util.spec_from_file_location("reference_kernel_test", path)
    mod = importlib.util.module_from_spec(spec)
    before = _native.launch_count()
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        spec.loader.exec_module(mod)
        assert mod.paged_attention_v1 is binding.paged_attention_v1 and mod.cache_ops is binding.cache_ops
        suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
        assert suite.countTestCases() == 1
        result = unittest.TextTestRunner(stream=io.StringIO(), verbosity=0).run(suite)
    assert result.wasSuccessful(), (result.failures, result.errors)
    launched = _native.launch_count() - before
    assert launched == (2 if kind == "new" else 0), launched    # reshape_and_cache + paged_attention_v1, ours or not


# ---------------------------------------------------------------------------------------------------------
# 2. KVCache + BlockManager + GPT2LMHeadModel: prefill + decode steps, stack vs stack
# ---------------------------------------------------------------------------------------------------------
def _make_model(mods, layers, seed):
    from transformers import GPT2Config
    from vllmini_b200.decode import GPT2Shape, synthetic_gpt2_state_dict
    shape = GPT2Shape(n_layer=layers)                                    # GPT-2 small widths: 12 heads x 64, vocab 50257
    cfg = GPT2Config(n_layer=layers)
    model = mods["gpt2"].GPT2LMHeadModel(cfg)
    missing, unexpected = model.load_state_dict(synthetic_gpt2_state_dict(shape, seed=seed), strict=False)
    assert not missing and not unexpected
    return model.to(torch.float16).cuda().eval(), cfg


def _run_stack(mods, layers, prompts, forced_tokens, max_blocks_per_seq, seed=3):
    """Drive the reference classes the way Scheduler.add_sequence / run do (scheduler.py:22-53, 75-101), but with a
    fixed token stream so that two runs see identical inputs: prefill each prompt, then decode round-robin.
    Returns per-step logits, the final caches and the block tables."""
    torch.manual_seed(0)
    model, cfg = _make_model(mods, layers, seed)
    BlockManager = mods["block_manager"].BlockManager
    mask_fn = mods["generate_triangular_mask"].generate_triangular_mask
    bm = BlockManager(num_blocks=layers * len(prompts) * max_blocks_per_seq + 5, block_size=16, num_heads=12,
                      head_size=64, max_blocks_per_seq=max_blocks_per_seq)
    kc, vc = bm.kv_cache.key_cache, bm.kv_cache.value_cache
    logits_log, lengths = [], {}
    with torch.no_grad():
        for sid, prompt in enumerate(prompts):
            ids = torch.tensor([prompt], device="cuda")
            T = ids.size(1)
            _, _, slot_mappings, tables = bm.allocate_for_prefill(sid, layers, T)
            logits, _ = model(input_ids=ids, position_ids=torch.arange(T, device="cuda"),
                              attention_mask=mask_fn(1, 12, T), use_cache=True, key_cache=kc, value_cache=vc,
                              slot_mappings=slot_mappings, block_tables=tables)
            logits_log.append(logits[0, -1].float().cpu().numpy())
            lengths[sid] = T
        cache_after_prefill = (kc.clone(), vc.clone())
        for step, toks in enumerate(forced_tokens):
            for sid, tok in enumerate(toks):
                L = lengths[sid]
                tables, new_slots = _silenced(bm.decode_step, sid, 1)
                logits, _ = model(input_ids=torch.tensor([[tok]], device="cuda"),
                                  position_ids=torch.tensor([L], device="cuda"), attention_mask=None, use_cache=True,
                                  is_prefill=False, key_cache=kc, value_cache=vc, slot_mappings=new_slots,
                                  block_tables=tables, seq_lens=torch.tensor([L], dtype=torch.int32, device="cuda"),
                                  max_seq_len=bm.kv_cache.max_blocks_per_seq * bm.block_size)
                logits_log.append(logits[0, -1].float().cpu().numpy())
                lengths[sid] = L + 1
    torch.cuda.synchronize()
    tables = {sid: [t.cpu().numpy().copy() for t in bm.kv_cache.paged_attention_block_tables[sid]] for sid in lengths}
    return dict(logits=logits_log, kc=kc, vc=vc, prefill=cache_after_prefill, tables=tables, lengths=lengths)


def _layer_blocks(run, layer):
    return sorted({int(b) for t in run["tables"].values() for b in t[layer][0] if b >= 0})


@pytest.mark.parametrize("layers,steps,max_blocks", [(12, 24, 4), (4, 70, 8)])
def test_reference_stack_on_dropin_matches_reference_stack_on_reference_kernels(layers, steps, max_blocks, ref_ext):
    """allocate_for_prefill + prefill + >= 24 decode_step/forward iterations of the reference's KVCache, BlockManager
    and GPT2LMHeadModel (GPT-2 small widths; 12 layers, and a 4-layer model run for 70 steps across 5 block
    boundaries), three times: on the reference's kernels, on the drop-in, and on the mixed binding."""
    from vllmini_b200 import _native
    rng = np.random.default_rng(layers)
    prompts = [rng.integers(0, 50257, size=n).tolist() for n in (5, 11)]
    forced = rng.integers(0, 50257, size=(steps, len(prompts))).tolist()
    runs = {}
    for kind in ("ref", "ref2", "new", "mixed"):
        before = _native.launch_count()
        mods = import_host("ref" if kind == "ref2" else kind, ref_ext)
        runs[kind] = _run_stack(mods, layers, prompts, forced, max_blocks)
        launched = _native.launch_count() - before
        per_forward = {"ref": 0, "ref2": 0, "new": 2, "mixed": 1}[kind]        # our launches per (layer, decode forward)
        n_fwd = steps * len(prompts)
        want = per_forward * layers * n_fwd + (layers * len(prompts) if per_forward else 0)   # + prefill scatters
        assert launched == want, (kind, launched, want)
    ref, new, mixed = runs["ref"], runs["new"], runs["mixed"]
    # same bookkeeping in every run (it is the same untouched host code fed the same tokens)
    for sid in ref["tables"]:
        for l in range(layers):
            assert np.array_equal(ref["tables"][sid][l], new["tables"][sid][l])
    assert ref["lengths"] == new["lengths"] == {0: 5 + steps, 1: 11 + steps}
    deterministic = all(np.array_equal(a, b) for a, b in zip(ref["logits"], runs["ref2"]["logits"]))

    # ---- reshape_and_cache inside the real stack: bit-exact ----
    # after prefill nothing has gone through attention kernels yet: whole caches identical in all runs
    for a, b in zip(ref["prefill"], new["prefill"]):
        assert torch.equal(a, b), "caches differ after prefill"
    # layer 0's K/V depend only on the embeddings: bit-exact after all decode steps too
    blk0 = torch.tensor(_layer_blocks(ref, 0), device="cuda")
    assert torch.equal(ref["kc"][blk0], new["kc"][blk0]) and torch.equal(ref["vc"][blk0], new["vc"][blk0])
    if deterministic:
        # mixed = reference attention + our cache ops: every byte of both caches and every logit must match
        assert torch.equal(ref["kc"], mixed["kc"]) and torch.equal(ref["vc"], mixed["vc"]), \
            "reshape_and_cache differs from the reference inside the stack"
        assert all(np.array_equal(a, b) for a, b in zip(ref["logits"], mixed["logits"]))

    # ---- paged_attention_v1 inside the real stack: SURVEY.md 8c tolerance, carried to the logits ----
    # deeper layers see each other's fp16-rounding differences; caches stay close, logits within 1e-2
    # (|logit| ~ 1; the attention outputs themselves differ by <= 2e-3 + 1e-2 |ref|; measured on B200: 3.9e-3 over
    # 12 layers x 24 steps, 3.2e-3 over 4 layers x 70 steps)
    assert (ref["kc"].float() - new["kc"].float()).abs().max().item() < 5e-2
    assert (ref["vc"].float() - new["vc"].float()).abs().max().item() < 5e-2
    worst = 0.0
    for i, (a, b) in enumerate(zip(ref["logits"], new["logits"])):
        d = float(np.abs(a - b).max())
        worst = max(worst, d)
        assert d < 1e-2, f"forward {i}: logits differ by {d:.4f}"
        if i < len(prompts) and deterministic:
            assert np.array_equal(a, b), "prefill logits must be identical (no paged kernel involved)"
        top2 = np.sort(a)[-2:]
        if top2[1] - top2[0] > 2e-2:
            assert int(a.argmax()) == int(b.argmax()), f"forward {i}: greedy token differs"
    print(f"[reference host] {layers} layers x {steps} steps: max |logit(new) - logit(ref)| = {worst:.2e}; "
          f"cuBLAS run-to-run deterministic: {deterministic}")


# ---------------------------------------------------------------------------------------------------------
# 3. the reference Scheduler end to end (scheduler.py:22-53 add_sequence, :55-115 run)
# ---------------------------------------------------------------------------------------------------------
def _run_scheduler(mods, prompts, max_length, seed):
    model, cfg = _make_model(mods, 12, seed=5)
    bm = mods["block_manager"].BlockManager(num_blocks=1000, block_size=16, num_heads=12, head_size=64,
                                            max_blocks_per_seq=4)          # server.py:37-41
    sch = mods["scheduler"].Scheduler(model, bm, max_length=max_length)
    torch.manual_seed(seed)
    ids = []
    with torch.no_grad():
        for p in prompts:
            ids.append(_silenced(sch.add_sequence, torch.tensor([p], device="cuda")))
        first_logits = {i: sch.last_logits[i].float().cpu().numpy() for i in ids}
        _silenced(sch.run)
    torch.cuda.synchronize()
    return sch, bm, ids, first_logits


def test_reference_scheduler_runs_to_completion_on_the_dropin(ref_ext):
    """Scheduler.add_sequence x 3 + Scheduler.run() with server.py's pool constants and max_length 20, untouched,
    on the drop-in: every sequence reaches max_length (or eos), every block returns to the free list.  The same
    run on the reference's kernels with the same sampler seed: prefill logits identical; generated tokens are
    compared and must agree at least up to the first sampled token (top-k multinomial on logits that differ at
    fp16 rounding level may legitimately diverge later)."""
    from vllmini_b200 import _native
    rng = np.random.default_rng(11)
    prompts = [rng.integers(0, 50257, size=n).tolist() for n in (4, 9, 14)]
    before = _native.launch_count()
    sch, bm, ids, first_new = _run_scheduler(import_host("new", ref_ext), prompts, 20, seed=1234)
    launched = _native.launch_count() - before
    assert not sch.active_sequences and sch.queue.empty()
    assert len(bm.kv_cache.free_blocks) == 1000 and not bm.kv_cache.allocated_blocks
    gen_new = [sch.sequences[i][0].tolist() for i in ids]
    eos = sch.model.config.eos_token_id
    for p, g in zip(prompts, gen_new):
        assert g[:len(p)] == p and (len(g) == 20 or g[-1] == eos), (len(p), len(g))
    decode_forwards = sum(len(g) - len(p) for p, g in zip(prompts, gen_new))
    assert launched == 12 * len(prompts) + 2 * 12 * decode_forwards      # prefill scatters + 2 launches/layer/step

    sch_r, _bm_r, ids_r, first_ref = _run_scheduler(import_host("ref", ref_ext), prompts, 20, seed=1234)
    gen_ref = [sch_r.sequences[i][0].tolist() for i in ids_r]
    for a, b in zip(ids, ids_r):
        assert np.array_equal(first_new[a], first_ref[b]), "prefill logits differ between the two stacks"
    agree = sum(int(x == y) for g, r in zip(gen_new, gen_ref) for x, y in zip(g, r))
    total = sum(min(len(g), len(r)) for g, r in zip(gen_new, gen_ref))
    for p, g, r in zip(prompts, gen_new, gen_ref):
        assert g[len(p)] == r[len(p)], "first sampled token differs although the logits are identical"
    print(f"[reference scheduler] {agree}/{total} generated positions agree between the drop-in and the reference "
          f"kernels under the same sampler seed")
