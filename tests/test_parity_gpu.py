"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the product
path: drop-in package -> vllmini_b200.ops -> C ABI -> hand-written sm_100a kernels.  The oracle
(oracle/) and the reference's own kernels (oracle/_ref) are only the checkers.

Tolerances:
  reshape_and_cache : bit-exact.
  paged_attention_v1: |new - ref_kernel| <= 2e-3 + 1e-2*|ref_kernel| elementwise, and
                      max|new - fp64_truth| <= max(1e-3 (+1 fp16 ulp), max|ref_kernel - fp64_truth|);
                      plus the reference test's own atol=1e-2 against eager fp16 attention.
"""
import glob
import os

import numpy as np
import pytest
import torch

from helpers import (ATOL_VS_REF, GOLDEN_DIR, PARITY_CASES, RTOL_VS_REF, assert_close_to_truth, build_case,
                     load_ref_extension, run_cuda_attention, run_ref_attention, to_cuda)
from oracle import c_oracle
from oracle import paged_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_ext():
    return load_ref_extension()


@pytest.fixture(autouse=True)
def _native_path_is_the_one_running():
    from vllmini_b200 import _native
    before = _native.launch_count()
    yield
    assert _native.launch_count() > before, "test did not launch any vllmini_b200 kernel"


# ------------------------------------------------------------------------------ reshape_and_cache
@pytest.mark.parametrize("B,H,D,T,strided", [(16, 12, 64, 3, False), (16, 12, 64, 256, True), (8, 20, 64, 77, True),
                                             (32, 12, 64, 1000, True), (16, 4, 128, 33, False),
                                             # the other head sizes, decode-sized (warp kernel) and prefill-sized (tile)
                                             (8, 3, 80, 40, True), (16, 3, 80, 700, True), (16, 2, 96, 513, False),
                                             (8, 2, 112, 600, True), (32, 2, 192, 90, True), (16, 2, 256, 1030, True)])
def test_reshape_and_cache_bit_exact(B, H, D, T, strided):
    from vllmini_b200 import ops
    rng = np.random.default_rng(T)
    NB = (T + B - 1) // B + 9
    width = 3 * H * D if strided else H * D
    qkv = rng.standard_normal((T, width)).astype(np.float16)
    # poison some values with NaN/Inf bit patterns: a copy must preserve bits, not values
    qkv.view(np.uint16)[::97] = 0x7e01
    k_np = qkv[:, width - 2 * H * D: width - H * D].reshape(T, H, D) if strided else qkv.reshape(T, H, D)
    v_np = qkv[:, width - H * D:].reshape(T, H, D) if strided else qkv[::-1].reshape(T, H, D)
    slots = rng.permutation(NB * B)[:T].astype(np.int64)
    slots[rng.integers(0, T, size=max(T // 10, 1))] = -1
    kc0 = rng.standard_normal((NB, H, D // 8, B, 8)).astype(np.float16)
    vc0 = rng.standard_normal((NB, H, D, B)).astype(np.float16)
    kc_want, vc_want = kc0.copy(), vc0.copy()
    c_oracle().reshape_and_cache(k_np, v_np, kc_want, vc_want, slots)

    dev = torch.device("cuda:0")
    qkv_t = torch.from_numpy(qkv).to(dev)
    if strided:
        k_t = qkv_t[:, width - 2 * H * D: width - H * D].view(T, H, D)
        v_t = qkv_t[:, width - H * D:].view(T, H, D)
        assert k_t.stride(0) == width
    else:
        k_t = qkv_t.view(T, H, D)
        v_t = torch.from_numpy(np.ascontiguousarray(v_np)).to(dev)
    kc_t, vc_t = torch.from_numpy(kc0).to(dev), torch.from_numpy(vc0).to(dev)
    ops.reshape_and_cache(k_t, v_t, kc_t, vc_t, torch.from_numpy(slots).to(dev), "auto", 1.0)
    torch.cuda.synchronize()
    assert np.array_equal(kc_t.cpu().numpy().view(np.uint16), kc_want.view(np.uint16))
    assert np.array_equal(vc_t.cpu().numpy().view(np.uint16), vc_want.view(np.uint16))


def test_reshape_and_cache_generic_path_unaligned_rows():
    """Rows that are not 16-byte aligned take the scalar kernel; still bit-exact."""
    from vllmini_b200 import ops
    rng = np.random.default_rng(5)
    T, H, D, B, NB = 9, 3, 64, 16, 4
    buf = rng.standard_normal((T, H * D + 3)).astype(np.float16)
    dev = torch.device("cuda:0")
    buf_t = torch.from_numpy(buf).to(dev)
    k_t = buf_t[:, 1:1 + H * D].view(T, H, D)       # offset 2 bytes, stride 195 elements
    v_t = buf_t[:, 3:3 + H * D].view(T, H, D)
    slots = rng.permutation(NB * B)[:T].astype(np.int64)
    kc_want = np.zeros((NB, H, D // 8, B, 8), dtype=np.float16)
    vc_want = np.zeros((NB, H, D, B), dtype=np.float16)
    po.reshape_and_cache(buf[:, 1:1 + H * D].reshape(T, H, D), buf[:, 3:3 + H * D].reshape(T, H, D), kc_want, vc_want, slots)
    kc_t = torch.zeros((NB, H, D // 8, B, 8), dtype=torch.float16, device=dev)
    vc_t = torch.zeros((NB, H, D, B), dtype=torch.float16, device=dev)
    ops.reshape_and_cache(k_t, v_t, kc_t, vc_t, torch.from_numpy(slots).to(dev), "auto", 1.0)
    torch.cuda.synchronize()
    assert np.array_equal(kc_t.cpu().numpy().view(np.uint16), kc_want.view(np.uint16))
    assert np.array_equal(vc_t.cpu().numpy().view(np.uint16), vc_want.view(np.uint16))


def _slot_pattern(kind, T, B, NB, rng):
    """Slot mappings that exercise the tile kernel's classification: whole runs (prefill), runs starting inside a
    block, short runs glued together (tiles just above / below the half-consecutive threshold), scattered, padding."""
    if kind == "consecutive":
        return np.arange(T, dtype=np.int64) + 3 * B
    if kind == "offset_run":
        return np.arange(T, dtype=np.int64) + 5 * B + 7
    if kind == "scattered":
        return rng.permutation(NB * B)[:T].astype(np.int64)
    if kind == "short_runs":            # runs of 2..12 consecutive slots at random block offsets, in random order
        free = rng.permutation(NB)
        out, b = [], 0
        while len(out) < T:
            n = int(rng.integers(2, 13))
            start = int(free[b]) * B + int(rng.integers(0, B - n + 1)) if n <= B else int(free[b]) * B
            out.extend(range(start, start + min(n, B)))
            b += 1
        return np.asarray(out[:T], dtype=np.int64)
    if kind == "mixed":                 # first half one run, second half scattered over the remaining blocks, some padding
        half = T // 2
        run = np.arange(half, dtype=np.int64)
        first_free = (half + B - 1) // B
        rest = rng.permutation((NB - first_free) * B)[: T - half].astype(np.int64) + first_free * B
        s = np.concatenate([run, rest])
        s[rng.integers(0, T, size=max(T // 15, 1))] = -1
        return s
    raise KeyError(kind)


@pytest.mark.parametrize("kind", ["consecutive", "offset_run", "scattered", "short_runs", "mixed"])
@pytest.mark.parametrize("T,B,H,D", [(300, 16, 12, 64),      # warp kernel
                                     (513, 16, 12, 64), (1500, 8, 3, 80), (2048, 32, 2, 128),   # 16-token tiles, 8 warps
                                     (2049, 16, 12, 64), (5000, 16, 2, 256)])                   # 32-token tiles
def test_reshape_and_cache_slot_patterns_bit_exact(kind, T, B, H, D, ref_ext):
    """Every dispatch regime of reshape_and_cache x every slot pattern its tile classification distinguishes,
    bit-exact against the C oracle and against the reference's own kernel (cache_kernels.cu:152-207)."""
    from vllmini_b200 import ops
    rng = np.random.default_rng(T * 7 + len(kind))
    NB = max(2 * ((T + B - 1) // B), T // 2) + 16       # short_runs spends one block per 2..12 tokens
    qkv = rng.standard_normal((T, 3 * H * D)).astype(np.float16)
    qkv.view(np.uint16)[::89] = 0x7e01
    slots = _slot_pattern(kind, T, B, NB, rng)
    live = slots[slots >= 0]
    assert len(np.unique(live)) == len(live) and live.max() < NB * B
    k_np, v_np = qkv[:, H * D:2 * H * D].reshape(T, H, D), qkv[:, 2 * H * D:].reshape(T, H, D)
    kc0 = rng.standard_normal((NB, H, D // 8, B, 8)).astype(np.float16)
    vc0 = rng.standard_normal((NB, H, D, B)).astype(np.float16)
    kc_want, vc_want = kc0.copy(), vc0.copy()
    c_oracle().reshape_and_cache(k_np, v_np, kc_want, vc_want, slots)
    dev = torch.device("cuda:0")
    qkv_t = torch.from_numpy(qkv).to(dev)
    k_t, v_t = qkv_t[:, H * D:2 * H * D].view(T, H, D), qkv_t[:, 2 * H * D:].view(T, H, D)
    slots_t = torch.from_numpy(slots).to(dev)
    kc_t, vc_t = torch.from_numpy(kc0).to(dev), torch.from_numpy(vc0).to(dev)
    ops.reshape_and_cache(k_t, v_t, kc_t, vc_t, slots_t, "auto", 1.0)
    torch.cuda.synchronize()
    assert np.array_equal(kc_t.cpu().numpy().view(np.uint16), kc_want.view(np.uint16))
    assert np.array_equal(vc_t.cpu().numpy().view(np.uint16), vc_want.view(np.uint16))
    if ref_ext is not None:
        kc_r, vc_r = torch.from_numpy(kc0).to(dev), torch.from_numpy(vc0).to(dev)
        ref_ext.cache_ops.reshape_and_cache(k_t, v_t, kc_r, vc_r, slots_t, "auto", 1.0)
        torch.cuda.synchronize()
        assert torch.equal(kc_t.view(torch.int16), kc_r.view(torch.int16))
        assert torch.equal(vc_t.view(torch.int16), vc_r.view(torch.int16))


# ------------------------------------------------------------------------------ paged_attention_v1
@pytest.mark.parametrize("name", [c[0] for c in PARITY_CASES])
def test_paged_attention_vs_oracle_and_reference_kernel(name, ref_ext):
    c = build_case(name)
    t = to_cuda(c)
    got = run_cuda_attention(c, t)
    assert np.isfinite(got.astype(np.float32)).all()
    truth = c_oracle().paged_attention_v1(c["q"], c["kc"], c["vc"], c["scale"], c["bt"], c["sl"], mode="exact")
    truth64 = po.paged_attention_v1_exact(c["q"], c["kc"], c["vc"], c["scale"], c["bt"], c["sl"]) \
        if c["S"] * int(c["sl"].max()) <= 4096 else truth.astype(np.float64)
    err_new = assert_close_to_truth(got, truth64, what=f"{name}: cuda vs fp64 oracle")
    for s in np.nonzero(c["sl"] == 0)[0]:
        assert not got[s].any(), "seq_len == 0 row must be zeros"
    if ref_ext is not None:
        ref = run_ref_attention(ref_ext, c, t).astype(np.float64)
        diff = np.abs(got.astype(np.float64) - ref)
        assert (diff <= ATOL_VS_REF + RTOL_VS_REF * np.abs(ref)).all(), f"{name}: vs reference kernel {diff.max():.3e}"
        err_ref = float(np.abs(ref - truth64).max())
        assert err_new <= max(1e-3 + float(np.abs(truth64).max()) * 2.0 ** -10, err_ref), \
            f"{name}: new kernel further from truth ({err_new:.3e}) than the reference kernel ({err_ref:.3e})"


@pytest.mark.parametrize("seed", range(24))
def test_paged_attention_random_shapes(seed, ref_ext):
    """Seeded random problems over the reference launcher's whole dispatch table: head size, block size, MHA /
    grouped-query head counts, batch, ragged lengths (empty sequences included), spare table columns, a max_seq_len
    hint anywhere between too small and too large (it may only change the split factor), a padded q stride --
    against the C oracle (fp64 accumulation) and the reference's own kernel."""
    from vllmini_b200 import ops
    rng = np.random.default_rng(9000 + seed)
    D = int(rng.choice([64, 64, 64, 80, 96, 112, 128, 192, 256]))
    B = int(rng.choice([8, 16, 32]))
    KVH = int(rng.choice([1, 2, 3, 4, 12]))
    G = int(rng.choice([1, 1, 2, 4]))
    H = KVH * G
    S = int(rng.choice([1, 2, 3, 7, 19, 40]))
    hi = int(rng.choice([40, 300, 900])) if S > 3 else int(rng.choice([300, 2500, 5000]))
    lens = rng.integers(0, hi + 1, size=S).astype(np.int32)
    lens[rng.integers(0, S)] = hi                        # at least one sequence of the full length
    spare = int(rng.integers(0, 4))
    nblk = int((int(lens.max()) + B - 1) // B)
    q, kc, vc, bt, sl = po.make_case(S, KVH, D, B, lens, num_blocks=int(((lens.astype(np.int64) + B - 1) // B).sum()) + 5,
                                     seed=100 + seed, max_blocks_per_seq=max(nblk, 1) + spare)
    qpad = int(rng.choice([0, 8, 2 * H * D]))
    qfull = rng.standard_normal((S, H * D + qpad)).astype(np.float16)
    q = qfull[:, : H * D].reshape(S, H, D)
    scale = float(D) ** -0.5
    # oracle on the MHA-expanded cache (query head h reads kv head h // G)
    kc_x, vc_x = np.repeat(kc, G, axis=1), np.repeat(vc, G, axis=1)
    truth = c_oracle().paged_attention_v1(np.ascontiguousarray(q), kc_x, vc_x, scale, bt, sl, mode="exact").astype(np.float64)
    dev = torch.device("cuda:0")
    q_t = torch.from_numpy(qfull).to(dev)[:, : H * D].view(S, H, D)
    kc_t, vc_t, bt_t, sl_t = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (kc, vc, bt, sl))
    msl = int(rng.choice([1, max(int(lens.max()) // 2, 1), max(int(lens.max()), 1), bt.shape[1] * B, 3 * bt.shape[1] * B]))
    out = torch.empty(S, H, D, dtype=torch.float16, device=dev)
    ops.paged_attention_v1(out, q_t, kc_t, vc_t, KVH, scale, bt_t, sl_t, B, msl, None, "auto", 1.0, 0, 0, 1, 1, 0)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    what = f"seed {seed}: S={S} H={H}/{KVH} D={D} B={B} lens<= {hi} max_seq_len={msl} qpad={qpad}"
    err_new = assert_close_to_truth(got, truth, atol=1.5e-3, what=what)
    for s_ in np.nonzero(sl == 0)[0]:
        assert not got[s_].any(), what + ": seq_len == 0 row must be zeros"
    if ref_ext is not None:
        ref = torch.empty_like(out)
        ref_ext.paged_attention_v1(ref, q_t, kc_t, vc_t, KVH, scale, bt_t, sl_t, B, int(bt.shape[1]) * B, None, "auto",
                                   1.0, 0, 0, 1, 1, 0)
        torch.cuda.synchronize()
        refn = ref.cpu().numpy().astype(np.float64)
        diff = np.abs(got.astype(np.float64) - refn)
        assert (diff <= ATOL_VS_REF + RTOL_VS_REF * np.abs(refn)).all(), what + f": vs reference kernel {diff.max():.3e}"
        assert err_new <= max(1.5e-3 + float(np.abs(truth).max()) * 2.0 ** -10, float(np.abs(refn - truth).max())), what


@pytest.mark.parametrize("max_seq_len", [64, 512, 1280, 2048, 4608])
def test_split_kv_factor_does_not_change_the_result(max_seq_len):
    """3 sequences x 12 heads leave SMs idle, so the launcher splits the context; the factor follows the caller's
    max_seq_len hint (1 / 2 / 5 / 8 / 9 splits for the values here).  Every factor must reproduce the fp64 truth --
    including hints SMALLER than the real lengths: max_seq_len only sizes the split, the last split runs to
    seq_len (the reference's v1 has no such hint dependence either).  The second call checks that the arrival
    counters were left reusable."""
    from vllmini_b200 import _native
    c = build_case("ctx_mixed_long")
    t = to_cuda(c)
    truth = c_oracle().paged_attention_v1(c["q"], c["kc"], c["vc"], c["scale"], c["bt"], c["sl"], mode="exact")
    plan = _native.plan(c["S"], c["H"], c["H"], c["D"], c["B"], max_seq_len)
    want_splits = {64: 1, 512: 2, 1280: 5, 2048: 8, 4608: 9}[max_seq_len]
    assert f"grid=(12,3,{want_splits})" in plan, plan
    got = run_cuda_attention(c, t, max_seq_len=max_seq_len)
    got2 = run_cuda_attention(c, t, max_seq_len=max_seq_len)
    assert np.abs(got.astype(np.float32) - got2.astype(np.float32)).max() < 1e-3
    assert_close_to_truth(got, truth.astype(np.float64), atol=1.5e-3, what=f"max_seq_len {max_seq_len}: {plan}")


def test_experiments_build_shapes_and_forced_splits_agree():
    """The sweep build of the library (-DVB200_EXPERIMENTS: other warp x ring shapes, forced split factor) is not
    the product, but tools/sweep.py numbers come from it: every shape x split factor must match the fp64 oracle.
    Runs in a subprocess with VLLMINI_B200_LIB pointing at lib/experiments/ (skipped when that build is absent)."""
    import subprocess
    import sys
    from vllmini_b200 import build as vb_build, ops
    ops.reshape_and_cache(*_tiny_reshape_args())          # this process launches something too (autouse check)
    if not os.path.isfile(vb_build.EXP_LIB_PATH):
        pytest.skip("no experiments build (python -m vllmini_b200.build --experiments)")
    code = r"""
import numpy as np, sys
sys.path.insert(0, 'tests')
from helpers import build_case, to_cuda, run_cuda_attention, assert_close_to_truth
from oracle import c_oracle
from vllmini_b200 import _native
assert _native.has_experiments()
c = build_case('ctx_mixed_long'); t = to_cuda(c)
truth = c_oracle().paged_attention_v1(c['q'], c['kc'], c['vc'], c['scale'], c['bt'], c['sl'], mode='exact').astype(np.float64)
for variant in (0, 1, 2, 3, 4):
    for splits in (0, 2, 5):
        _native.set_tuning(splits, variant)
        got = run_cuda_attention(c, t)
        assert_close_to_truth(got, truth, atol=1.5e-3, what=f'variant {variant} splits {splits}')
_native.set_tuning(0, 0)
print('ok')
"""
    env = dict(os.environ, VLLMINI_B200_LIB=vb_build.EXP_LIB_PATH, VLLMINI_B200_NO_SHIM="1")
    r = subprocess.run([sys.executable, "-c", code], cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_speculative_prefetch_reads_stale_indices():
    """The first wave of an attention launch prefetches its first ring into L2 AHEAD of griddepcontrol.wait, from
    seq_lens / block-table entries that the previous kernel on the stream may not have written yet.  Here that
    previous kernel really is a PDL-aware producer of both tensors (experiments build: lets its dependents launch,
    holds them for 40 us, then writes the indices): the attention CTAs are resident and speculating on GARBAGE --
    random int32 block ids (negative, far beyond the pool) and lengths -- the whole time.  The hints must be harmless
    (no fault, nothing stale left behind) and the result must be the oracle's.  Subprocess: a fault would poison the
    CUDA context."""
    import subprocess
    import sys
    from vllmini_b200 import build as vb_build, ops
    ops.reshape_and_cache(*_tiny_reshape_args())          # this process launches something too (autouse check)
    if not os.path.isfile(vb_build.EXP_LIB_PATH):
        pytest.skip("no experiments build (python -m vllmini_b200.build --experiments)")
    code = r"""
import ctypes, numpy as np, sys, torch
sys.path.insert(0, 'tests')
from helpers import assert_close_to_truth
from oracle import c_oracle, paged_oracle as po
from vllmini_b200 import _native, ops
lib = _native.load()
assert _native.has_experiments()
fn = lib.vb200_debug_delayed_copy2
vp, i64, u32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint32
fn.restype, fn.argtypes = ctypes.c_int32, [vp, vp, i64, vp, vp, i64, u32, vp]
dev = torch.device('cuda:0')
rng = np.random.default_rng(77)
for S, lens_hi, B in ((64, 512, 16), (16, 900, 8), (2, 5000, 32), (200, 300, 16)):
    lens = rng.integers(1, lens_hi + 1, size=S).astype(np.int32); lens[0] = lens_hi
    q, kc, vc, bt, sl = po.make_case(S, 12, 64, B, lens, seed=int(S))
    truth = c_oracle().paged_attention_v1(q, kc, vc, 0.125, bt, sl, mode='exact').astype(np.float64)
    q_t, kc_t, vc_t, bt_good, sl_good = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (q, kc, vc, bt, sl))
    for trial in range(3):
        garbage_bt = torch.from_numpy(rng.integers(-2**31, 2**31, size=bt.shape, dtype=np.int64).astype(np.int32)).to(dev)
        garbage_sl = torch.from_numpy(rng.integers(-2**31, 2**31, size=sl.shape, dtype=np.int64).astype(np.int32)).to(dev)
        if trial == 1:
            garbage_sl.fill_(int(bt.shape[1]) * B)            # plausible lengths, garbage block ids: every hint is wild
        bt_t, sl_t = garbage_bt.clone(), garbage_sl.clone()
        out = torch.zeros(S, 12, 64, dtype=torch.float16, device=dev)
        torch.cuda.synchronize()
        st = torch.cuda.current_stream(dev).cuda_stream
        rc = fn(bt_t.data_ptr(), bt_good.data_ptr(), bt_t.numel(), sl_t.data_ptr(), sl_good.data_ptr(), sl_t.numel(), 40000, st)
        assert rc == 0
        ops.paged_attention_v1(out, q_t, kc_t, vc_t, 12, 0.125, bt_t, sl_t, B, int(bt.shape[1]) * B, None, 'auto', 1.0,
                               0, 0, 1, 1, 0)
        torch.cuda.synchronize()
        assert torch.equal(bt_t, bt_good) and torch.equal(sl_t, sl_good)
        assert_close_to_truth(out.cpu().numpy(), truth, atol=1.5e-3, what=f'S={S} trial {trial}')
print('ok')
"""
    env = dict(os.environ, VLLMINI_B200_LIB=vb_build.EXP_LIB_PATH, VLLMINI_B200_NO_SHIM="1")
    r = subprocess.run([sys.executable, "-c", code], cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def _tiny_reshape_args():
    dev = torch.device("cuda:0")
    k = torch.zeros(1, 1, 64, dtype=torch.float16, device=dev)
    kc = torch.zeros(1, 1, 8, 16, 8, dtype=torch.float16, device=dev)
    vc = torch.zeros(1, 1, 64, 16, dtype=torch.float16, device=dev)
    return k, k.clone(), kc, vc, torch.zeros(1, dtype=torch.int64, device=dev), "auto", 1.0


def test_two_streams_splitting_concurrently():
    """Two small-batch, long-context calls (both split-KV) issued on two streams so that they overlap on the GPU.
    The library keeps no state and needs no workspace (the splits merge inside their thread-block cluster), so --
    like the reference's v1 -- concurrent launches cannot disturb each other.  (Round 1's shared per-device
    workspace would have let the arrival counters and partials of the two launches collide.)"""
    from vllmini_b200 import ops
    dev = torch.device("cuda:0")
    cases = []
    for seed, lens in ((31, [8192, 5000]), (32, [7000, 8192])):
        q, kc, vc, bt, sl = po.make_case(2, 12, 64, 16, np.array(lens, dtype=np.int32), seed=seed)
        truth = c_oracle().paged_attention_v1(q, kc, vc, 0.125, bt, sl, mode="exact").astype(np.float64)
        t = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (q, kc, vc, bt, sl)]
        cases.append((t, truth, int(bt.shape[1]) * 16))
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    outs = [[torch.empty(2, 12, 64, dtype=torch.float16, device=dev) for _ in range(40)] for _ in cases]
    torch.cuda.synchronize()
    for i in range(40):                       # interleave the enqueues: the two streams run side by side
        for (t, _truth, msl), st, o in zip(cases, streams, outs):
            with torch.cuda.stream(st):
                ops.paged_attention_v1(o[i], t[0], t[1], t[2], 12, 0.125, t[3], t[4], 16, msl, None, "auto", 1.0,
                                       0, 0, 1, 1, 0)
    torch.cuda.synchronize()
    for (t, truth, _msl), o in zip(cases, outs):
        for i in range(40):
            assert_close_to_truth(o[i].cpu().numpy(), truth, what=f"stream-concurrent split-KV, call {i}")


def test_host_threads_calling_concurrently():
    """Four host threads, each on its own stream, call the C ABI through the ctypes binding (which releases the GIL)
    at the same time: first-use initialisation (function attributes, the cluster-size query), the launch counter and
    the per-thread error text must hold up, and every result must be the single-threaded one.  Shapes differ per
    thread so that different instantiations are initialised concurrently; one thread keeps provoking an error."""
    import threading
    from vllmini_b200 import _native, ops
    dev = torch.device("cuda:0")
    specs = [("ragged_small", None), ("head128_bs16", None), ("ctx_mixed_long", 4608), ("head80_bs8", None)]
    work = []
    for name, msl in specs:
        c = build_case(name)
        t = to_cuda(c)
        want = run_cuda_attention(c, t, max_seq_len=msl)          # single-threaded result, default stream
        work.append((c, t, msl, want))
    lib = _native.load()
    errors, results = [], [None] * len(work)
    barrier = threading.Barrier(len(work))

    def body(i):
        try:
            c, t, msl, _ = work[i]
            st = torch.cuda.Stream(dev)
            outs = []
            barrier.wait()
            with torch.cuda.stream(st):
                for it in range(150):
                    out = torch.empty((c["S"], c["H"], c["D"]), dtype=torch.float16, device=dev)
                    rc = lib.vb200_paged_attention_v1_f16(
                        out.data_ptr(), t["q"].data_ptr(), t["kc"].data_ptr(), t["vc"].data_ptr(), t["bt"].data_ptr(),
                        t["sl"].data_ptr(), c["S"], c["H"], c["kc"].shape[1], c["D"], c["B"], int(t["bt"].shape[1]),
                        msl or int(t["bt"].shape[1]) * c["B"], t["q"].stride(0), t["kc"].stride(0), t["kc"].stride(1),
                        float(c["scale"]), 0, st.cuda_stream)
                    assert rc == 0, _native.last_error()
                    outs.append(out)
                    if i == 0 and it % 10 == 0:                    # an error on this thread: its text stays this thread's
                        rc = lib.vb200_paged_attention_v1_f16(
                            out.data_ptr(), t["q"].data_ptr(), t["kc"].data_ptr(), t["vc"].data_ptr(), t["bt"].data_ptr(),
                            t["sl"].data_ptr(), c["S"], c["H"], c["kc"].shape[1], 72, c["B"], int(t["bt"].shape[1]),
                            64, t["q"].stride(0), t["kc"].stride(0), t["kc"].stride(1), 0.1, 0, st.cuda_stream)
                        assert rc != 0 and "head size" in _native.last_error()
                    elif it % 10 == 0:
                        assert "head size" not in _native.last_error()
            st.synchronize()
            results[i] = [o.cpu().numpy() for o in (outs[0], outs[len(outs) // 2], outs[-1])]
        except Exception as e:                                       # surfaced in the main thread
            errors.append((i, repr(e)))

    n0 = _native.launch_count()
    threads = [threading.Thread(target=body, args=(i,)) for i in range(len(work))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    torch.cuda.synchronize()
    assert not errors, errors
    assert _native.launch_count() - n0 == 150 * len(work)
    for (c, t, msl, want), got in zip(work, results):
        for g in got:
            assert np.array_equal(g.view(np.uint16), want.view(np.uint16)), "result differs from the single-threaded call"


def test_first_use_inside_cuda_graph_capture():
    """A fresh process whose FIRST call into the library happens while a CUDA graph is being captured: the once-per-
    device set-up (function attributes, the largest cluster size, how many clusters are resident at once) must not
    disturb the capture -- for a split launch (clusters) at a head size that holds one CTA per SM, an unsplit launch,
    reshape_and_cache and the fused append.  The replayed graph must give the eager results."""
    import subprocess
    import sys
    from vllmini_b200 import ops
    ops.reshape_and_cache(*_tiny_reshape_args())          # this process launches something too (autouse check)
    code = r"""
import numpy as np, sys, torch
sys.path.insert(0, 'tests')
from helpers import assert_close_to_truth
from oracle import c_oracle, paged_oracle as po
from vllmini_b200 import _native, ops
dev = torch.device('cuda:0')
cases = []
for S, H, D, B, lens in ((6, 8, 192, 16, [2048, 1500, 7, 640, 2000, 1]), (70, 12, 64, 16, None), (3, 4, 256, 32, [900, 50, 3000])):
    lens = np.asarray(lens if lens is not None else np.random.default_rng(1).integers(1, 300, size=S), dtype=np.int32)
    q, kc, vc, bt, sl = po.make_case(S, H, D, B, lens, seed=int(D))
    truth = c_oracle().paged_attention_v1(q, kc, vc, D ** -0.5, bt, sl, mode='exact').astype(np.float64)
    t = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (q, kc, vc, bt, sl)]
    cases.append((S, H, D, B, t, torch.zeros(S, H, D, dtype=torch.float16, device=dev), truth, int(bt.shape[1]) * B))
k_new = torch.randn(4, 12, 64, device=dev).to(torch.float16)
kc2 = torch.zeros(8, 12, 8, 16, 8, dtype=torch.float16, device=dev)
vc2 = torch.zeros(8, 12, 64, 16, dtype=torch.float16, device=dev)
slots = torch.tensor([5, 17, 40, 99], dtype=torch.int64, device=dev)
torch.cuda.synchronize()
assert _native.launch_count() == 0                       # nothing of the library has run in this process yet
side = torch.cuda.Stream(dev)
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(side):
    with torch.cuda.graph(g, stream=side):
        ops.reshape_and_cache(k_new, k_new, kc2, vc2, slots, 'auto', 1.0)
        for S, H, D, B, t, out, truth, msl in cases:
            ops.paged_attention_v1(out, t[0], t[1], t[2], H, D ** -0.5, t[3], t[4], B, msl, None, 'auto', 1.0, 0, 0, 1, 1, 0)
for _ in range(2):
    g.replay()
torch.cuda.synchronize()
for S, H, D, B, t, out, truth, msl in cases:
    assert_close_to_truth(out.cpu().numpy(), truth, atol=1.5e-3, what=f'captured cold, D={D}')
assert float(kc2.abs().sum()) > 0
print('ok')
"""
    r = subprocess.run([sys.executable, "-c", code], cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_small_batch_long_context_uses_split_kv():
    """S*H far below the SM count: the launcher splits the context (flash-decoding style) -- the splits of a
    (sequence, head) form a thread-block cluster and merge through distributed shared memory -- and must reproduce
    the unsplit result."""
    rng = np.random.default_rng(3)
    lens = np.array([4096, 3000], dtype=np.int32)
    q, kc, vc, bt, sl = po.make_case(2, 12, 64, 16, lens, seed=21)
    c = dict(S=2, H=12, D=64, B=16, q=q, kc=kc, vc=vc, bt=bt, sl=sl, scale=0.125)
    from vllmini_b200 import _native
    assert "grid=(12,2,11)" in _native.plan(2, 12, 12, 64, 16, int(bt.shape[1]) * 16)             # 11 splits, one cluster each
    got = run_cuda_attention(c)
    truth = c_oracle().paged_attention_v1(q, kc, vc, 0.125, bt, sl, mode="exact").astype(np.float64)
    assert_close_to_truth(got, truth, what="split-KV")


def test_strided_query_and_4d_out_like_vllmini():
    """vllmini passes q as a view of the fused qkv projection (stride(0) = 3*hidden, gpt2.py:36-41)
    and the reference test passes `out` as [S, H, 1, D] (tests/kernels/paged_attention.py:114)."""
    from vllmini_b200 import ops
    c = build_case("ragged_small")
    t = to_cuda(c)
    S, H, D = c["S"], c["H"], c["D"]
    qkv = torch.randn(S, 3 * H * D, device="cuda:0").to(torch.float16)
    qkv[:, :H * D] = t["q"].view(S, H * D)
    q_view = qkv[:, :H * D].view(S, H, D)
    assert q_view.stride(0) == 3 * H * D
    out4 = torch.empty(S, H, 1, D, dtype=torch.float16, device="cuda:0")
    msl = int(t["bt"].shape[1]) * c["B"]      # same max_seq_len => same split factor => same arithmetic
    ops.paged_attention_v1(out4, q_view, t["kc"], t["vc"], H, c["scale"], t["bt"], t["sl"], c["B"], msl,
                           None, "auto", 1.0, 0, 0, 1, 1, 0)
    torch.cuda.synchronize()
    want = run_cuda_attention(c, t, max_seq_len=msl)
    assert np.array_equal(out4.view(S, H, D).cpu().numpy().view(np.uint16), want.view(np.uint16))


def test_ctypes_and_torch_shim_bindings_are_the_same_path():
    """ops binds the C ABI either through the compiled torch shim or through ctypes; both must reach the
    same kernels and give bit-identical results."""
    from vllmini_b200 import _native, ops
    c = build_case("ragged_small")
    t = to_cuda(c)
    S, H, D, B = c["S"], c["H"], c["D"], c["B"]
    msl = int(t["bt"].shape[1]) * B
    outs = []
    for fn in (ops.paged_attention_v1, ops.paged_attention_v1_ctypes):
        o = torch.zeros(S, H, D, dtype=torch.float16, device="cuda:0")
        n0 = _native.launch_count()
        fn(o, t["q"], t["kc"], t["vc"], H, c["scale"], t["bt"], t["sl"], B, msl, None, "auto", 1.0, 0, 0, 1, 1, 0)
        torch.cuda.synchronize()
        assert _native.launch_count() == n0 + 1
        outs.append(o.cpu().numpy())
    assert np.array_equal(outs[0].view(np.uint16), outs[1].view(np.uint16))
    kcs = []
    for fn in (ops.reshape_and_cache, ops.reshape_and_cache_ctypes):
        kc, vc = torch.zeros_like(t["kc"]), torch.zeros_like(t["vc"])
        k = torch.randn(7, H, D, device="cuda:0", generator=torch.Generator(device="cuda:0").manual_seed(3)).half()
        fn(k, k, kc, vc, torch.arange(7, device="cuda:0") * 3, "auto", 1.0)
        torch.cuda.synchronize()
        kcs.append((kc.cpu(), vc.cpu()))
    assert torch.equal(kcs[0][0], kcs[1][0]) and torch.equal(kcs[0][1], kcs[1][1])
    with pytest.raises(RuntimeError, match="Unsupported block size"):
        ops.paged_attention_v1(outs and torch.zeros(S, H, D, dtype=torch.float16, device="cuda:0"), t["q"], t["kc"], t["vc"],
                               H, c["scale"], t["bt"], t["sl"], 4, msl, None, "auto", 1.0, 0, 0, 1, 1, 0)
    with pytest.raises(RuntimeError, match="Int"):
        ops.paged_attention_v1(torch.zeros(S, H, D, dtype=torch.float16, device="cuda:0"), t["q"], t["kc"], t["vc"],
                               H, c["scale"], t["bt"].long(), t["sl"], B, msl, None, "auto", 1.0, 0, 0, 1, 1, 0)


def test_garbage_beyond_seq_len_cannot_leak():
    c = build_case("block_plus_one")
    clean = run_cuda_attention(c)
    kc, vc = c["kc"].copy(), c["vc"].copy()
    for s in range(c["S"]):
        last = c["bt"][s, 1]
        kc[last, :, :, 1:, :] = np.nan
        vc[last, :, :, 1:] = np.inf
    unused = [b for b in range(kc.shape[0]) if b not in set(c["bt"][c["bt"] >= 0].tolist())]
    kc[unused] = np.nan
    vc[unused] = np.nan
    dirty = dict(c, kc=kc, vc=vc)
    got = run_cuda_attention(dirty)
    assert np.array_equal(got.view(np.uint16), clean.view(np.uint16))


def test_gqa_heads_share_kv_head():
    rng = np.random.default_rng(9)
    S, H, KVH, D, B = 3, 12, 4, 64, 16
    lens = np.array([40, 7, 300], dtype=np.int32)
    q, kc, vc, bt, sl = po.make_case(S, KVH, D, B, lens, seed=33)
    q = rng.standard_normal((S, H, D)).astype(np.float16)
    c = dict(S=S, H=H, D=D, B=B, q=q, kc=kc, vc=vc, bt=bt, sl=sl, scale=0.125)
    got = run_cuda_attention(c)
    truth = po.paged_attention_v1_exact(q, kc, vc, 0.125, bt, sl, num_kv_heads=KVH)
    assert_close_to_truth(got, truth, what="GQA")


def test_reference_kernel_test_restated():
    """vllmini/tests/kernels/paged_attention.py:84-139, same shapes and tolerances, against the
    drop-in package: reshape_and_cache 3 tokens into block 0, layout check, attention vs eager."""
    import vllmini_b200
    vllmini_b200.install_dropin()
    from paged_attention_cuda import cache_ops, paged_attention_v1
    torch.manual_seed(0)
    NB, H, D, B, bs, L = 1000, 12, 64, 16, 1, 3
    dev = torch.device("cuda:0")
    key_cache = torch.zeros(NB, H, D // 8, B, 8, dtype=torch.float16, device=dev)
    value_cache = torch.zeros(NB, H, D, B, dtype=torch.float16, device=dev)
    key = torch.randn(bs, L, H, D, dtype=torch.float16, device=dev)
    value = torch.randn(bs, L, H, D, dtype=torch.float16, device=dev)
    slot_mapping = torch.arange(L, dtype=torch.long, device=dev).repeat(bs)
    cache_ops.reshape_and_cache(key.reshape(-1, H, D), value.reshape(-1, H, D), key_cache, value_cache,
                                slot_mapping, "auto", 1.0)
    for h in range(H):
        for i in range(L):
            slot = int(slot_mapping[i])
            assert torch.allclose(key_cache[slot // B, h, :, slot % B, :].reshape(D), key[0, i, h], atol=1e-3)
            assert torch.allclose(value_cache[slot // B, h, :, slot % B], value[0, i, h], atol=1e-3)
    query = torch.randn(bs, 1, H, D, dtype=torch.float16, device=dev)
    scale = 1.0 / (D ** 0.5)
    w = torch.matmul(query.transpose(1, 2), key.transpose(1, 2).transpose(-1, -2)) * scale
    vanilla = torch.matmul(torch.softmax(w, dim=-1), value.transpose(1, 2))
    paged = torch.empty_like(vanilla)
    paged_attention_v1(paged, query.reshape(-1, H, D), key_cache, value_cache, H, scale,
                       torch.tensor([[0, -1, -1, -1]], dtype=torch.int32, device=dev),
                       torch.full((bs,), L, dtype=torch.int32, device=dev), B, L, None, "auto", 1.0, 0, 0, 1, 1, 0)
    assert torch.allclose(vanilla, paged, atol=1e-2)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN_DIR, "vanilla_*.npz"))))
def test_cuda_path_against_reference_vanilla_golden(path):
    """Full hot path on the GPU (reshape_and_cache then paged_attention_v1) against the committed
    outputs of the reference's own Python attention."""
    from vllmini_b200 import ops
    g = np.load(path)
    q, k, v, lens, want, scale = g["q"], g["k"], g["v"], g["seq_lens"], g["out"], float(g["scale"])
    S, Lmax, H, D = k.shape
    B = 16
    dev = torch.device("cuda:0")
    nblk = (lens.astype(np.int64) + B - 1) // B
    NB = int(nblk.sum()) + 3
    perm = np.random.default_rng(1).permutation(NB).astype(np.int32)
    bt = np.full((S, int(nblk.max())), -1, dtype=np.int32)
    kc = torch.zeros(NB, H, D // 8, B, 8, dtype=torch.float16, device=dev)
    vc = torch.zeros(NB, H, D, B, dtype=torch.float16, device=dev)
    o = 0
    for s in range(S):
        bt[s, :nblk[s]] = perm[o:o + nblk[s]]
        o += nblk[s]
        L = int(lens[s])
        slots = bt[s, np.arange(L) // B].astype(np.int64) * B + np.arange(L) % B
        ops.reshape_and_cache(torch.from_numpy(k[s, :L]).to(dev), torch.from_numpy(v[s, :L]).to(dev), kc, vc,
                              torch.from_numpy(slots).to(dev), "auto", 1.0)
    out = torch.empty(S, H, D, dtype=torch.float16, device=dev)
    ops.paged_attention_v1(out, torch.from_numpy(q).to(dev), kc, vc, H, scale, torch.from_numpy(bt).to(dev),
                           torch.from_numpy(lens).to(dev), B, int(bt.shape[1]) * B, None, "auto", 1.0, 0, 0, 1, 1, 0)
    torch.cuda.synchronize()
    assert np.abs(out.cpu().numpy().astype(np.float64) - want).max() < 1.5e-3


# ------------------------------------------------------------------------------ full-size properties
def _metric_config(S=256, H=12, D=64, B=16, ctx=1024, seed=0):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(seed)
    nblk = (ctx + B - 1) // B
    NB = S * nblk
    kc = torch.randn(NB, H, D // 8, B, 8, generator=g, device=dev, dtype=torch.float16)
    vc = torch.randn(NB, H, D, B, generator=g, device=dev, dtype=torch.float16)
    q = torch.randn(S, H, D, generator=g, device=dev, dtype=torch.float16)
    bt = torch.randperm(NB, generator=g, device=dev).to(torch.int32).view(S, nblk)
    sl = torch.full((S,), ctx, dtype=torch.int32, device=dev)
    return q, kc, vc, bt, sl


def test_full_size_properties_metric_config():
    """BASELINE metric shape (bs=256, ctx=1024, 12x64, block 16): too big for the CPU oracle in
    full, so check size-independent properties + an oracle spot check on a few sequences.
      * V == const 1  -> output == 1 (softmax weights sum to 1)
      * permuting the physical blocks (and the table with them) leaves the output unchanged
      * sequences are independent: a sub-batch gives the same rows
      * linearity in V: out(V1 + V2) ~= out(V1) + out(V2)
    """
    from vllmini_b200 import ops
    q, kc, vc, bt, sl = _metric_config()
    S, H, D = q.shape
    B = 16

    def attn(q_, kc_, vc_, bt_, sl_):
        o = torch.empty(q_.shape, dtype=torch.float16, device=q_.device)
        ops.paged_attention_v1(o, q_, kc_, vc_, H, 0.125, bt_, sl_, B, int(bt_.shape[1]) * B, None, "auto", 1.0,
                               0, 0, 1, 1, 0)
        return o

    base = attn(q, kc, vc, bt, sl)
    torch.cuda.synchronize()
    assert torch.isfinite(base).all()
    # weights sum to one
    ones = attn(q, kc, torch.ones_like(vc), bt, sl)
    assert (ones.float() - 1.0).abs().max().item() <= 2e-3
    # block permutation invariance (bit-exact: same arithmetic, different addresses)
    perm = torch.randperm(kc.shape[0], device=kc.device)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(perm.numel(), device=perm.device)
    moved = attn(q, kc[perm], vc[perm], inv[bt.long()].to(torch.int32), sl)
    assert torch.equal(moved, base)
    # independence of sequences.  A 3-sequence batch is launched with split-KV (different
    # summation order), so equality is to fp16 rounding, not bit-exact.
    idx = torch.tensor([0, 17, 255], device=q.device)
    sub = attn(q[idx].contiguous(), kc, vc, bt[idx].contiguous(), sl[idx].contiguous())
    assert (sub.float() - base[idx].float()).abs().max().item() <= 2.5e-4
    # linearity in V
    v2 = torch.randn_like(vc)
    lin = attn(q, kc, (vc.float() + v2.float()).to(torch.float16), bt, sl)
    parts = base.float() + attn(q, kc, v2, bt, sl).float()
    assert (lin.float() - parts).abs().max().item() < 4e-3
    # oracle spot check on three sequences
    k_np, v_np = kc.cpu().numpy(), vc.cpu().numpy()
    rows = [0, 100, 255]
    truth = c_oracle().paged_attention_v1(q[rows].cpu().numpy(), k_np, v_np, 0.125, bt[rows].cpu().numpy(),
                                          sl[rows].cpu().numpy(), mode="exact").astype(np.float64)
    assert_close_to_truth(base[rows].cpu().numpy(), truth, atol=5e-4, what="metric config spot check")


def test_full_size_round_trip_reshape_then_attend():
    """reshape_and_cache -> paged_attention_v1 round trip at bs=256: appending a token whose key
    equals a huge multiple of q makes attention collapse onto that token's value row."""
    from vllmini_b200 import ops
    q, kc, vc, bt, sl = _metric_config(ctx=1008, seed=1)      # 63 blocks used; slot for token 1008 is free
    S, H, D = q.shape
    B = 16
    dev = q.device
    extra_blocks = torch.arange(S, device=dev, dtype=torch.int32) + kc.shape[0]
    kc = torch.cat([kc, torch.zeros(S, H, D // 8, B, 8, dtype=torch.float16, device=dev)])
    vc = torch.cat([vc, torch.zeros(S, H, D, B, dtype=torch.float16, device=dev)])
    bt = torch.cat([bt, extra_blocks[:, None]], dim=1).contiguous()
    new_k = (q.float() * 8.0).to(torch.float16)               # logit = 8*|q|^2*0.125 ~ 64 >> others
    new_v = torch.randn(S, H, D, device=dev, dtype=torch.float16)
    slots = extra_blocks.long() * B                            # offset 0 of each new block
    ops.reshape_and_cache(new_k, new_v, kc, vc, slots, "auto", 1.0)
    sl2 = sl + 1
    out = torch.empty_like(q)
    ops.paged_attention_v1(out, q, kc, vc, H, 0.125, bt, sl2, B, int(bt.shape[1]) * B, None, "auto", 1.0, 0, 0, 1, 1, 0)
    torch.cuda.synchronize()
    assert (out.float() - new_v.float()).abs().max().item() < 2e-2


# ------------------------------------------------------------------------------ rest of the exported surface
@pytest.mark.parametrize("name,msl", [("ctx_mixed_long", None), ("ctx_mixed_long", 2048), ("ragged_small", None),
                                      ("with_empty_seq", 600), ("head128_bs16", None), ("head80_bs8", 1024),
                                      ("bs32_ragged", None), ("head256_bs32", 1030)])
def test_paged_attention_v2_writes_the_callers_scratch_like_the_reference(name, msl, ref_ext):
    """paged_attention_v2 (attention_kernels.cu:966-990): `out` AND the three scratch tensors are compared with the
    reference's own v2 kernels on the same inputs.  Partitions are the reference's fixed 512 tokens; a partition
    without tokens is left untouched by both (the sentinel survives).  Tolerances: max_logits are the same fp32
    dot products in a different summation order (1e-4 relative); exp_sums differ by ex2.approx vs __expf and by the
    fp16 rounding of p this kernel sums (2e-3 relative); tmp_out / out as for v1."""
    from vllmini_b200 import ops
    c = build_case(name)
    t = to_cuda(c)
    S, H, D, B = c["S"], c["H"], c["D"], c["B"]
    msl = msl or int(t["bt"].shape[1]) * B
    P = (msl + 511) // 512
    dev = torch.device("cuda:0")

    def run(fn):
        out = torch.empty(S, H, D, dtype=torch.float16, device=dev)
        exp_sums = torch.full((S, H, P), -777.0, dtype=torch.float32, device=dev)
        max_logits = torch.full((S, H, P), -777.0, dtype=torch.float32, device=dev)
        tmp_out = torch.full((S, H, P, D), -7.0, dtype=torch.float16, device=dev)
        fn(out, exp_sums, max_logits, tmp_out, t["q"], t["kc"], t["vc"], c["kc"].shape[1], c["scale"], t["bt"], t["sl"],
           B, msl, None, "auto", 1.0, 0, 0, 1, 1, 0)
        torch.cuda.synchronize()
        return [x.cpu().numpy() for x in (out, exp_sums, max_logits, tmp_out)]

    out, es, ml, to = run(ops.paged_attention_v2)
    lens = np.minimum(c["sl"].astype(np.int64), msl)
    live = (np.arange(P)[None, :] * 512 < lens[:, None])                     # [S, P] partitions that hold tokens
    assert (es[~live[:, None, :].repeat(H, 1)] == -777.0).all() and (ml[~live[:, None, :].repeat(H, 1)] == -777.0).all()
    assert (to[~live[:, None, :].repeat(H, 1)] == -7.0).all(), "a partition beyond seq_len was written"
    truth = c_oracle().paged_attention_v1(c["q"], c["kc"], c["vc"], c["scale"], c["bt"],
                                          lens.astype(np.int32), mode="exact").astype(np.float64)
    assert_close_to_truth(out, truth, what=f"v2 out ({name})")
    # the records reproduce the output on the host: out = sum_j tmp_out_j * es_j * exp(ml_j - M) / (sum + 1e-6)
    for s_ in range(S):
        n = int(live[s_].sum())
        if n == 0:
            assert not out[s_].any()
            continue
        M = ml[s_, :, :n].max(axis=-1, keepdims=True)
        w = es[s_, :, :n] * np.exp(ml[s_, :, :n] - M)
        host = (to[s_, :, :n].astype(np.float64) * w[..., None]).sum(axis=1) / (w.sum(axis=1, keepdims=True) + 1e-6)
        assert np.abs(host - out[s_].astype(np.float64)).max() < 2e-3
    if ref_ext is not None:
        r_out, r_es, r_ml, r_to = run(ref_ext.paged_attention_v2)
        m = live[:, None, :].repeat(H, 1)
        assert np.allclose(ml[m], r_ml[m], rtol=1e-4, atol=1e-4), np.abs(ml[m] - r_ml[m]).max()
        assert np.allclose(es[m], r_es[m], rtol=2e-3, atol=1e-4), np.abs(es[m] / r_es[m] - 1).max()
        d = np.abs(to[m].astype(np.float64) - r_to[m].astype(np.float64))
        assert (d <= ATOL_VS_REF + RTOL_VS_REF * np.abs(r_to[m].astype(np.float64))).all(), d.max()
        d = np.abs(out.astype(np.float64) - r_out.astype(np.float64))
        assert (d <= ATOL_VS_REF + RTOL_VS_REF * np.abs(r_out.astype(np.float64))).all(), d.max()
        # untouched partitions: the reference leaves the sentinel too
        assert (r_es[~m] == -777.0).all()


def test_paged_attention_v2_validates_like_the_reference():
    from vllmini_b200 import ops
    ops.reshape_and_cache(*_tiny_reshape_args())
    c = build_case("ragged_small")
    t = to_cuda(c)
    S, H, D, B = c["S"], c["H"], c["D"], c["B"]
    dev = torch.device("cuda:0")
    out = torch.empty(S, H, D, dtype=torch.float16, device=dev)
    es = torch.empty(S, H, 1, dtype=torch.float32, device=dev)
    to = torch.empty(S, H, 1, D, dtype=torch.float16, device=dev)
    args = (t["q"], t["kc"], t["vc"], H, c["scale"], t["bt"], t["sl"])
    with pytest.raises(RuntimeError, match="Unsupported block size"):
        ops.paged_attention_v2(out, es, es.clone(), to, *args, 4, 512, None, "auto", 1.0, 0, 0, 1, 1, 0)
    with pytest.raises(RuntimeError, match="exp_sums/max_logits must be"):
        ops.paged_attention_v2(out, es, es.clone(), to, *args, B, 513, None, "auto", 1.0, 0, 0, 1, 1, 0)
    with pytest.raises(RuntimeError, match="Unsupported data type of kv cache"):
        ops.paged_attention_v2(out, es, es.clone(), to, *args, B, 512, None, "fp8", 1.0, 0, 0, 1, 1, 0)
    with pytest.raises(RuntimeError, match="Int"):
        ops.paged_attention_v2(out, es, es.clone(), to, t["q"], t["kc"], t["vc"], H, c["scale"], t["bt"].long(), t["sl"],
                               B, 512, None, "auto", 1.0, 0, 0, 1, 1, 0)


@pytest.mark.parametrize("T,H,D,B,NB", [(53, 12, 64, 16, 6), (20, 3, 80, 8, 4), (300, 2, 256, 32, 10)])
def test_reshape_and_cache_flash_bit_exact(T, H, D, B, NB, ref_ext):
    from vllmini_b200 import ops
    rng = np.random.default_rng(11)
    qkv = rng.standard_normal((T, 3 * H * D)).astype(np.float16)
    slots = rng.permutation(NB * B)[:T].astype(np.int64)
    slots[7] = -1
    dev = torch.device("cuda:0")
    qkv_t = torch.from_numpy(qkv).to(dev)
    k_t, v_t = qkv_t[:, H * D:2 * H * D].view(T, H, D), qkv_t[:, 2 * H * D:].view(T, H, D)
    kc = torch.full((NB, B, H, D), 3.0, dtype=torch.float16, device=dev)
    vc = torch.full((NB, B, H, D), 3.0, dtype=torch.float16, device=dev)
    ops.reshape_and_cache_flash(k_t, v_t, kc, vc, torch.from_numpy(slots).to(dev), "auto")
    torch.cuda.synchronize()
    kc_w = np.full((NB, B, H, D), 3.0, dtype=np.float16)
    vc_w = kc_w.copy()
    po.reshape_and_cache_flash(qkv[:, H * D:2 * H * D].reshape(T, H, D), qkv[:, 2 * H * D:].reshape(T, H, D), kc_w, vc_w, slots)
    assert np.array_equal(kc.cpu().numpy().view(np.uint16), kc_w.view(np.uint16))
    assert np.array_equal(vc.cpu().numpy().view(np.uint16), vc_w.view(np.uint16))
    if ref_ext is not None:      # and against the reference's own kernel
        kc_r = torch.full((NB, B, H, D), 3.0, dtype=torch.float16, device=dev)
        vc_r = torch.full((NB, B, H, D), 3.0, dtype=torch.float16, device=dev)
        ref_ext.cache_ops.reshape_and_cache_flash(k_t, v_t, kc_r, vc_r, torch.from_numpy(slots).to(dev), "auto")
        torch.cuda.synchronize()
        assert torch.equal(kc_r, kc) and torch.equal(vc_r, vc)


def test_copy_blocks_and_swap_blocks_bit_exact(ref_ext):
    from vllmini_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    L, NB, H, D, B = 3, 10, 4, 64, 16
    kcs = [torch.randn(NB, H, D // 8, B, 8, generator=g, device=dev).half() for _ in range(L)]
    vcs = [torch.randn(NB, H, D, B, generator=g, device=dev).half() for _ in range(L)]
    mapping = torch.tensor([[0, 5], [1, 6], [2, 9]], dtype=torch.int64)
    k_np = [k.cpu().numpy().copy() for k in kcs]
    v_np = [v.cpu().numpy().copy() for v in vcs]
    po.copy_blocks(k_np, v_np, mapping.numpy())
    if ref_ext is not None:
        kr, vr = [k.clone() for k in kcs], [v.clone() for v in vcs]
        ref_ext.cache_ops.copy_blocks(kr, vr, mapping.to(dev))
    ops.copy_blocks(kcs, vcs, mapping.to(dev))
    torch.cuda.synchronize()
    for l in range(L):
        assert np.array_equal(kcs[l].cpu().numpy().view(np.uint16), k_np[l].view(np.uint16))
        assert np.array_equal(vcs[l].cpu().numpy().view(np.uint16), v_np[l].view(np.uint16))
        if ref_ext is not None:
            assert torch.equal(kr[l], kcs[l]) and torch.equal(vr[l], vcs[l])
    # swap: GPU -> CPU, CPU -> GPU, GPU -> GPU
    src = kcs[0]
    host = torch.zeros(NB, H, D // 8, B, 8, dtype=torch.float16).pin_memory()
    ops.swap_blocks(src, host, torch.tensor([[3, 0], [4, 7]], dtype=torch.int64))
    torch.cuda.synchronize()
    assert torch.equal(host[0], src[3].cpu()) and torch.equal(host[7], src[4].cpu()) and not host[1].any()
    back = torch.zeros_like(src)
    ops.swap_blocks(host, back, torch.tensor([[0, 2], [7, 1]], dtype=torch.int64))
    other = torch.zeros_like(src)
    ops.swap_blocks(src, other, torch.tensor([[9, 9]], dtype=torch.int64))
    torch.cuda.synchronize()
    assert torch.equal(back[2], src[3]) and torch.equal(back[1], src[4]) and torch.equal(other[9], src[9])
    if ref_ext is not None:
        host_r = torch.zeros(NB, H, D // 8, B, 8, dtype=torch.float16).pin_memory()
        ref_ext.cache_ops.swap_blocks(src, host_r, torch.tensor([[3, 0], [4, 7]], dtype=torch.int64))
        torch.cuda.synchronize()
        assert torch.equal(host_r, host)


# ------------------------------------------------------------------------------ fused append + attend
@pytest.mark.parametrize("seed", range(10))
def test_fused_append_random_shapes(seed):
    """Seeded random decode steps through the fused launch: head size, block size, MHA / grouped-query, batch (from a
    split-KV handful to several hundred rows, the latter large enough for the speculative first-ring prefetch),
    histories from brand-new sequences to a few thousand tokens, both position conventions -- caches and output bit
    for bit equal to reshape_and_cache + paged_attention_v1, output against the C oracle on the updated cache."""
    from vllmini_b200 import ops
    rng = np.random.default_rng(4000 + seed)
    D = int(rng.choice([64, 64, 80, 96, 112, 128, 192, 256]))
    B = int(rng.choice([8, 16, 32]))
    KVH = int(rng.choice([1, 2, 4, 12]))
    G = int(rng.choice([1, 1, 2, 4]))
    H = KVH * G
    S = int(rng.choice([1, 2, 5, 33, 150, 300]))
    hi = int(rng.choice([60, 400])) if S > 5 else int(rng.choice([400, 3000]))
    hist = rng.integers(0, hi + 1, size=S)
    hist[rng.integers(0, S)] = 0                              # a brand-new sequence
    hist[rng.integers(0, S)] = hi
    self_attend = bool(seed % 2)
    lens_after = hist + 1
    q, kc, vc, bt, _ = po.make_case(S, KVH, D, B, lens_after, seed=500 + seed,
                                    max_blocks_per_seq=int(lens_after.max() // B + 2))
    q = rng.standard_normal((S, H, D)).astype(np.float16)
    dev = torch.device("cuda:0")
    new_qkv = torch.from_numpy(rng.standard_normal((S, (H + 2 * KVH) * D)).astype(np.float16)).to(dev)
    k_new = new_qkv[:, H * D:(H + KVH) * D].view(S, KVH, D)
    v_new = new_qkv[:, (H + KVH) * D:].view(S, KVH, D)
    slots = torch.from_numpy(bt[np.arange(S), hist // B].astype(np.int64) * B + hist % B).to(dev)
    seq_lens_np = (lens_after if self_attend else hist).astype(np.int32)
    seq_lens = torch.from_numpy(seq_lens_np).to(dev)
    bt_t, q_t = torch.from_numpy(bt).to(dev), torch.from_numpy(q).to(dev)
    msl = int(bt.shape[1]) * B
    scale = float(D) ** -0.5
    kc_a, vc_a = torch.from_numpy(kc).to(dev), torch.from_numpy(vc).to(dev)
    out_a = torch.empty(S, H, D, dtype=torch.float16, device=dev)
    ops.reshape_and_cache(k_new, v_new, kc_a, vc_a, slots, "auto", 1.0)
    ops.paged_attention_v1(out_a, q_t, kc_a, vc_a, KVH, scale, bt_t, seq_lens, B, msl, None, "auto", 1.0, 0, 0, 1, 1, 0)
    kc_b, vc_b = torch.from_numpy(kc).to(dev), torch.from_numpy(vc).to(dev)
    out_b = torch.empty(S, H, D, dtype=torch.float16, device=dev)
    ops.paged_attention_v1_append(out_b, q_t, k_new, v_new, kc_b, vc_b, slots, KVH, scale, bt_t, seq_lens, B, msl)
    torch.cuda.synchronize()
    what = f"seed {seed}: S={S} H={H}/{KVH} D={D} B={B} hist<= {hi} self_attend={self_attend}"
    assert torch.equal(kc_a, kc_b) and torch.equal(vc_a, vc_b), what + ": fused scatter differs from reshape_and_cache"
    assert torch.equal(out_a, out_b), what + ": fused attention differs from the two-launch result"
    kc_x = np.repeat(kc_a.cpu().numpy(), G, axis=1)
    vc_x = np.repeat(vc_a.cpu().numpy(), G, axis=1)
    truth = c_oracle().paged_attention_v1(q, kc_x, vc_x, scale, bt, seq_lens_np, mode="exact").astype(np.float64)
    assert_close_to_truth(out_b.cpu().numpy(), truth, atol=1.5e-3, what=what)


@pytest.mark.parametrize("self_attend", [True, False])
@pytest.mark.parametrize("shape", ["batch", "long_small_batch", "gqa", "head80_blk8", "head256_blk32"])
def test_fused_append_equals_reshape_then_attention(self_attend, shape):
    """vb200_paged_attention_v1_append_f16 == reshape_and_cache followed by paged_attention_v1, bit for bit
    (same kernel arithmetic), for both position conventions: the new token attends to itself
    (seq_len counts it) or not (the reference scheduler's off-by-one, scheduler.py:96-101)."""
    from vllmini_b200 import ops
    rng = np.random.default_rng(17)
    if shape == "batch":
        S, H, KVH, D, B = 37, 12, 12, 64, 16
        hist = rng.integers(0, 200, size=S)                 # tokens already cached (0 = brand-new sequence)
    elif shape == "long_small_batch":
        S, H, KVH, D, B = 2, 12, 12, 64, 16                 # split-KV path
        hist = np.array([3000, 2047])
    elif shape == "head80_blk8":
        S, H, KVH, D, B = 6, 4, 4, 80, 8                    # partial last chunk per lane
        hist = np.array([0, 7, 8, 9, 63, 200])
    elif shape == "head256_blk32":
        S, H, KVH, D, B = 3, 4, 2, 256, 32                  # q in shared memory, 3-warp CTA
        hist = np.array([0, 95, 300])
    else:
        S, H, KVH, D, B = 5, 8, 2, 128, 32
        hist = np.array([0, 31, 32, 33, 100])
    lens_after = hist + 1
    q, kc, vc, bt, _ = po.make_case(S, KVH, D, B, lens_after, seed=23, max_blocks_per_seq=int(lens_after.max() // B + 2))
    q = rng.standard_normal((S, H, D)).astype(np.float16)
    dev = torch.device("cuda:0")
    new_qkv = torch.from_numpy(rng.standard_normal((S, (H + 2 * KVH) * D)).astype(np.float16)).to(dev)
    k_new = new_qkv[:, H * D:(H + KVH) * D].view(S, KVH, D)          # strided views like gpt2.py:36-41
    v_new = new_qkv[:, (H + KVH) * D:].view(S, KVH, D)
    slots = torch.from_numpy(bt[np.arange(S), hist // B].astype(np.int64) * B + hist % B).to(dev)
    seq_lens = torch.from_numpy((lens_after if self_attend else hist).astype(np.int32)).to(dev)
    bt_t, q_t = torch.from_numpy(bt).to(dev), torch.from_numpy(q).to(dev)
    msl = int(bt.shape[1]) * B

    kc_a, vc_a = torch.from_numpy(kc).to(dev), torch.from_numpy(vc).to(dev)
    out_a = torch.empty(S, H, D, dtype=torch.float16, device=dev)
    ops.reshape_and_cache(k_new, v_new, kc_a, vc_a, slots, "auto", 1.0)
    ops.paged_attention_v1(out_a, q_t, kc_a, vc_a, KVH, 0.1, bt_t, seq_lens, B, msl, None, "auto", 1.0, 0, 0, 1, 1, 0)

    kc_b, vc_b = torch.from_numpy(kc).to(dev), torch.from_numpy(vc).to(dev)
    out_b = torch.empty(S, H, D, dtype=torch.float16, device=dev)
    ops.paged_attention_v1_append(out_b, q_t, k_new, v_new, kc_b, vc_b, slots, KVH, 0.1, bt_t, seq_lens, B, msl)
    torch.cuda.synchronize()
    assert torch.equal(kc_a, kc_b) and torch.equal(vc_a, vc_b), "fused scatter differs from reshape_and_cache"
    assert torch.equal(out_a, out_b), "fused attention differs from the two-launch result"
    # and both agree with the oracle run on the updated cache
    truth = po.paged_attention_v1_exact(q, kc_a.cpu().numpy(), vc_a.cpu().numpy(), 0.1, bt, seq_lens.cpu().numpy(),
                                        num_kv_heads=KVH)
    assert_close_to_truth(out_b.cpu().numpy(), truth, what="fused append")


# ------------------------------------------------------------------------------ size extremes
@pytest.mark.parametrize("S,H,ctx_lo,ctx_hi,B", [(1, 12, 32768, 32768, 16),   # one very long sequence: split-KV, 2048-entry table row
                                                  (4096, 12, 1, 64, 16),       # many short sequences: 49 152 CTAs
                                                  (65535, 2, 1, 16, 8)])       # the grid.y limit (the reference's grid is (H, S) too)
def test_size_extremes_against_c_oracle(S, H, ctx_lo, ctx_hi, B):
    D = 64
    rng = np.random.default_rng(S)
    lens = rng.integers(ctx_lo, ctx_hi + 1, size=S).astype(np.int32)
    q, kc, vc, bt, sl = po.make_case(S, H, D, B, lens, seed=S % 1000)
    case = dict(S=S, H=H, D=D, B=B, q=q, kc=kc, vc=vc, bt=bt, sl=sl, scale=0.125)
    got = run_cuda_attention(case)
    rows = np.unique(np.concatenate([np.arange(min(S, 64)), rng.integers(0, S, size=64), [S - 1]]))
    used = np.unique(bt[rows][bt[rows] >= 0])
    remap = np.full(kc.shape[0], -1, dtype=np.int32)
    remap[used] = np.arange(used.size, dtype=np.int32)
    bt_small = np.where(bt[rows] >= 0, remap[np.maximum(bt[rows], 0)], -1).astype(np.int32)
    truth = c_oracle().paged_attention_v1(np.ascontiguousarray(q[rows]), kc[used], vc[used], 0.125, bt_small, sl[rows],
                                          mode="exact").astype(np.float64)
    assert_close_to_truth(got[rows], truth, atol=1.5e-3, what=f"S={S} ctx {ctx_lo}..{ctx_hi}")
    assert np.isfinite(got.astype(np.float32)).all()


def test_more_sequences_than_the_grid_allows_is_an_error():
    from vllmini_b200 import ops
    dev = torch.device("cuda:0")
    S, H, D, B = 65536, 1, 64, 8
    q = torch.zeros(S, H, D, dtype=torch.float16, device=dev)
    kc = torch.zeros(2, H, D // 8, B, 8, dtype=torch.float16, device=dev)
    vc = torch.zeros(2, H, D, B, dtype=torch.float16, device=dev)
    bt = torch.zeros(S, 1, dtype=torch.int32, device=dev)
    sl = torch.ones(S, dtype=torch.int32, device=dev)
    with pytest.raises(RuntimeError, match="grid.y"):
        ops.paged_attention_v1(torch.empty_like(q), q, kc, vc, H, 0.125, bt, sl, B, B, None, "auto", 1.0, 0, 0, 1, 1, 0)
    from vllmini_b200 import _native
    # keep the autouse launch-count fixture satisfied: one valid launch
    ops.paged_attention_v1(torch.empty_like(q[:2]), q[:2], kc, vc, H, 0.125, bt[:2], sl[:2], B, B, None, "auto", 1.0,
                           0, 0, 1, 1, 0)
    torch.cuda.synchronize()
    assert _native.launch_count() > 0


# ------------------------------------------------------------------------------ randomized shapes
def test_randomized_shapes_against_c_oracle():
    """60 seeded random problems: batch 1..40, heads (MHA and GQA), every head size, block 8/16/32, ragged
    lengths 0..700, padded tables, strided queries, random max_seq_len slack -- all against the fp64 oracle."""
    from vllmini_b200 import ops
    rng = np.random.default_rng(2024)
    co = c_oracle()
    dev = torch.device("cuda:0")
    worst = 0.0
    for trial in range(60):
        D = int(rng.choice([64, 128, 64, 80, 96, 112, 192, 256]))
        B = int(rng.choice([8, 16, 32]))
        KVH = int(rng.choice([1, 2, 4, 12]))
        H = KVH * int(rng.choice([1, 1, 2, 4]))
        S = int(rng.integers(1, 41))
        hi = int(rng.choice([20, 100, 700]))
        lens = rng.integers(0, hi + 1, size=S).astype(np.int32)
        if trial % 7 == 0:
            lens[:] = int(rng.integers(1, hi + 1))
        mb = int(max((int(lens.max()) + B - 1) // B, 1)) + int(rng.integers(0, 3))
        q, kc, vc, bt, sl = po.make_case(S, KVH, D, B, lens, seed=trial, max_blocks_per_seq=mb,
                                         num_blocks=int(((lens.astype(np.int64) + B - 1) // B).sum()) + 3)
        pad = int(rng.choice([0, 8, 2 * H * D]))
        qbuf = rng.standard_normal((S, H * D + pad)).astype(np.float16)
        q_np = qbuf[:, :H * D].reshape(S, H, D)
        q_t = torch.from_numpy(qbuf).to(dev)[:, :H * D].view(S, H, D)
        assert S == 1 or q_t.stride(0) == H * D + pad      # torch normalises strides of size-1 dims
        out = torch.empty(S, H, D, dtype=torch.float16, device=dev)
        scale = float(rng.choice([D ** -0.5, 0.05, 0.3]))
        msl = mb * B if trial % 3 else max(int(lens.max()), 1)
        ops.paged_attention_v1(out, q_t, torch.from_numpy(kc).to(dev), torch.from_numpy(vc).to(dev), KVH, scale,
                               torch.from_numpy(bt).to(dev), torch.from_numpy(sl).to(dev), B, msl, None, "auto", 1.0,
                               0, 0, 1, 1, 0)
        torch.cuda.synchronize()
        truth = co.paged_attention_v1(np.ascontiguousarray(q_np), kc, vc, scale, bt, sl, num_kv_heads=KVH,
                                      mode="exact").astype(np.float64)
        worst = max(worst, assert_close_to_truth(out.cpu().numpy(), truth, atol=1.5e-3,
                                                 what=f"trial {trial} S={S} H={H}/{KVH} D={D} B={B}"))
    assert worst < 2.5e-3


# ------------------------------------------------------------------------------ BASELINE configs at full size
def _device_case(S, H, D, B, lens, seed):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(seed)
    lens_t = torch.as_tensor(lens, dtype=torch.int32, device=dev)
    nblk = (int(lens_t.max()) + B - 1) // B
    NB = S * nblk
    kc = torch.empty(NB, H, D // 8, B, 8, device=dev, dtype=torch.float16).normal_(generator=g)
    vc = torch.empty(NB, H, D, B, device=dev, dtype=torch.float16).normal_(generator=g)
    q = torch.randn(S, H, D, generator=g, device=dev, dtype=torch.float16)
    bt = torch.randperm(NB, generator=g, device=dev).to(torch.int32).view(S, nblk)
    return q, kc, vc, bt, lens_t


@pytest.mark.parametrize("name,S,H,D,B,ctx", [
    ("configs[1] bs=64 ctx=512", 64, 12, 64, 16, 512),
    ("metric bs=256 ctx=1024", 256, 12, 64, 16, 1024),
    ("configs[2] mixed 128-2048", 256, 12, 64, 16, -2048),
    ("configs[3] blk8 ctx4k 20x64", 256, 20, 64, 8, 4096),
    ("configs[3] blk32 ctx8k 20x64", 256, 20, 64, 32, 8192),
])
def test_baseline_configs_full_size_vs_reference_kernel(name, S, H, D, B, ctx, ref_ext):
    """Every BASELINE.json config at its real size: the new kernel against the reference's own kernel on the
    same tensors (tolerance of SURVEY.md 8c), plus an fp64-oracle spot check on two sequences."""
    if ref_ext is None:
        pytest.skip("oracle/_ref not built")
    from vllmini_b200 import ops
    if ctx < 0:
        lens = torch.randint(128, -ctx + 1, (S,), generator=torch.Generator().manual_seed(0)).tolist()
    else:
        lens = [ctx] * S
    q, kc, vc, bt, sl = _device_case(S, H, D, B, lens, seed=S + B)
    msl = int(bt.shape[1]) * B
    new = torch.empty_like(q)
    ref = torch.empty_like(q)
    ops.paged_attention_v1(new, q, kc, vc, H, D ** -0.5, bt, sl, B, msl, None, "auto", 1.0, 0, 0, 1, 1, 0)
    ref_ext.paged_attention_v1(ref, q, kc, vc, H, D ** -0.5, bt, sl, B, msl, None, "auto", 1.0, 0, 0, 1, 1, 0)
    torch.cuda.synchronize()
    diff = (new.float() - ref.float()).abs()
    assert bool((diff <= ATOL_VS_REF + RTOL_VS_REF * ref.float().abs()).all()), f"{name}: max diff {diff.max().item():.3e}"
    rows = [0, S - 1]
    kc_np, vc_np = None, None
    # gather only the blocks of the two spot-check sequences for the CPU oracle
    used = torch.unique(bt[rows].flatten()).long()
    remap = torch.full((kc.shape[0],), -1, dtype=torch.int32, device=kc.device)
    remap[used] = torch.arange(used.numel(), dtype=torch.int32, device=kc.device)
    truth = c_oracle().paged_attention_v1(q[rows].cpu().numpy(), kc[used].cpu().numpy(), vc[used].cpu().numpy(),
                                          D ** -0.5, remap[bt[rows].long()].cpu().numpy(), sl[rows].cpu().numpy(),
                                          mode="exact").astype(np.float64)
    err_new = np.abs(new[rows].cpu().numpy().astype(np.float64) - truth).max()
    err_ref = np.abs(ref[rows].cpu().numpy().astype(np.float64) - truth).max()
    assert err_new <= max(5e-4, err_ref), f"{name}: new {err_new:.2e} vs reference kernel {err_ref:.2e} from fp64 truth"


# ------------------------------------------------------------------------------ grouped-query kernel
@pytest.mark.parametrize("D,B,KVH,ratio", [(64, 16, 3, 4), (64, 8, 2, 2), (64, 32, 1, 8), (128, 16, 2, 4), (128, 32, 2, 4),
                                           (128, 8, 1, 6), (64, 16, 2, 3)])
def test_grouped_query_kernel_is_bit_identical_to_the_per_head_grid(D, B, KVH, ratio):
    """num_kv_heads < num_heads takes the kernel that applies each landed tile to 4 (or 2) query heads of its kv
    head; per head it performs the same operations in the same order as the per-query-head grid, so outputs must
    match bit for bit -- and the fp64 oracle within tolerance.  The per-head grid is reached by giving every query
    head its own copy of its kv head (an MHA cache with repeated heads).  Ratio 3 has no grouped instantiation and
    runs the per-head grid either way."""
    from vllmini_b200 import _native, ops
    rng = np.random.default_rng(D + B + KVH + ratio)
    H = KVH * ratio
    S = 4 * 148 // H + 2              # the grouped kernel is only chosen from 4 units per SM up (bandwidth-bound)
    lens = rng.integers(1, 200, size=S).astype(np.int32)   # >= 1: the append part below re-writes the last token
    lens[0], lens[1] = 1, 257
    q, kc, vc, bt, sl = po.make_case(S, KVH, D, B, lens, seed=ratio, max_blocks_per_seq=int(lens.max() // B + 2))
    q = rng.standard_normal((S, H, D)).astype(np.float16)
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (q, kc, vc, bt, sl)]
    msl = int(bt.shape[1]) * B
    grouped = "pa_v1_gqa_kernel" in _native.plan(S, H, KVH, D, B, msl)
    assert grouped == (ratio % 2 == 0)
    kc_mha = t[1].repeat_interleave(ratio, dim=1).contiguous()        # [NB, H, ...]: head h = kv head h // ratio
    vc_mha = t[2].repeat_interleave(ratio, dim=1).contiguous()
    outs = []
    for kc_, vc_, kvh in ((t[1], t[2], KVH), (kc_mha, vc_mha, H)):
        out = torch.empty(S, H, D, dtype=torch.float16, device=dev)
        ops.paged_attention_v1(out, t[0], kc_, vc_, kvh, 0.11, t[3], t[4], B, msl, None, "auto", 1.0, 0, 0, 1, 1, 0)
        torch.cuda.synchronize()
        outs.append(out)
    del kc_mha, vc_mha
    if D == 128 and B <= 16 and ratio % 4 == 0:
        # this instantiation runs 4 warps x 1 stage (the per-head grid: 3 x 2): stages are dealt to the warps
        # differently, so the cross-warp merge rounds differently -- same values to fp32 rounding
        assert (outs[0].float() - outs[1].float()).abs().max().item() < 1e-3
    else:
        assert torch.equal(outs[0], outs[1]), "grouped-query kernel differs from the per-head grid"
    truth = c_oracle().paged_attention_v1(q, kc, vc, 0.11, bt, sl, num_kv_heads=KVH, mode="exact").astype(np.float64)
    assert_close_to_truth(outs[0].cpu().numpy(), truth, atol=1.5e-3, what="grouped-query kernel")
    # fused append through the grouped kernel == reshape_and_cache + attention (same kernel arithmetic)
    hist = np.maximum(lens.astype(np.int64) - 1, 0)
    lens1 = (hist + 1).astype(np.int32)
    slots = torch.from_numpy(bt[np.arange(S), hist // B].astype(np.int64) * B + hist % B).to(dev)
    k_new = torch.from_numpy(rng.standard_normal((S, KVH, D)).astype(np.float16)).to(dev)
    v_new = torch.from_numpy(rng.standard_normal((S, KVH, D)).astype(np.float16)).to(dev)
    sl1 = torch.from_numpy(lens1).to(dev)
    kc_a, vc_a, kc_b, vc_b = t[1].clone(), t[2].clone(), t[1].clone(), t[2].clone()
    out_a, out_b = torch.empty_like(outs[0]), torch.empty_like(outs[0])
    ops.reshape_and_cache(k_new, v_new, kc_a, vc_a, slots, "auto", 1.0)
    ops.paged_attention_v1(out_a, t[0], kc_a, vc_a, KVH, 0.11, t[3], sl1, B, msl, None, "auto", 1.0, 0, 0, 1, 1, 0)
    ops.paged_attention_v1_append(out_b, t[0], k_new, v_new, kc_b, vc_b, slots, KVH, 0.11, t[3], sl1, B, msl)
    torch.cuda.synchronize()
    assert torch.equal(kc_a, kc_b) and torch.equal(vc_a, vc_b) and torch.equal(out_a, out_b)
