"""CPU tests: the oracle against the reference-derived golden vectors and against itself
(numpy restatement vs independent C restatement vs torch-eager port)."""
import glob
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, PARITY_CASES, build_case
from oracle import c_oracle, eager_attention
from oracle import paged_oracle as po

SMALL = [c[0] for c in PARITY_CASES if c[0] not in ("ctx512", "ctx_mixed_long", "many_seqs_short")]


# ---------------------------------------------------------------- fp16 conversion of the C oracle
def test_c_half_conversions_match_numpy():
    co = c_oracle()
    bits = np.arange(0, 1 << 16, dtype=np.uint32).astype(np.uint16)
    f = bits.view(np.float16).astype(np.float32)
    for b in list(range(0, 1 << 16, 97)) + [0, 1, 0x3ff, 0x400, 0x7bff, 0x7c00, 0x8000, 0xfbff, 0xfc00]:
        got = co.lib.oracle_h2f(int(b))
        want = float(f[b])
        assert (np.isnan(got) and np.isnan(want)) or got == want, hex(b)
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.standard_normal(2000).astype(np.float32) * s for s in (1e-8, 1e-5, 1e-3, 1.0, 300.0, 7e4)])
    xs = np.concatenate([xs, np.array([0.0, -0.0, 65504.0, 65519.99, 65520.0, 2.0 ** -24, 2.0 ** -25, 1.5 * 2.0 ** -25,
                                       np.inf, -np.inf, 1.0009765625, 1.00048828125, 1.00146484375], dtype=np.float32)])
    with np.errstate(over="ignore"):
        want = xs.astype(np.float16).view(np.uint16)
    for x, w in zip(xs, want):
        assert co.lib.oracle_f2h(float(x)) == int(w), float(x)


# ---------------------------------------------------------------- reshape_and_cache
@pytest.mark.parametrize("B,H,D,T", [(16, 12, 64, 3), (8, 3, 64, 11), (32, 2, 128, 40), (16, 20, 64, 5)])
def test_reshape_and_cache_three_restatements_agree(B, H, D, T):
    rng = np.random.default_rng(B + H + D + T)
    NB = (T + B - 1) // B + 3
    # strided rows, like the q/k/v views of a fused qkv projection (gpt2.py:36-41): stride 3*H*D
    qkv = rng.standard_normal((T, 3 * H * D)).astype(np.float16)
    key = qkv[:, H * D:2 * H * D].reshape(T, H, D)
    value = qkv[:, 2 * H * D:].reshape(T, H, D)
    slots = rng.permutation(NB * B)[:T].astype(np.int64)
    slots[T // 2] = -1                                              # padding token is skipped
    caches = []
    for impl in ("numpy", "loops", "c"):
        kc = np.full((NB, H, D // 8, B, 8), 7.0, dtype=np.float16)
        vc = np.full((NB, H, D, B), 7.0, dtype=np.float16)
        if impl == "numpy":
            po.reshape_and_cache(key, value, kc, vc, slots)
        elif impl == "loops":
            po.reshape_and_cache_loops(key, value, kc, vc, slots)
        else:
            c_oracle().reshape_and_cache(key, value, kc, vc, slots)
        caches.append((kc, vc))
    for kc, vc in caches[1:]:
        assert np.array_equal(kc.view(np.uint16), caches[0][0].view(np.uint16))
        assert np.array_equal(vc.view(np.uint16), caches[0][1].view(np.uint16))
    # the reference test's layout property (tests/kernels/paged_attention.py:63-82), bit-exact
    kc, vc = caches[0]
    for t in range(T):
        if slots[t] < 0:
            continue
        blk, off = slots[t] // B, slots[t] % B
        for h in range(H):
            assert np.array_equal(kc[blk, h, :, off, :].reshape(D), key[t, h])
            assert np.array_equal(vc[blk, h, :, off], value[t, h])
    # untouched slots keep their old contents
    assert (kc == 7.0).sum() == kc.size - (T - 1) * H * D


# ---------------------------------------------------------------- golden: reference's own CPU attention
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN_DIR, "vanilla_*.npz"))))
@pytest.mark.parametrize("B", [8, 16, 32])
def test_oracle_matches_reference_vanilla_attention(path, B):
    """Scatter the fixture's dense K/V into paged caches (oracle reshape_and_cache, shuffled block
    table) and check every oracle mode against the reference's `_vanilla_attention` output."""
    g = np.load(path)
    q, k, v, lens, want, scale = g["q"], g["k"], g["v"], g["seq_lens"], g["out"], float(g["scale"])
    S, Lmax, H, D = k.shape
    nblk = (lens.astype(np.int64) + B - 1) // B
    NB = int(nblk.sum()) + 5
    rng = np.random.default_rng(B)
    perm = rng.permutation(NB).astype(np.int32)
    bt = np.full((S, int(nblk.max()) + 1), -1, dtype=np.int32)
    kc = np.zeros((NB, H, D // 8, B, 8), dtype=np.float16)
    vc = np.zeros((NB, H, D, B), dtype=np.float16)
    o = 0
    for s in range(S):
        bt[s, :nblk[s]] = perm[o:o + nblk[s]]
        o += nblk[s]
        L = int(lens[s])
        slots = bt[s, np.arange(L) // B].astype(np.int64) * B + np.arange(L) % B
        po.reshape_and_cache(k[s, :L], v[s, :L], kc, vc, slots)
    exact = po.paged_attention_v1_exact(q, kc, vc, scale, bt, lens)
    assert np.abs(exact - want).max() < 2e-5            # fp64 truth vs the reference's fp32 eager
    co = c_oracle()
    for mode, tol in (("exact", 1.2e-3), ("fp32", 1.2e-3), ("ref_rounding", 2.5e-3)):
        got = co.paged_attention_v1(q, kc, vc, scale, bt, lens, mode=mode).astype(np.float64)
        assert np.abs(got - want).max() < tol, mode     # fp16 output rounding (+ fp16 PV partials)
    rr = po.paged_attention_v1_ref_rounding(q, kc, vc, scale, bt, lens).astype(np.float64)
    assert np.abs(rr - want).max() < 2.5e-3
    eager = eager_attention.paged_attention_eager_cpu(torch.from_numpy(q), torch.from_numpy(kc), torch.from_numpy(vc),
                                                      torch.from_numpy(bt), torch.from_numpy(lens), scale).numpy()
    assert np.abs(eager - want).max() < 2e-5


# ---------------------------------------------------------------- numpy vs C restatements
@pytest.mark.parametrize("name", SMALL)
def test_numpy_and_c_oracles_agree(name):
    c = build_case(name)
    truth = po.paged_attention_v1_exact(c["q"], c["kc"], c["vc"], c["scale"], c["bt"], c["sl"])
    co = c_oracle()
    c_exact = co.paged_attention_v1(c["q"], c["kc"], c["vc"], c["scale"], c["bt"], c["sl"], mode="exact")
    # both round the same fp64 value to fp16: allow 1 ulp for the fp64 summation-order difference
    assert np.abs(c_exact.astype(np.float64) - truth).max() <= np.abs(truth).max() * 2.0 ** -10 + 1e-7
    c_f32 = co.paged_attention_v1(c["q"], c["kc"], c["vc"], c["scale"], c["bt"], c["sl"], mode="fp32")
    assert np.abs(c_f32.astype(np.float64) - truth).max() < 1e-3
    # reference-rounding emulations: the two restatements follow the same operation order, so they
    # may differ only through libm expf vs numpy exp (1 ulp fp32) -> at most 1 fp16 ulp
    c_rr = co.paged_attention_v1(c["q"], c["kc"], c["vc"], c["scale"], c["bt"], c["sl"], mode="ref_rounding")
    n_rr = po.paged_attention_v1_ref_rounding(c["q"], c["kc"], c["vc"], c["scale"], c["bt"], c["sl"])
    diff = np.abs(c_rr.astype(np.float64) - n_rr.astype(np.float64))
    assert diff.max() <= np.abs(truth).max() * 2.0 ** -9 + 1e-6
    # (a 1-ulp exp difference moves one whole (sequence, head) row: cases with only 4 rows may lose one)
    assert (c_rr.view(np.uint16) == n_rr.view(np.uint16)).mean() > (0.97 if c["S"] * c["H"] >= 16 else 0.94)
    assert np.abs(n_rr.astype(np.float64) - truth).max() < 3e-3
    # rows of empty sequences are zeros
    for s in np.nonzero(c["sl"] == 0)[0]:
        assert not c_exact[s].any() and not c_rr[s].any() and not truth[s].any()


def test_eager_port_matches_oracle_on_ragged_batch():
    c = build_case("ragged_small")
    truth = po.paged_attention_v1_exact(c["q"], c["kc"], c["vc"], c["scale"], c["bt"], c["sl"], add_eps=False)
    eager = eager_attention.paged_attention_eager_cpu(*(torch.from_numpy(np.ascontiguousarray(c[k])) for k in
                                                        ("q", "kc", "vc", "bt", "sl")), c["scale"]).numpy()
    assert np.abs(eager - truth).max() < 2e-5


def test_binding_level_oracle_errors():
    c = build_case("one_token")
    out = np.zeros((c["S"], c["H"], c["D"]), dtype=np.float16)
    args = [out, c["q"], c["kc"], c["vc"], c["H"], c["scale"], c["bt"], c["sl"], c["B"], 64, None, "auto", 1.0,
            0, 0, 1, 1, 0]
    po.paged_attention_v1(*args)
    assert out.any()
    bad = list(args); bad[8] = 4
    with pytest.raises(RuntimeError, match="block size"):
        po.paged_attention_v1(*bad)
    bad = list(args); bad[11] = "fp8"
    with pytest.raises(RuntimeError, match="kv cache"):
        po.paged_attention_v1(*bad)


def test_nan_garbage_beyond_seq_len_is_ignored():
    """Slots past seq_len in the last block may hold anything (uninitialised cache memory): the
    reference masks logits and zeroes V there (attention_kernels.cu:303, 420-430)."""
    c = build_case("block_plus_one")
    kc, vc = c["kc"].copy(), c["vc"].copy()
    last = c["bt"][0, 1]
    kc[last, :, :, 1:, :] = np.nan
    vc[last, :, :, 1:] = np.nan
    ref = po.paged_attention_v1_exact(c["q"], c["kc"], c["vc"], c["scale"], c["bt"], c["sl"])
    got = po.paged_attention_v1_exact(c["q"], kc, vc, c["scale"], c["bt"], c["sl"])
    assert np.array_equal(ref[0], got[0])
    for mode in ("exact", "fp32", "ref_rounding"):
        g = c_oracle().paged_attention_v1(c["q"], kc, vc, c["scale"], c["bt"], c["sl"], mode=mode)
        assert np.isfinite(g.astype(np.float32)).all(), mode


def test_uncalled_cache_ops_oracles():
    rng = np.random.default_rng(0)
    T, H, D, B, NB = 9, 2, 8, 4, 4
    key = rng.standard_normal((T, H, D)).astype(np.float16)
    val = rng.standard_normal((T, H, D)).astype(np.float16)
    kc = np.zeros((NB, B, H, D), dtype=np.float16)
    vc = np.zeros_like(kc)
    slots = np.array([0, 1, 2, 3, 4, -1, 9, 15, 12], dtype=np.int64)
    po.reshape_and_cache_flash(key, val, kc, vc, slots)
    assert np.array_equal(kc[1, 0], key[4]) and np.array_equal(vc[3, 3], val[7]) and not kc[1, 1].any()
    a, b = [kc.copy()], [vc.copy()]
    po.copy_blocks(a, b, np.array([[0, 2], [3, 1]]))
    assert np.array_equal(a[0][2], kc[0]) and np.array_equal(b[0][1], vc[3])
    dst = np.zeros_like(kc)
    po.swap_blocks(kc, dst, np.array([[1, 0]]))
    assert np.array_equal(dst[0], kc[1]) and not dst[1].any()
