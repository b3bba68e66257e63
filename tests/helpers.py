"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import paged_oracle as po  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# (name, S, H, D, B, seq_lens spec, seed) -- small enough for the CPU oracle to finish in seconds.
# seq_lens spec: int -> all equal; tuple -> explicit; ("rand", lo, hi) -> seeded uniform.
PARITY_CASES = [
    ("ref_test_shape",   1, 12, 64, 16, 3, 1),                      # tests/kernels/paged_attention.py:8-13
    ("one_token",        2, 12, 64, 16, 1, 2),
    ("exact_block",      2, 12, 64, 16, 16, 3),
    ("block_plus_one",   2, 12, 64, 16, 17, 4),
    ("stage_boundary",   3, 12, 64, 16, (31, 32, 33), 5),
    ("ragged_small",     5, 12, 64, 16, (1, 15, 40, 129, 300), 6),
    ("with_empty_seq",   3, 12, 64, 16, (0, 5, 64), 7),
    ("bs8_ragged",       4, 12, 64, 8, (7, 8, 9, 200), 8),
    ("bs32_ragged",      4, 12, 64, 32, (31, 32, 33, 250), 9),
    ("gpt2_large_heads", 3, 20, 64, 16, (100, 257, 511), 10),       # BASELINE configs[3] head shape
    ("head128_bs16",     2, 8, 128, 16, (33, 190), 11),
    ("head128_bs32",     2, 4, 128, 32, (5, 97), 12),
    ("head128_bs8",      2, 4, 128, 8, (5, 97), 13),
    ("ctx512",           4, 12, 64, 16, 512, 14),                   # BASELINE configs[1] shape, fewer seqs
    ("ctx_mixed_long",   3, 12, 64, 16, (128, 1024, 2048), 15),     # BASELINE configs[2] range
    ("many_seqs_short",  64, 12, 64, 16, ("rand", 1, 80), 16),
    # the reference launcher's other head sizes (attention_kernels.cu:738-766)
    ("head80_bs8",       3, 4, 80, 8, (1, 50, 131), 17),            # partial last chunk per lane (80*8/8 = 80 chunks)
    ("head80_bs16",      2, 4, 80, 16, (17, 140), 18),
    ("head80_bs32",      2, 4, 80, 32, (33, 140), 19),
    ("head96_bs8",       2, 4, 96, 8, (9, 99), 20),
    ("head96_bs16",      2, 4, 96, 16, (16, 300), 21),
    ("head112_bs8",      3, 2, 112, 8, (7, 64, 257), 22),           # partial last chunk per lane
    ("head112_bs32",     2, 2, 112, 32, (31, 130), 23),
    ("head192_bs8",      2, 2, 192, 8, (12, 150), 24),
    ("head192_bs16",     2, 2, 192, 16, (40, 333), 25),
    ("head192_bs32",     2, 2, 192, 32, (64, 200), 26),             # q kept in shared memory (24 chunks per lane)
    ("head256_bs8",      2, 2, 256, 8, (3, 150), 27),
    ("head256_bs16",     2, 2, 256, 16, (100, 513), 28),            # 3-warp CTA
    ("head256_bs32",     2, 2, 256, 32, (95, 289), 29),             # q kept in shared memory (32 chunks per lane)
]


def case_seq_lens(spec, S, seed):
    if isinstance(spec, int):
        return np.full(S, spec, dtype=np.int32)
    if isinstance(spec, tuple) and spec and spec[0] == "rand":
        rng = np.random.default_rng(seed + 1000)
        return rng.integers(spec[1], spec[2] + 1, size=S).astype(np.int32)
    return np.asarray(spec, dtype=np.int32)


def build_case(name):
    for c in PARITY_CASES:
        if c[0] == name:
            _, S, H, D, B, spec, seed = c
            lens = case_seq_lens(spec, S, seed)
            need = int(((lens.astype(np.int64) + B - 1) // B).sum())
            q, kc, vc, bt, sl = po.make_case(S, H, D, B, lens, num_blocks=need + 7, seed=seed,
                                             max_blocks_per_seq=int(max(((lens.max() + B - 1) // B), 1)) + 2)
            return dict(name=name, S=S, H=H, D=D, B=B, q=q, kc=kc, vc=vc, bt=bt, sl=sl, scale=float(D) ** -0.5)
    raise KeyError(name)


def to_cuda(case):
    import torch
    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(np.ascontiguousarray(case[k])).to(dev) for k in ("q", "kc", "vc", "bt", "sl")}
    return t


def run_cuda_attention(case, tensors=None, max_seq_len=None, q_override=None):
    """Call the product path (dropin -> ops -> C ABI -> CUDA) and return fp16 numpy [S, H, D]."""
    import torch
    from vllmini_b200 import ops
    t = tensors or to_cuda(case)
    q = q_override if q_override is not None else t["q"]
    out = torch.empty((case["S"], case["H"], case["D"]), dtype=torch.float16, device=q.device)
    if max_seq_len is None:
        max_seq_len = int(t["bt"].shape[1]) * case["B"]
    ops.paged_attention_v1(out, q, t["kc"], t["vc"], case["kc"].shape[1], case["scale"], t["bt"], t["sl"],
                           case["B"], max_seq_len, None, "auto", 1.0, 0, 0, 1, 1, 0)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def load_ref_extension():
    """The reference's own kernels built by oracle/build_ref.py (None if absent)."""
    import torch  # noqa: F401
    from oracle import build_ref
    path = build_ref.ref_so_path()
    if not os.path.isfile(path):
        return None
    spec = importlib.util.spec_from_file_location("paged_attention_cuda", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_ref_attention(ref, case, tensors=None):
    import torch
    t = tensors or to_cuda(case)
    out = torch.empty((case["S"], case["H"], case["D"]), dtype=torch.float16, device=t["q"].device)
    ref.paged_attention_v1(out, t["q"], t["kc"], t["vc"], case["kc"].shape[1], case["scale"], t["bt"], t["sl"],
                           case["B"], int(t["bt"].shape[1]) * case["B"], None, "auto", 1.0, 0, 0, 1, 1, 0)
    torch.cuda.synchronize()
    return out.cpu().numpy()


# Tolerances: |new - ref_kernel| <= 2e-3 + 1e-2*|ref| elementwise, and
# max|new - truth| <= max(1e-3, max|ref_kernel - truth|).
ATOL_VS_REF = 2e-3
RTOL_VS_REF = 1e-2
ATOL_VS_TRUTH = 1e-3


def assert_close_to_truth(got_f16, truth_f64, atol=ATOL_VS_TRUTH, what=""):
    err = np.abs(got_f16.astype(np.float64) - truth_f64)
    # one fp16 ulp of the value itself is unavoidable: allow atol + 2^-10 * |truth|
    bound = atol + np.abs(truth_f64) * 2.0 ** -10
    bad = err > bound
    assert not bad.any(), f"{what}: max err {err.max():.3e} (bound {bound[np.unravel_index(err.argmax(), err.shape)]:.3e})"
    return float(err.max())
