"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, rejects bad arguments with the documented codes (paths that return before any CUDA
call), and the Python shim mirrors the reference's surface and error behaviour."""
import ctypes
import os
import re
import sys

import pytest
import torch

from helpers import ROOT

import vllmini_b200
from vllmini_b200 import _native, ops

HEADER = os.path.join(ROOT, "include", "vllmini_b200.h")


def header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    assert os.path.isfile(_native.LIB_PATH), "build with `python -m vllmini_b200.build`"
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = header_functions()
    assert {"vb200_paged_attention_v1_f16", "vb200_reshape_and_cache_f16", "vb200_last_error"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vllmini_b200.h but not exported"
    assert set(names) == set(_native.SYMBOLS), "ctypes table and header disagree"


def test_abi_version_and_counters():
    lib = _native.load()
    assert lib.vb200_abi_version() == _native.ABI_VERSION == 2
    assert _native.launch_count() >= 0
    # the shipped library carries no sweep knobs (they live in the -DVB200_EXPERIMENTS build only)
    assert not _native.has_experiments()
    with pytest.raises(RuntimeError, match="no sweep knobs"):
        _native.set_tuning(0, 0)


def test_plan_names_the_launched_instantiation():
    """vb200_paged_attention_plan is a pure function of the shape: no device needed."""
    assert _native.plan(256, 12, 12, 64, 16, 1024) == \
        "vb200::pa_v1_kernel<64,16,4,2,0> grid=(12,256,1) block=128 smem=67040"
    assert "pa_v1_kernel<64,16,4,2,1>" in _native.plan(256, 12, 12, 64, 16, 1024, fused_append=True)
    assert "grid=(12,1,16)" in _native.plan(1, 12, 12, 64, 16, 8192)                  # small batch: split-KV cluster
    assert "pa_v1_gqa_kernel<128,16,4,1,0,4> grid=(8,256,1)" in _native.plan(256, 32, 8, 128, 16, 1024)
    assert "pa_v1_kernel<128,16,3,2,0>" in _native.plan(8, 32, 8, 128, 16, 1024)      # small batch: per-head grid
    with pytest.raises(RuntimeError, match="Unsupported head size"):
        _native.plan(1, 12, 12, 72, 16, 64)


def test_split_factor_never_exceeds_the_resident_slots():
    """A split launch must fit the GPU at once (a second wave of clusters starts when the first has drained): the
    factor is capped by SMs x CTAs per SM / units -- 3 CTAs per SM at head sizes 64 / 80, 2 at 96 / 112 / 128, 1 at
    192 / 256 -- besides the 16-CTA cluster limit and the 256-token minimum per split.  (On a device the launcher
    additionally asks cudaOccupancyMaxActiveClusters; a plan query is device-free and uses the static estimate.)"""
    import itertools
    slots = {64: 3, 80: 3, 96: 2, 112: 2, 128: 2, 192: 1, 256: 1}
    for D, S, H, ctx in itertools.product(slots, (1, 2, 3, 6, 10, 16, 40, 100, 148, 149), (1, 3, 8, 12), (256, 1024, 3072, 32768)):
        plan = _native.plan(S, H, H, D, 16, ctx)
        gx, gy, gz = (int(x) for x in re.search(r"grid=\((\d+),(\d+),(\d+)\)", plan).groups())
        assert (gx, gy) == (H, S) and 1 <= gz <= 16, plan
        if gz > 1:
            assert S * H <= 148, plan                                    # split only below one unit per SM
            assert S * H * gz <= 148 * slots[D], (plan, "more CTAs than resident slots")
            assert ctx // gz >= 128, plan
    # the cases that motivated the cap (profiles/r02m_split_occupancy.md)
    assert "grid=(3,40,2)" in _native.plan(40, 3, 3, 96, 32, 3072)
    assert "grid=(10,10,1)" in _native.plan(10, 10, 10, 256, 16, 2048)
    assert "grid=(8,16,2)" in _native.plan(16, 8, 8, 128, 16, 1024)


def test_library_size_and_kernel_inventory():
    """No dead experiment ships: the product library holds the default instantiation per (head size, block size)
    [+ fused append], the grouped-query kernels, the v2 reduce kernels and the cache / driver helpers -- no
    persistent kernel, no alternative CTA shapes."""
    import shutil
    import subprocess
    assert os.path.getsize(_native.LIB_PATH) < 3 * 1024 * 1024
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    syms = subprocess.run([cuobjdump, "-elf", _native.LIB_PATH], capture_output=True, text=True).stdout
    names = set(re.findall(r"\.text\.(_ZN5vb200\w+)", syms))
    assert not [n for n in names if "persistent" in n]
    classic = [n for n in names if "12pa_v1_kernel" in n]
    assert len(classic) == 7 * 3 * 2, len(classic)
    assert any("19pa_v2_reduce_kernel" in n for n in names)


def test_library_has_sm100a_code_and_blackwell_instructions():
    """The hot kernel must be real sm_100a code: bulk async copies (UBLKCP), mbarrier transactions
    (SYNCS) and the mixed-precision FHFMA -- checked in the SASS of the built library."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run([cuobjdump, "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "_ZN5vb20012pa_v1_kernelILi64ELi16ELi4ELi2ELb0EEEvNS_8PaParamsE",
                           _native.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UBLKCP", "SYNCS.ARRIVE.TRANS64", "FHFMA", "CREDUX.MAX.F32", "LDS.128"):
        assert mnemonic in sass, mnemonic


def test_c_abi_argument_validation_without_gpu():
    lib = _native.load()
    one = ctypes.c_void_p(16)  # never dereferenced: validation fails first
    def attn(**kw):
        a = dict(num_seqs=1, num_heads=12, num_kv_heads=12, head_size=64, block_size=16, mb=4, msl=64,
                 qs=768, kbs=12 * 64 * 16, khs=64 * 16)
        a.update(kw)
        return lib.vb200_paged_attention_v1_f16(one, one, one, one, one, one, a["num_seqs"], a["num_heads"],
                                                a["num_kv_heads"], a["head_size"], a["block_size"], a["mb"], a["msl"],
                                                a["qs"], a["kbs"], a["khs"], 0.125, 0, None)
    assert attn(block_size=4) == -2 and "Unsupported block size: 4" in _native.last_error()
    assert attn(block_size=64) == -2
    assert attn(head_size=72) == -3 and "Unsupported head size: 72" in _native.last_error()
    assert attn(num_heads=12, num_kv_heads=5) == -4
    assert attn(qs=770) == -4 and "stride" in _native.last_error()
    assert attn(num_seqs=0) == 0                       # empty batch is a no-op
    assert attn(num_seqs=-1) == -4
    assert lib.vb200_reshape_and_cache_f16(one, one, one, one, one, 0, 12, 64, 16, 8, 768, 768, 0, None) == 0
    assert lib.vb200_reshape_and_cache_f16(one, one, one, one, one, 1, 12, 64, 16, 7, 768, 768, 0, None) == -4
    assert lib.vb200_reshape_and_cache_f16(None, one, one, one, one, 1, 12, 64, 16, 8, 768, 768, 0, None) == -4
    # no workspace anywhere in the ABI: split-KV merges inside a thread-block cluster (reference v1: no scratch either)
    assert not hasattr(lib, "vb200_paged_attention_v1_workspace_bytes")
    two = ctypes.cast(ctypes.c_void_p(16), ctypes.c_void_p)
    def v2(**kw):
        a = dict(block_size=16, head_size=64, msl=1024)
        a.update(kw)
        return lib.vb200_paged_attention_v2_f16(one, two, two, one, one, one, one, one, one, 1, 12, 12, a["head_size"],
                                                a["block_size"], 64, a["msl"], 768, 12 * 64 * 16, 64 * 16, 0.125, 0, None)
    assert v2(block_size=4) == -2 and v2(head_size=100) == -3 and v2(msl=0) == -4
    assert lib.vb200_add_layernorm_f16(one, one, one, None, 3, 768, one, None, 8, 768, one, one, 8, 768, 1e-5, 0,
                                       None) == -4 and "fewer source rows" in _native.last_error()


def test_dropin_surface_matches_reference_package():
    vllmini_b200.install_dropin()
    sys.modules.pop("paged_attention_cuda", None)
    import paged_attention_cuda as pac
    assert pac.__all__ == ["paged_attention_v1", "paged_attention_v2", "cache_ops"]   # reference __init__.py:4-8
    for name in ("swap_blocks", "copy_blocks", "reshape_and_cache", "reshape_and_cache_flash", "convert_fp8"):
        assert callable(getattr(pac.cache_ops, name))                                  # paged_attention_cuda.cpp:57-61
    from paged_attention_cuda import paged_attention_v1, cache_ops                     # gpt2.py:5
    assert paged_attention_v1 is ops.paged_attention_v1 and cache_ops.reshape_and_cache is ops.reshape_and_cache


def test_shim_error_behaviour_mirrors_reference():
    h = torch.float16
    q = torch.zeros(1, 12, 64, dtype=h)
    kc = torch.zeros(4, 12, 8, 16, 8, dtype=h)
    vc = torch.zeros(4, 12, 64, 16, dtype=h)
    bt = torch.zeros(1, 4, dtype=torch.int32)
    sl = torch.ones(1, dtype=torch.int32)
    tail = [None, "auto", 1.0, 0, 0, 1, 1, 0]
    # CPU tensors: the reference's device guard raises RuntimeError before anything else
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.paged_attention_v1(torch.empty_like(q), q, kc, vc, 12, 0.125, bt, sl, 16, 64, *tail)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.reshape_and_cache(q, q, kc, vc, torch.zeros(1, dtype=torch.int64), "auto", 1.0)
    # the rest of the exported surface validates like the reference
    with pytest.raises(RuntimeError, match="Invalid device combination"):       # cache_kernels.cu:37-39
        ops.swap_blocks(q, q, torch.zeros(1, 2, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.copy_blocks([kc], [vc], torch.zeros(1, 2, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.reshape_and_cache_flash(q, q, kc, vc, torch.zeros(1, dtype=torch.int64), "auto")
    with pytest.raises(RuntimeError, match="CUDA tensor"):                      # device guard first, like v1
        ops.paged_attention_v2(q, q, q, q, q, kc, vc, 12, 0.125, bt, sl, 16, 64, *tail)
    with pytest.raises(NotImplementedError):
        ops.convert_fp8(kc, kc, 1.0, "auto")


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _native.load()


def test_header_is_plain_c_and_abi_validates_arguments_from_a_c_program(tmp_path):
    """tests/c/abi_smoke.c includes include/vllmini_b200.h as C (gcc -std=c99 -pedantic), links the library and
    exercises every argument-validation path that needs no GPU."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "abi_smoke")
    src = os.path.join(ROOT, "tests", "c", "abi_smoke.c")
    libdir = os.path.dirname(_native.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                    "-L", libdir, "-lvllmini_b200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "abi_smoke: ok" in out.stdout, out.stdout + out.stderr
