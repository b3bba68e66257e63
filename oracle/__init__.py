"""oracle/ -- CPU restatement of the reference's paged-attention decode path.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from vllmini_b200/.

  paged_oracle.py   numpy restatement (exact fp64 + reference-rounding modes)
  paged_oracle.c    plain-C restatement (same three modes, OpenMP) -> _build/
  eager_attention.py torch-eager CPU port of the reference's own CPU attention
                    (vllmini/model/gpt2.py:71-78) over block-table-gathered K/V
  build_ref.py      recipe that compiles the reference's own CUDA kernels from
                    /root/reference into oracle/_ref/ (GPU-side checker)
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_C_SO = os.path.join(_BUILD, "libpaged_oracle.so")
_C_SRC = os.path.join(_HERE, "paged_oracle.c")


def build_c_oracle(force: bool = False) -> str:
    """gcc -O3 -fopenmp the C restatement into oracle/_build/ (git-ignored *.so)."""
    if (not force and os.path.isfile(_C_SO)
            and os.path.getmtime(_C_SO) >= os.path.getmtime(_C_SRC)):
        return _C_SO
    os.makedirs(_BUILD, exist_ok=True)
    cmd = ["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-o", _C_SO + ".tmp", _C_SRC, "-lm"]
    subprocess.run(cmd, check=True)
    os.replace(_C_SO + ".tmp", _C_SO)
    return _C_SO


class COracle:
    """ctypes view of libpaged_oracle.so operating on numpy arrays."""

    MODES = {"exact": 0, "fp32": 1, "ref_rounding": 2}

    def __init__(self):
        path = build_c_oracle()
        self.lib = ctypes.CDLL(path)
        L = self.lib
        vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
        L.oracle_reshape_and_cache_f16.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i64, i64]
        L.oracle_reshape_and_cache_f16.restype = None
        L.oracle_paged_attention_v1_f16.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32,
                                                    i64, i64, i64, f32, i32]
        L.oracle_paged_attention_v1_f16.restype = i32
        L.oracle_num_threads.restype = i32
        L.oracle_set_num_threads.argtypes = [i32]
        L.oracle_f2h.argtypes = [f32]
        L.oracle_f2h.restype = ctypes.c_uint16
        L.oracle_h2f.argtypes = [ctypes.c_uint16]
        L.oracle_h2f.restype = f32

    @staticmethod
    def _p(a):
        return ctypes.c_void_p(a.ctypes.data)

    def num_threads(self) -> int:
        return int(self.lib.oracle_num_threads())

    def set_num_threads(self, n: int) -> None:
        self.lib.oracle_set_num_threads(int(n))

    def reshape_and_cache(self, key, value, key_cache, value_cache, slot_mapping):
        T, H, D = key.shape
        NB, _, DX, B, x = key_cache.shape
        assert key.strides[2] == 2 and key.strides[1] == 2 * D
        assert value.strides[2] == 2 and value.strides[1] == 2 * D
        assert key_cache.flags.c_contiguous and value_cache.flags.c_contiguous
        sm = np.ascontiguousarray(slot_mapping, dtype=np.int64)
        self.lib.oracle_reshape_and_cache_f16(self._p(key), self._p(value), self._p(key_cache), self._p(value_cache),
                                              self._p(sm), T, H, D, B, x, key.strides[0] // 2, value.strides[0] // 2)

    def paged_attention_v1(self, query, key_cache, value_cache, scale, block_tables, seq_lens,
                           num_kv_heads=None, mode="exact"):
        S, H, D = query.shape
        NB, KVH, DX, B, x = key_cache.shape
        assert x == 8 and key_cache.flags.c_contiguous and value_cache.flags.c_contiguous
        assert query.strides[2] == 2 and query.strides[1] == 2 * D
        bt = np.ascontiguousarray(block_tables, dtype=np.int32)
        sl = np.ascontiguousarray(seq_lens, dtype=np.int32)
        out = np.empty((S, H, D), dtype=np.float16)
        rc = self.lib.oracle_paged_attention_v1_f16(
            self._p(out), self._p(query), self._p(key_cache), self._p(value_cache), self._p(bt), self._p(sl),
            S, H, KVH if num_kv_heads is None else num_kv_heads, D, B, bt.shape[1] if bt.ndim == 2 else 0,
            query.strides[0] // 2, key_cache.strides[0] // 2, key_cache.strides[1] // 2,
            float(scale), self.MODES[mode])
        if rc == -2:
            raise RuntimeError(f"Unsupported block size: {B}")
        if rc == -3:
            raise RuntimeError(f"Unsupported head size: {D}")
        if rc != 0:
            raise RuntimeError(f"oracle failed: rc={rc}")
        return out


_c_oracle = None


def c_oracle() -> COracle:
    global _c_oracle
    if _c_oracle is None:
        _c_oracle = COracle()
    return _c_oracle
