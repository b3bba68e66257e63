"""vllmini_b200 -- B200-native (sm_100a) paged-attention decode path for MDK8888/vllmini.

Scope: the two native operations vllmini's Python host calls (vllmini/model/gpt2.py:81, :94):
`cache_ops.reshape_and_cache` and `paged_attention_v1`, hand-written in CUDA behind a C ABI
(include/vllmini_b200.h), plus the Python shim that reproduces the reference extension's
interface (`vllmini_b200.ops`, `vllmini_b200/dropin/paged_attention_cuda`).
"""
import os as _os
import sys as _sys

__version__ = "0.1.0"

_DROPIN_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "dropin")


def install_dropin() -> str:
    """Make `import paged_attention_cuda` resolve to this project's drop-in package."""
    if _DROPIN_DIR not in _sys.path:
        _sys.path.insert(0, _DROPIN_DIR)
    return _DROPIN_DIR


def native_library_path() -> str:
    from . import _native
    return _native.LIB_PATH
