#!/usr/bin/env python
"""Recipe for oracle/_ref/vllmini_host: the reference's OWN Python host, staged unmodified for the GPU box.

TEST INFRASTRUCTURE ONLY (sibling of build_ref.py).  Nothing under vllmini_b200/ imports what this stages.

Why
---
The north star is a drop-in: `vllmini/model/gpt2.py:5` does `from paged_attention_cuda import paged_attention_v1,
cache_ops` and the rest of the host -- `kv_cache.py`, `block_manager.py`, `scheduler.py` -- stays untouched.  The
proof is to RUN that untouched host on a B200 twice, once with `paged_attention_cuda` resolving to the reference's
own kernels (oracle/_ref/*.so, build_ref.py) and once resolving to this project's drop-in package, and compare
(tests/test_reference_host_gpu.py).  /root/reference does not exist on the GPU box, and reference sources must not
enter this repository's history, so -- exactly like the reference-derived .so -- the host files travel as a
git-ignored artefact under oracle/_ref/ (git-ignored, NOT gpurun-ignored).

What
----
Byte-for-byte copies of

    vllmini/model/gpt2.py                               the model; the two call sites :81, :94
    vllmini/model/helpers/generate_triangular_mask.py   prefill mask (scheduler.py:36)
    vllmini/kv_cache.py, block_manager.py, scheduler.py the KV bookkeeping and the scheduler
    vllmini/tests/kernels/paged_attention.py            the reference's kernel acceptance test

into oracle/_ref/vllmini_host/vllmini/... .  oracle/ref_host_manifest.json (committed: names, sizes and SHA-256
digests only, no source text) lets the GPU-side test prove the staged files are the unmodified reference files.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("VLLMINI_REFERENCE_ROOT", "/root/reference")
STAGE_DIR = os.path.join(HERE, "_ref", "vllmini_host")
MANIFEST = os.path.join(HERE, "ref_host_manifest.json")
FILES = [
    "vllmini/model/gpt2.py",
    "vllmini/model/helpers/generate_triangular_mask.py",
    "vllmini/kv_cache.py",
    "vllmini/block_manager.py",
    "vllmini/scheduler.py",
    "vllmini/tests/kernels/paged_attention.py",
]


def _sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def reference_available() -> bool:
    return all(os.path.isfile(os.path.join(REF_ROOT, f)) for f in FILES)


def staged_dir() -> str:
    return STAGE_DIR


def load_manifest() -> dict:
    with open(MANIFEST) as f:
        return json.load(f)


def verify_staged() -> bool:
    """True when every staged file exists and has the digest recorded in the committed manifest."""
    if not os.path.isfile(MANIFEST):
        return False
    man = load_manifest()["files"]
    for rel, meta in man.items():
        p = os.path.join(STAGE_DIR, rel)
        if not os.path.isfile(p) or _sha256(p) != meta["sha256"]:
            return False
    return True


def stage(write_manifest: bool = False) -> str:
    """Copy the files from the reference tree (must be present).  write_manifest=True refreshes the committed
    digests (authoring container only); otherwise the copies are checked against the committed manifest."""
    if not reference_available():
        raise RuntimeError(f"reference host files not found under {REF_ROOT}")
    entries = {}
    for rel in FILES:
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(STAGE_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        entries[rel] = {"bytes": os.path.getsize(src), "sha256": _sha256(src)}
    if write_manifest or not os.path.isfile(MANIFEST):
        with open(MANIFEST, "w") as f:
            json.dump({"source": "MDK8888/vllmini @ /root/reference (commit bb64268c, .SUBMODULES.json:3)",
                       "note": "digests of the reference host files staged by oracle/stage_ref_host.py; "
                               "no reference source text is stored in this repository",
                       "files": entries}, f, indent=1)
            f.write("\n")
    elif load_manifest()["files"] != entries:
        raise RuntimeError("reference host files differ from oracle/ref_host_manifest.json "
                           "(re-run with --write-manifest if the reference really changed)")
    return STAGE_DIR


if __name__ == "__main__":
    print(stage(write_manifest="--write-manifest" in sys.argv))
    print("verified" if verify_staged() else "NOT verified")
