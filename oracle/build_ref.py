#!/usr/bin/env python
"""Build recipe for oracle/_ref: the reference's OWN kernels, compiled for B200.

TEST INFRASTRUCTURE ONLY.  Nothing under vllmini_b200/ may import, link or
execute what this produces; only tests/, __graft_entry__.smoke() and bench.py's
baseline legs use it, and only as the checker.

What it does
------------
Compiles the three source files of the reference extension *where they lie*
under /root/reference/paged_attention_ext/paged_attention_cuda/ (no source is
copied into this repository) with our own nvcc / g++ command lines and links
them into

    oracle/_ref/paged_attention_cuda.cpython-3XX-x86_64-linux-gnu.so

The reference's setup.py (paged_attention_ext/setup.py:34-38) only lists
sm_70..sm_89 SASS with no PTX, so the shipped build cannot run on a B200; the
ONLY deviation here is the -gencode list (sm_100) and dropping -march=native
(setup.py:31) because the .so is built in the authoring container and executed
on the GPU box.  All other nvcc flags are the reference's (setup.py:33-45).

oracle/_ref/ is git-ignored (it is a build artefact of reference code) but is
NOT gpurun-ignored, so the .so travels to the GPU box, which has no
/root/reference.

Modes
-----
default : compile attention_kernels.cu unmodified (all 7 head sizes x 3 block
          sizes x 3 dtypes x 3 kv dtypes x 2 sparsity = ~750 kernel
          instantiations; ~10 min of one core, -split-compile spreads it).
--trim  : compile a *transient* preprocessed view of attention_kernels.cu in
          which the `switch (head_size)` launch lists (attention_kernels.cu:
          738-766 and the v2 twin) keep only `case 64:` and `case 128:`.
          Kernel bodies are untouched; the transient file lives in a temp dir
          outside the repository and is deleted after compilation.  Used by
          __graft_entry__.build() when the full artefact is missing, to keep
          the build check to ~2 minutes.
"""
from __future__ import annotations

import argparse
import os
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("VLLMINI_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "paged_attention_ext", "paged_attention_cuda")
EXT_SUFFIX = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
SO_NAME = "paged_attention_cuda" + EXT_SUFFIX
SO_PATH = os.path.join(OUT_DIR, SO_NAME)


def ref_so_path() -> str:
    return SO_PATH


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "attention_kernels.cu"))


def _torch_paths():
    import torch  # noqa: F401  (heavy import kept local)
    from torch.utils import cpp_extension as ce

    return ce.include_paths(), ce.library_paths()


def _run(cmd, **kw):
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True, **kw)


def _trimmed_view(src_path: str, tmpdir: str) -> str:
    """Return path of a transient copy of attention_kernels.cu whose head-size
    launch switches keep only 64 and 128.  Lives outside the repo."""
    text = open(src_path).read()
    keep = {"64", "128"}

    def drop_cases(m):
        return m.group(0) if m.group(1) in keep else ""

    text2 = re.sub(
        r"[ \t]*case (\d+):\s*\n[ \t]*LAUNCH_PAGED_ATTENTION_V[12]\(\d+\);\s*\n[ \t]*break;\s*\n",
        drop_cases,
        text,
    )
    assert text2 != text, "trim pattern did not match the reference source"
    out = os.path.join(tmpdir, "attention_kernels.cu")
    with open(out, "w") as f:
        f.write(text2)
    return out


def build(trim: bool = False, jobs: int = 0, force: bool = False) -> str:
    if os.path.isfile(SO_PATH) and not force:
        print(f"[oracle/_ref] up to date: {SO_PATH}")
        return SO_PATH
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REF_SRC}")
    os.makedirs(OUT_DIR, exist_ok=True)
    inc, libs = _torch_paths()
    py_inc = sysconfig.get_paths()["include"]
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    nvcc = os.path.join(cuda_home, "bin", "nvcc")
    jobs = jobs or os.cpu_count() or 1

    tmp = tempfile.mkdtemp(prefix="vllmini_ref_build_")
    t0 = time.time()
    try:
        common_inc = [f"-I{p}" for p in inc] + [f"-I{py_inc}", f"-I{REF_SRC}", f"-I{cuda_home}/include"]
        defs = ["-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=1"]
        nvcc_flags = [
            "-O3",
            "-gencode", "arch=compute_100,code=sm_100",  # the one deviation from setup.py:34-38
            "--use_fast_math", "-lineinfo",
            "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
            "--expt-relaxed-constexpr", "-std=c++17", "-Xcompiler", "-fPIC", "-w",
            "-split-compile", str(jobs),
        ]
        attn_src = os.path.join(REF_SRC, "attention_kernels.cu")
        if trim:
            attn_src = _trimmed_view(attn_src, tmp)
        objs = []
        procs = []
        for name, src in (
            ("attention_kernels", attn_src),
            ("cache_kernels", os.path.join(REF_SRC, "cache_kernels.cu")),
        ):
            obj = os.path.join(tmp, name + ".o")
            objs.append(obj)
            cmd = [nvcc, "-c", src, "-o", obj] + nvcc_flags + defs + common_inc
            print("+", " ".join(cmd), flush=True)
            procs.append(subprocess.Popen(cmd))
        obj = os.path.join(tmp, "binding.o")
        objs.append(obj)
        _run(["g++", "-c", os.path.join(REF_SRC, "paged_attention_cuda.cpp"), "-o", obj,
              "-O3", "-w", "-fPIC", "-std=c++17"] + defs + common_inc)
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("nvcc failed on a reference source")
        link = ["g++", "-shared", "-o", SO_PATH + ".tmp"] + objs
        for l in libs:
            link += [f"-L{l}", f"-Wl,-rpath,{l}"]
        link += [f"-L{cuda_home}/lib64", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
                 "-ltorch_python", "-lcudart"]
        _run(link)
        os.replace(SO_PATH + ".tmp", SO_PATH)
        with open(os.path.join(OUT_DIR, "BUILD_INFO.txt"), "w") as f:
            f.write(f"mode={'trim(head_size in {64,128})' if trim else 'unmodified'}\n"
                    f"gencode=arch=compute_100,code=sm_100\nseconds={time.time()-t0:.0f}\n"
                    f"sources={REF_SRC}\n")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print(f"[oracle/_ref] built {SO_PATH} in {time.time()-t0:.0f}s")
    return SO_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--trim", action="store_true")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-j", type=int, default=0)
    a = ap.parse_args()
    build(trim=a.trim, jobs=a.j, force=a.force)
    sys.exit(0)
