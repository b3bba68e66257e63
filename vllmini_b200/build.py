"""Build libvllmini_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m vllmini_b200.build [--force] [--verbose]

The .so lands in vllmini_b200/lib/ (git-ignored, but shipped to the GPU box by gpurun).
nvcc cross-compiles without a GPU.  No torch involvement: the library only needs the CUDA
runtime (linked statically), so it can be bound from any host language.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libvllmini_b200.so")
SOURCES = ["abi.cu", "cache_kernels.cu", "cache_aux.cu", "model_aux.cu", "paged_attention.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(ROOT, "include", "vllmini_b200.h")]


def _nvcc() -> str:
    home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cand = os.path.join(home, "bin", "nvcc")
    return cand if os.path.isfile(cand) else "nvcc"


def needs_build() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


EXP_LIB_PATH = os.path.join(LIB_DIR, "experiments", "libvllmini_b200.so")


def variant_path(name: str) -> str:
    return os.path.join(LIB_DIR, name, "libvllmini_b200.so")


def build(force: bool = False, verbose: bool = False, experiments: bool = False, variant: str = "",
          defines=()) -> str:
    """Default: the shipped library (vllmini_b200/lib/libvllmini_b200.so).

    experiments=True builds the sweep variant (-DVB200_EXPERIMENTS: extra CTA shapes, forced split factor, merge
    mode and cache-kernel knobs behind vb200_debug_* entries) to lib/experiments/; `variant` + `defines` build any
    other development variant to lib/<variant>/ (e.g. "trace" with -DVB200_TRACE for tools/trace_timeline.py).
    Tools load a variant through VLLMINI_B200_LIB; the product always loads the default build."""
    if experiments:
        variant, defines = "experiments", ("-DVB200_EXPERIMENTS",) + tuple(defines)
    out_path = variant_path(variant) if variant else LIB_PATH
    if variant:
        deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
        if not force and os.path.isfile(out_path) and all(os.path.getmtime(d) <= os.path.getmtime(out_path) for d in deps):
            return out_path
    elif not force and not needs_build():
        return LIB_PATH
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    cmd = [
        _nvcc(), "-shared", "-o", out_path + ".tmp",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-lineinfo", "-std=c++17",
        "-Xcompiler", "-fPIC,-O3,-Wall",
        "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
        "-t", "0",
        # smaller fatbin (2.7 MB instead of 5.6 MB); needs an R570+ driver to load, which every sm_100 box has
        "--compress-mode=size",
    ]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += list(defines)
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    os.replace(out_path + ".tmp", out_path)
    return out_path


SHIM_SRC = os.path.join(CSRC, "torch_shim.cpp")


def shim_path() -> str:
    import sysconfig
    return os.path.join(PKG, "_torch_shim" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_shim(force: bool = False) -> str:
    """Compile the torch-facing pybind11 shim (g++ only; links libvllmini_b200.so and torch).  It is an
    optimisation of call overhead, not a different path: without it ops.py binds the same C ABI through
    ctypes."""
    out = shim_path()
    deps = [SHIM_SRC, os.path.join(ROOT, "include", "vllmini_b200.h")]
    if not force and os.path.isfile(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    import sysconfig

    from torch.utils import cpp_extension as ce
    build(force=False)
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-shared", "-fPIC", "-O2", "-std=c++17", "-w", SHIM_SRC, "-o", out + ".tmp",
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-DTORCH_EXTENSION_NAME=_torch_shim", "-D_GLIBCXX_USE_CXX11_ABI=1",
           "-I" + os.path.join(ROOT, "include"), "-I" + sysconfig.get_paths()["include"], "-I" + cuda_home + "/include"]
    cmd += ["-I" + p for p in ce.include_paths()]
    for l in ce.library_paths():
        cmd += ["-L" + l, "-Wl,-rpath," + l]
    cmd += ["-L" + LIB_DIR, "-Wl,-rpath,$ORIGIN/lib", "-lvllmini_b200", "-lc10", "-lc10_cuda", "-ltorch_cpu",
            "-ltorch", "-ltorch_python"]
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    os.replace(out + ".tmp", out)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--experiments", action="store_true", help="also build lib/experiments/ (sweep knobs)")
    ap.add_argument("--variant", default="", help="also build lib/<variant>/ with the given --define flags")
    ap.add_argument("--define", action="append", default=[], help="e.g. --define=-DVB200_TRACE")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
    if a.experiments:
        print(build(force=a.force, verbose=a.verbose, experiments=True))
    if a.variant:
        print(build(force=a.force, verbose=a.verbose, variant=a.variant, defines=a.define))
    print(build_shim(force=a.force))
    sys.exit(0)
