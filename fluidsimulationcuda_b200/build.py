"""Build the CUDA library in-tree: fluidsimulationcuda_b200/libstablefluids_b200.so (sm_100a).

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
gpurun snapshot.  `python -m fluidsimulationcuda_b200.build [--force] [--verbose]`."""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libstablefluids_b200.so")
SOURCES = ["sf_api.cu", "sf_jacobi.cu", "sf_stages.cu", "sf_slab.cu"]
HEADERS = ["sf_common.cuh", "sf_internal.h", os.path.join("..", "..", "include", "stablefluids.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                      # parity: never contract mul+add (FAST mode asks for FMA explicitly)
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]    # this image's $CC/$CXX wrapper lacks pieces the host pass needs
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        sys.stdout.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libstablefluids_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
