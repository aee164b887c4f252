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
SOURCES = ["sf_api.cu", "sf_jacobi.cu", "sf_stages.cu", "sf_slab.cu", "sf_solvers.cu"]
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


def build(force: bool = False, verbose: bool = False, extra_flags=(), target: str = LIB) -> str:
    """One nvcc -c per source file, in parallel (sf_jacobi.cu with its ~150 kernel instantiations dominates),
    then one link step.  No relocatable device code is needed: no device function crosses a file.
    `extra_flags` / `target` build an experimental variant next to the product library (A/B timing through
    SF_LIBRARY, see solver.load_library): `python -m fluidsimulationcuda_b200.build --out build/x.so -DSF_PACKED_F32=0`."""
    if target == LIB and not extra_flags and not force and not needs_build():
        return LIB
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    base = [_nvcc()] + [f for f in NVCC_FLAGS if f != "-shared"] + list(extra_flags)
    if os.path.exists("/usr/bin/g++"):
        base += ["-ccbin", "/usr/bin/g++"]    # this image's $CC/$CXX wrapper lacks pieces the host pass needs
    if verbose:
        base += ["-Xptxas", "-v"]
    with tempfile.TemporaryDirectory(prefix="sf_build_") as tmp:
        objs = [os.path.join(tmp, s.replace(".cu", ".o")) for s in SOURCES]

        def compile_one(args):
            src, obj = args
            return subprocess.run(base + ["-c", "-o", obj, os.path.join(CSRC, src)], stdout=subprocess.PIPE,
                                  stderr=subprocess.STDOUT, text=True)
        with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
            results = list(pool.map(compile_one, zip(SOURCES, objs)))
        out = "".join(r.stdout for r in results)
        ok = all(r.returncode == 0 for r in results)
        if ok:
            os.makedirs(os.path.dirname(os.path.abspath(target)), exist_ok=True)
            link = [_nvcc(), "-shared", "-o", target] + objs
            if os.path.exists("/usr/bin/g++"):
                link += ["-ccbin", "/usr/bin/g++"]
            r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            out += r.stdout
            ok = r.returncode == 0
    if verbose or not ok:
        sys.stdout.write(out)
    if not ok:
        raise RuntimeError("nvcc failed building libstablefluids_b200.so")
    return target


if __name__ == "__main__":
    argv = sys.argv[1:]
    target = LIB
    if "--out" in argv:
        k = argv.index("--out")
        target = os.path.abspath(argv[k + 1])
        del argv[k:k + 2]
    print(build(force="--force" in argv, verbose="--verbose" in argv,
                extra_flags=[a for a in argv if a.startswith("-D")], target=target))
