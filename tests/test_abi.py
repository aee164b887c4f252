"""CPU-side checks of the drop-in boundary: the library builds/loads and exports every symbol that
include/stablefluids.h declares; argument validation works without a GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _lib():
    from fluidsimulationcuda_b200 import build, solver
    build.build()
    return solver.load_library(), solver


def test_header_symbols_exported():
    L, solver = _lib()
    hdr = open(os.path.join(ROOT, "include", "stablefluids.h")).read()
    declared = set(re.findall(r"\b(sf_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"sf_context", "sf_status"}
    assert declared, "no declarations parsed"
    assert declared == set(solver.ABI_SYMBOLS), declared ^ set(solver.ABI_SYMBOLS)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in stablefluids.h but not exported"


def test_create_rejects_bad_arguments():
    L, _ = _lib()
    h = C.c_void_p()
    assert L.sf_create(C.byref(h), 0, 0) == -1            # N < 1
    assert L.sf_create(None, 16, 0) == -1
    assert L.sf_create_slab(C.byref(h), 14, 0, None, 4, 2, 0) == -1   # empty slab
    assert L.sf_destroy(None) == -1
    assert L.sf_set_option(None, 1, 0) == -1


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from fluidsimulationcuda_b200.solver import StableFluids, StableFluidsError
    with pytest.raises(StableFluidsError):
        StableFluids(14)
    L, _ = _lib()
    h = C.c_void_p()
    assert L.sf_create(C.byref(h), 14, 0) == -2           # SF_ERR_CUDA


def test_product_does_not_touch_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may reference it."""
    bad = []
    for base in ("fluidsimulationcuda_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(^|\W)(import oracle|from oracle|oracle/|stam_oracle|pyoracle)", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_option_constants_match_the_header():
    """solver.py mirrors the SF_OPT_* / SF_SOLVER_* / SF_ARITH_* enumerators of stablefluids.h by value."""
    from fluidsimulationcuda_b200 import solver
    hdr = open(os.path.join(ROOT, "include", "stablefluids.h")).read()
    enums = {k: int(v) for k, v in re.findall(r"\b(SF_[A-Z_]+)\s*=\s*(-?\d+)", hdr)}
    opts = {k: v for k, v in enums.items() if k.startswith("SF_OPT_")}
    assert len(opts) >= 12 and len(set(opts.values())) == len(opts), "option numbers must be unique"
    for name, value in opts.items():
        assert getattr(solver, name) == value, name
    assert (solver.STRICT, solver.FAST) == (enums["SF_ARITH_STRICT"], enums["SF_ARITH_FAST"])
    assert (solver.SOLVER_JACOBI, solver.SOLVER_RBGS) == (enums["SF_SOLVER_JACOBI"], enums["SF_SOLVER_RBGS"])
    assert (solver.SF_SLAB_ERR_TIMEOUT, solver.SF_SLAB_ERR_REACH) == (enums["SF_SLAB_ERR_TIMEOUT"], enums["SF_SLAB_ERR_REACH"])


def test_c_example_compiles_and_links_as_c(tmp_path):
    """The header pair (stablefluids.h + stablefluids_compat.h, the reference's own function names) is plain C and the
    example links against the library -- and, with no CUDA device here, the program reports the error instead of
    computing anything (the GPU suite runs it for real)."""
    import shutil, subprocess
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else (shutil.which("gcc") or shutil.which("cc"))
    if cc is None:
        pytest.skip("no C compiler")
    _lib()
    exe = str(tmp_path / "fluid_main")
    lib = os.path.join(ROOT, "fluidsimulationcuda_b200")
    subprocess.check_call([cc, "-std=c11", "-Wall", "-Werror", "-O1", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "fluid_main.c"), "-L" + lib, "-lstablefluids_b200",
                           "-Wl,-rpath," + lib, "-o", exe])
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "14", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode != 0 and "sum(dens)" not in r.stdout, (r.returncode, r.stdout, r.stderr)
