"""Pin the CPU oracle (oracle/stam_oracle.c) to the reference.

(1) against the committed fixtures in tests/golden/, which were produced by the reference's own
    translation unit (tests/golden/make_golden.py) -- runs everywhere;
(2) against oracle/_ref/libref_seq_N*_K*.so directly, when those builds are present (they are
    built in the container that has /root/reference and travel to the GPU box as .so files).
Everything is compared BITWISE (uint32 views)."""
import os

import numpy as np
import pytest

from oracle.pyoracle import Oracle, ReferenceSeq
from conftest import GOLDEN

DT, VIS, DIFF = 0.016, 0.0025, 0.1


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def same(a, b):
    return np.array_equal(bits(a), bits(b))


def test_stage_fixture(oracle):
    g = np.load(os.path.join(GOLDEN, "stage_N30.npz"))
    N, K = (int(v) for v in g["meta_N_K"])
    for b in (0, 1, 2):
        x = g[f"set_bnd{b}_in"].copy(); oracle.set_bnd(N, b, x); assert same(x, g[f"set_bnd{b}_out"])
    x = g["add_source_x"].copy(); oracle.add_source(N, x, g["add_source_s"].copy(), DT)
    assert same(x, g["add_source_out"])
    for b in (0, 1, 2):
        x = g[f"diffuse{b}_x"].copy(); al, be = (float(v) for v in g[f"diffuse{b}_ab"])
        oracle.diffuse(N, b, x, g[f"diffuse{b}_x0"].copy(), al, be, K); assert same(x, g[f"diffuse{b}_out"])
    for b in (0, 1, 2):
        d = np.zeros_like(g[f"advect{b}_d0"])
        # the reference leaves corners/ring of d to set_bnd; start from the same garbage-free state
        oracle.advect(N, b, d, g[f"advect{b}_d0"].copy(), g[f"advect{b}_u"].copy(), g[f"advect{b}_v"].copy(), DT)
        assert same(d, g[f"advect{b}_out"])
    p = np.empty_like(g["div_u"]); div = np.empty_like(p)
    oracle.computeDivergenceAndPressure(N, g["div_u"].copy(), g["div_v"].copy(), p, div)
    assert same(p, g["div_p_out"]) and same(div, g["div_div_out"])
    u, v = g["lp_u"].copy(), g["lp_v"].copy()
    oracle.lastProject(N, u, v, g["lp_p"].copy(), np.zeros_like(u))
    assert same(u, g["lp_u_out"]) and same(v, g["lp_v_out"])
    x, x0 = g["dens_x"].copy(), g["dens_x0"].copy()
    oracle.dens_step(N, x, x0, g["dens_u"].copy(), g["dens_v"].copy(), DIFF, DT, K)
    assert same(x, g["dens_x_out"]) and same(x0, g["dens_x0_out"])
    u, v, u0, v0 = (g[k].copy() for k in ("vel_u", "vel_v", "vel_u0", "vel_v0"))
    oracle.vel_step(N, u, v, u0, v0, VIS, DT, K)
    assert same(u, g["vel_u_out"]) and same(v, g["vel_v_out"])
    assert same(u0, g["vel_u0_out"]) and same(v0, g["vel_v0_out"])


@pytest.mark.parametrize("name", ["run_N14_K40", "run_N62_K20", "run_N126_K20", "run_N126_K40"])
def test_run_fixture(oracle, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    N, K = (int(v) for v in g["meta_N_K"])
    s = oracle.init_reference_rand(N)
    for f in ("dens_prev", "u_prev", "v_prev"):
        assert same(s[f], g[f"ic_{f}"]), "glibc rand() initial condition differs"
    done = 0
    for upto in (int(v) for v in g["steps"]):
        oracle.run_steps(N, upto - done, s, VIS, DIFF, DT, K, first_step=done)
        done = upto
        for f in ("dens", "u", "v"):
            assert same(s[f], g[f"{f}_step{upto}"]), (name, f, upto)


def test_threaded_oracle_identical(oracle, oracle_mt):
    N, K = 62, 6
    a, b = oracle.init_synthetic(N, 7), oracle_mt.init_synthetic(N, 7)
    oracle.run_steps(N, 3, a, VIS, DIFF, DT, K)
    oracle_mt.run_steps(N, 3, b, VIS, DIFF, DT, K)
    for f in a:
        assert same(a[f], b[f]), f


def test_odd_iteration_count_lands_in_x(oracle):
    """The reference only supports even K (FluidSequential.c:100-103); the oracle's odd-K result
    must equal one more sweep applied to the even-K result."""
    N = 30
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (N + 2, N + 2)).astype(np.float32); x0 = rng.uniform(-1, 1, x.shape).astype(np.float32)
    a = x.copy(); oracle.diffuse(N, 1, a, x0, 0.7, 3.8, 4)
    b = a.copy(); oracle.diffuse(N, 1, b, x0, 0.7, 3.8, 1)
    c = x.copy(); oracle.diffuse(N, 1, c, x0, 0.7, 3.8, 5)
    assert same(b, c)


# the first four: the smallest grids (G = 3, 4, 5, 8), where every cell touches a wall
REF_CASES = [(1, 2, 4), (2, 4, 4), (3, 2, 4), (6, 4, 6), (14, 40, 5), (30, 4, 5), (62, 20, 10), (126, 20, 30), (126, 40, 10), (254, 20, 4), (1022, 20, 2)]


@pytest.mark.parametrize("N,K,steps", REF_CASES)
def test_against_reference_build(oracle, N, K, steps):
    if not ReferenceSeq.available(N, K):
        pytest.skip("oracle/_ref build not present")
    R = ReferenceSeq(N, K)
    a = R.initializeParameters()
    b = oracle.init_reference_rand(N)
    for s in range(steps):
        R.run_steps(1, a, first_step=s)
        oracle.run_steps(N, 1, b, VIS, DIFF, DT, K, first_step=s)
        for f in a:  # state AND scratch fields (u_prev = pressure, v_prev = divergence, ...)
            assert same(a[f], b[f]), (N, K, s, f)
