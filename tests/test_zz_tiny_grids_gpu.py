"""GPU parity on the smallest grids (G = 3 .. 12), where every cell touches a wall and a warp's band,
chunk and pipeline depth are all larger than the grid.  The oracle is pinned to the reference build at
G = 3, 4, 5, 8 (tests/test_oracle_pin.py, REF_CASES).  Bit-identical, through the C ABI.

This file sorts last on purpose: it was written after the round's last GPU session, so its first run
is the driver's round-end run; if a tiny grid is broken, every other parity test has reported already."""
import numpy as np
import pytest

from gpu_util import bits_equal, dev, host, mismatch_report

pytestmark = pytest.mark.gpu

DT, VIS, DIFF = 0.016, 0.0025, 0.1
TINY_N = [1, 2, 3, 4, 5, 6, 10]        # G = 4, 8, 12 take the streaming kernels, the others the generic ones


@pytest.fixture(scope="module")
def SF():
    from fluidsimulationcuda_b200 import solver
    return solver


def rnd(rng, G, lo=-1.0, hi=1.0):
    return rng.uniform(lo, hi, (G, G)).astype(np.float32)


def assert_same(got, want, name):
    assert bits_equal(got, want), mismatch_report(got, want, name)


@pytest.mark.parametrize("N", TINY_N)
def test_tiny_stage_functions(SF, oracle, N):
    G = N + 2
    rng = np.random.default_rng(N)
    s = SF.StableFluids(N)
    for b in (0, 1, 2):
        x = rnd(rng, G); want = x.copy(); oracle.set_bnd(N, b, want)
        dx = dev(x); s.set_bnd(b, dx)
        assert_same(host(dx), want, f"set_bnd N={N} b={b}")
    x, src = rnd(rng, G), rnd(rng, G)
    want = x.copy(); oracle.add_source(N, want, src, DT)
    dx = dev(x); s.add_source(dx, dev(src), DT)
    assert_same(host(dx), want, f"add_source N={N}")
    u, v = rnd(rng, G, -.3, .3), rnd(rng, G, -.3, .3)
    for b in (0, 1, 2):
        d0 = rnd(rng, G)
        want = np.zeros((G, G), np.float32); oracle.advect(N, b, want, d0, u, v, DT)
        dd = dev(np.zeros((G, G), np.float32)); s.advect(b, dd, dev(d0), dev(u), dev(v), DT)
        assert_same(host(dd), want, f"advect N={N} b={b}")
    wp, wdiv = np.zeros((G, G), np.float32), np.zeros((G, G), np.float32)
    oracle.computeDivergenceAndPressure(N, u, v, wp, wdiv)
    dp, ddiv = dev(rnd(rng, G)), dev(rnd(rng, G))
    s.computeDivergenceAndPressure(dev(u), dev(v), dp, ddiv)
    assert_same(host(dp), wp, f"divergence p N={N}"); assert_same(host(ddiv), wdiv, f"divergence div N={N}")
    p = rnd(rng, G)
    wu, wv = u.copy(), v.copy(); oracle.lastProject(N, wu, wv, p, wdiv)
    du, dv = dev(u), dev(v); s.lastProject(du, dv, dev(p), ddiv)
    assert_same(host(du), wu, f"lastProject u N={N}"); assert_same(host(dv), wv, f"lastProject v N={N}")


@pytest.mark.parametrize("T", [0, 1, 2, 3, 5, 7, 8])
@pytest.mark.parametrize("N", TINY_N)
def test_tiny_diffuse(SF, oracle, N, T):
    G = N + 2
    rng = np.random.default_rng(100 * N + T)
    s = SF.StableFluids(N, sweeps_per_launch=T)
    for b, (alpha, beta), iters in ((0, (1.0, 4.0), 1), (0, (1.0, 4.0), 20), (1, (0.635, 3.54), 7), (2, (2683.2, 10733.8), 40),
                                    (0, (107322.0, 429289.0), 9)):
        x, x0 = rnd(rng, G), rnd(rng, G)
        want = x.copy(); oracle.diffuse(N, b, want, x0, alpha, beta, iters)
        dx = dev(x); s.diffuse(b, dx, dev(x0), alpha, beta, iters)
        assert_same(host(dx), want, f"diffuse N={N} T={T} b={b} alpha={alpha} iters={iters}")


@pytest.mark.parametrize("K", [1, 2, 7, 20])
@pytest.mark.parametrize("N", TINY_N)
def test_tiny_steps(SF, oracle, N, K):
    """Three steps of the reference loop body (sources zeroed after step 0), device fields, graph replay from step 2 on."""
    rng = np.random.default_rng(10 * N + K)
    s = SF.StableFluids(N)
    names = ("dens", "dens_prev", "u", "u_prev", "v", "v_prev")
    w = oracle.init_synthetic(N, 3)
    w["u"][...] = rnd(rng, N + 2, -.2, .2); w["v"][...] = rnd(rng, N + 2, -.2, .2)
    f = {k: dev(w[k]) for k in names}
    for step in range(3):
        if step > 0:
            for k in ("dens_prev", "u_prev", "v_prev"):
                f[k].zero_()
        s.step(*[f[k] for k in names], VIS, DIFF, DT, K)
        oracle.run_steps(N, 1, w, VIS, DIFF, DT, K, first_step=step)
        for k in names:
            assert_same(host(f[k]), w[k], f"N={N} K={K} step {step} field {k}")
