"""GPU parity of the OPT-IN red-black Gauss-Seidel / SOR solver (SF_OPT_SOLVER = SF_SOLVER_RBGS, SURVEY.md
section 8f-3) against the CPU build of the same scheme (oracle/rbgs_check.c): bit-identical, through the C ABI.
The reference's own scheme (Jacobi) stays the default and is what every other parity test runs.

Sorted late on purpose: written after the round's last GPU session, so its first run is the driver's."""

import numpy as np
import pytest

from gpu_util import bits_equal, dev, host, mismatch_report

pytestmark = pytest.mark.gpu

DT, VIS, DIFF = 0.016, 0.0025, 0.1


@pytest.fixture(scope="module")
def SF():
    from fluidsimulationcuda_b200 import solver
    return solver


@pytest.fixture(scope="module")
def rb():
    from oracle.pyoracle import RedBlackCheck
    return RedBlackCheck()


def rnd(rng, G, lo=-1.0, hi=1.0):
    return rng.uniform(lo, hi, (G, G)).astype(np.float32)


def assert_same(got, want, name):
    assert bits_equal(got, want), mismatch_report(got, want, name)


@pytest.mark.parametrize("omega_milli", [1000, 1500, 800])
@pytest.mark.parametrize("N", [1, 2, 5, 13, 14, 30, 126, 130, 222, 510])
def test_rbgs_diffuse(SF, rb, N, omega_milli):
    G = N + 2
    rng = np.random.default_rng(N + omega_milli)
    s = SF.StableFluids(N)
    s.set_option(SF.SF_OPT_SOLVER, SF.SOLVER_RBGS)
    s.set_option(SF.SF_OPT_SOR_OMEGA_MILLI, omega_milli)
    assert s.get_option(SF.SF_OPT_SOLVER) == SF.SOLVER_RBGS and s.get_option(SF.SF_OPT_SOR_OMEGA_MILLI) == omega_milli
    omega = float(np.float32(omega_milli) / np.float32(1000))
    for b, (alpha, beta), iters in ((0, (1.0, 4.0), 7), (1, (0.635, 3.54), 20), (2, (2683.2, 10733.8), 5),
                                    (0, (107322.0, 429289.0), 3)):
        x, x0 = rnd(rng, G), rnd(rng, G)
        want = x.copy(); rb.rb_diffuse(N, b, want, x0, alpha, beta, iters, omega)
        dx = dev(x); s.diffuse(b, dx, dev(x0), alpha, beta, iters)
        assert_same(host(dx), want, f"rbgs diffuse N={N} b={b} alpha={alpha} omega={omega}")


def test_rbgs_subnormal_and_zero_fields(SF, rb):
    N = 62; G = N + 2
    rng = np.random.default_rng(5)
    s = SF.StableFluids(N)
    s.set_option(SF.SF_OPT_SOLVER, SF.SOLVER_RBGS)
    x = (rnd(rng, G, 0, 1) * 1e-41).astype(np.float32); x0 = (rnd(rng, G, 0, 1) * 3e-39).astype(np.float32)
    x[10:20, :] = 0.0; x0[:, 30:40] = 0.0
    want = x.copy(); rb.rb_diffuse(N, 0, want, x0, 6.15, 25.6, 20, 1.0)
    dx = dev(x); s.diffuse(0, dx, dev(x0), 6.15, 25.6, 20)
    assert_same(host(dx), want, "rbgs subnormal diffuse")


@pytest.mark.parametrize("N,K,omega_milli", [(30, 6, 1000), (126, 20, 1000), (62, 10, 1700), (130, 4, 1000)])
def test_rbgs_steps(SF, rb, N, K, omega_milli):
    """Four steps of the loop body with every solve routed through the red-black scheme (graph replay from step 2 on)."""
    s = SF.StableFluids(N)
    s.set_option(SF.SF_OPT_SOLVER, SF.SOLVER_RBGS)
    s.set_option(SF.SF_OPT_SOR_OMEGA_MILLI, omega_milli)
    names = ("dens", "dens_prev", "u", "u_prev", "v", "v_prev")
    w = rb.init_synthetic(N, 11)
    f = {k: dev(w[k]) for k in names}
    rb.set_solver(1, float(np.float32(omega_milli) / np.float32(1000)))
    try:
        for step in range(4):
            if step > 0:
                for k in ("dens_prev", "u_prev", "v_prev"):
                    f[k].zero_()
            s.step(*[f[k] for k in names], VIS, DIFF, DT, K)
            rb.run_steps(N, 1, w, VIS, DIFF, DT, K, first_step=step)
            for k in names:
                assert_same(host(f[k]), w[k], f"rbgs N={N} K={K} step {step} field {k}")
    finally:
        rb.set_solver(0)


def test_switching_back_restores_the_reference_scheme(SF, oracle):
    N, K = 62, 8
    s = SF.StableFluids(N)
    rng = np.random.default_rng(1)
    x, x0 = rnd(rng, N + 2), rnd(rng, N + 2)
    want = x.copy(); oracle.diffuse(N, 1, want, x0, 0.635, 3.54, K)
    s.set_option(SF.SF_OPT_SOLVER, SF.SOLVER_RBGS)
    d1 = dev(x); s.diffuse(1, d1, dev(x0), 0.635, 3.54, K)
    assert not bits_equal(host(d1), want)
    s.set_option(SF.SF_OPT_SOLVER, SF.SOLVER_JACOBI)
    d2 = dev(x); s.diffuse(1, d2, dev(x0), 0.635, 3.54, K)
    assert_same(host(d2), want, "Jacobi after switching back")


def test_rbgs_option_validation(SF):
    s = SF.StableFluids(30)
    with pytest.raises(SF.StableFluidsError):
        s.set_option(SF.SF_OPT_SOLVER, 2)
    with pytest.raises(SF.StableFluidsError):
        s.set_option(SF.SF_OPT_SOR_OMEGA_MILLI, 2000)
    with pytest.raises(SF.StableFluidsError):
        s.set_option(SF.SF_OPT_SOR_OMEGA_MILLI, 0)
    slab = SF.StableFluids(62, row_lo=0, row_hi=32, halo=8)
    with pytest.raises(SF.StableFluidsError):
        slab.set_option(SF.SF_OPT_SOLVER, SF.SOLVER_RBGS)       # slabs: not yet
