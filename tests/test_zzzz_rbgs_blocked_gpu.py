"""GPU parity of SF_OPT_RBGS_BLOCKED: the opt-in red-black Gauss-Seidel / SOR solver on the temporally blocked streaming
pipeline (three iterations per launch) must give the bits of the in-place scheme (oracle/rbgs_check.c) -- the design is
checked on the CPU in tools/models/rbgs_blocked_model.py.  Off by default; sorted last: first run = the driver's."""
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


def blocked(SF, N, omega_milli=1000, chunk=0):
    s = SF.StableFluids(N)
    s.set_option(SF.SF_OPT_SOLVER, SF.SOLVER_RBGS)
    s.set_option(SF.SF_OPT_SOR_OMEGA_MILLI, omega_milli)
    s.set_option(SF.SF_OPT_RBGS_BLOCKED, 1)
    s.set_option(SF.SF_OPT_CHUNK_ROWS, chunk)
    assert s.get_option(SF.SF_OPT_RBGS_BLOCKED) == 1
    return s


@pytest.mark.parametrize("omega_milli", [1000, 1500])
@pytest.mark.parametrize("N,chunk", [(2, 0), (6, 0), (10, 0), (14, 0), (30, 0), (126, 0), (222, 16), (226, 0), (254, 24), (510, 0), (1022, 0)])
def test_blocked_rbgs_diffuse(SF, rb, N, chunk, omega_milli):
    G = N + 2
    rng = np.random.default_rng(N + omega_milli)
    s = blocked(SF, N, omega_milli, chunk)
    omega = float(np.float32(omega_milli) / np.float32(1000))
    for b, (alpha, beta), iters in ((0, (1.0, 4.0), 7), (1, (0.635, 3.54), 20), (2, (2683.2, 10733.8), 5),
                                    (0, (107322.0, 429289.0), 3), (1, (0.635, 3.54), 1), (2, (1.0, 4.0), 2)):
        x, x0 = rnd(rng, G), rnd(rng, G)
        want = x.copy(); rb.rb_diffuse(N, b, want, x0, alpha, beta, iters, omega)
        dx = dev(x); s.diffuse(b, dx, dev(x0), alpha, beta, iters)
        assert_same(host(dx), want, f"blocked rbgs N={N} b={b} alpha={alpha} iters={iters} omega={omega}")


def test_blocked_rbgs_subnormal_front(SF, rb):
    """A compactly supported field decaying into the subnormal range: the strict kernel's guarded ticks and restarts."""
    N = 254; G = N + 2
    rng = np.random.default_rng(5)
    s = blocked(SF, N)
    x = np.zeros((G, G), np.float32); x0 = np.zeros((G, G), np.float32)
    x0[100:140, 90:150] = rnd(rng, G, 0, 1)[100:140, 90:150] * np.float32(1e-25)
    alpha, beta = 107322.0, 429289.0
    want = x.copy(); rb.rb_diffuse(N, 0, want, x0, alpha, beta, 40, 1.0)
    dx = dev(x); s.diffuse(0, dx, dev(x0), alpha, beta, 40)
    assert np.any((want != 0) & (np.abs(want) < 1e-30)), "test must exercise the low end of the division's range"
    assert_same(host(dx), want, "blocked rbgs decaying front")


@pytest.mark.parametrize("N,K,omega_milli", [(30, 6, 1000), (126, 20, 1000), (62, 10, 1700)])
def test_blocked_rbgs_steps(SF, rb, N, K, omega_milli):
    s = blocked(SF, N, omega_milli)
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
                assert_same(host(f[k]), w[k], f"blocked rbgs N={N} K={K} step {step} field {k}")
    finally:
        rb.set_solver(0)
