"""CPU checks of oracle/rbgs_check.c, the CPU build the product's opt-in red-black Gauss-Seidel / SOR solver
(SF_OPT_SOLVER = SF_SOLVER_RBGS) is validated against on the GPU (tests/test_zzz_solvers_gpu.py).  Not a
reference path: the reference solves with Jacobi only."""
import numpy as np
import pytest

from oracle.pyoracle import RedBlackCheck

DT, VIS, DIFF = 0.016, 0.0025, 0.1


@pytest.fixture(scope="module")
def rb():
    return RedBlackCheck()


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def py_rbgs(N, b, x, x0, alpha, beta, iters, omega):
    """The scheme once more in numpy binary32, vectorised per colour (order-independent by construction)."""
    f = np.float32
    G = N + 2
    rows, cols = np.meshgrid(np.arange(G), np.arange(G), indexing="ij")
    inner = (rows >= 1) & (rows <= N) & (cols >= 1) & (cols <= N)
    for _ in range(iters):
        for colour in (0, 1):
            m = inner & (((rows + cols) & 1) == colour)
            nb = (np.roll(x, 1, 1) + np.roll(x, -1, 1)).astype(f)
            nb = (nb + np.roll(x, 1, 0)).astype(f)
            nb = (nb + np.roll(x, -1, 0)).astype(f)
            gs = ((x0 + (f(alpha) * nb).astype(f)).astype(f) / f(beta)).astype(f)
            if omega != 1.0:
                gs = (x + (f(omega) * (gs - x).astype(f)).astype(f)).astype(f)
            x[m] = gs[m]
        sx = -1.0 if b == 1 else 1.0
        sy = -1.0 if b == 2 else 1.0
        x[1:N + 1, 0] = f(sx) * x[1:N + 1, 1]; x[1:N + 1, N + 1] = f(sx) * x[1:N + 1, N]
        x[0, 1:N + 1] = f(sy) * x[1, 1:N + 1]; x[N + 1, 1:N + 1] = f(sy) * x[N, 1:N + 1]
        x[0, 0] = f(0.5) * (x[0, 1] + x[1, 0]); x[N + 1, 0] = f(0.5) * (x[N + 1, 1] + x[N, 0])
        x[0, N + 1] = f(0.5) * (x[0, N] + x[1, N + 1]); x[N + 1, N + 1] = f(0.5) * (x[N + 1, N] + x[N, N + 1])


@pytest.mark.parametrize("N", [1, 2, 5, 14, 30, 63])
@pytest.mark.parametrize("omega", [1.0, 1.5, 0.8])
def test_c_build_matches_numpy_statement(rb, N, omega):
    rng = np.random.default_rng(N)
    for b, (alpha, beta), iters in ((0, (1.0, 4.0), 5), (1, (0.635, 3.54), 4), (2, (41.8, 168.2), 3)):
        x = rng.uniform(-1, 1, (N + 2, N + 2)).astype(np.float32); x0 = rng.uniform(-1, 1, x.shape).astype(np.float32)
        want = x.copy(); py_rbgs(N, b, want, x0, alpha, beta, iters, omega)
        rb.rb_diffuse(N, b, x, x0, alpha, beta, iters, omega)
        assert np.array_equal(bits(x), bits(want)), (N, b, omega)


def test_dispatch_off_is_the_reference_path(rb, oracle):
    """With the solver switch off, the second compilation of the restatement is the restatement."""
    N, K = 30, 6
    rb.set_solver(0)
    a, b = oracle.init_synthetic(N, 5), rb.init_synthetic(N, 5)
    oracle.run_steps(N, 3, a, VIS, DIFF, DT, K)
    rb.run_steps(N, 3, b, VIS, DIFF, DT, K)
    for f in a:
        assert np.array_equal(bits(a[f]), bits(b[f])), f


def test_red_black_converges_faster_than_jacobi(rb, oracle):
    """Why the option exists: same number of iterations, smaller residual of the linear system."""
    N, K = 126, 40
    rng = np.random.default_rng(3)
    alpha = np.float32(DT) * np.float32(DIFF) * np.float32(N) * np.float32(N)
    beta = np.float32(1) + np.float32(4) * alpha
    x0 = rng.uniform(0, 1, (N + 2, N + 2)).astype(np.float32)
    xj = np.zeros_like(x0); oracle.diffuse(N, 0, xj, x0, float(alpha), float(beta), K)
    xg = np.zeros_like(x0); rb.rb_diffuse(N, 0, xg, x0, float(alpha), float(beta), K, 1.0)
    xs = np.zeros_like(x0); rb.rb_diffuse(N, 0, xs, x0, float(alpha), float(beta), K, 1.7)
    rj = rb.residual_sumsq(N, xj, x0, float(alpha), float(beta))
    rg = rb.residual_sumsq(N, xg, x0, float(alpha), float(beta))
    rs = rb.residual_sumsq(N, xs, x0, float(alpha), float(beta))
    # (white-noise right-hand side: plain Gauss-Seidel gains little on the high-frequency residual, over-relaxation a lot)
    assert rg < rj and rs < 0.1 * rj, (rj, rg, rs)


def test_steps_run_through_the_red_black_solver(rb, oracle):
    N, K = 30, 6
    a, b = oracle.init_synthetic(N, 5), rb.init_synthetic(N, 5)
    oracle.run_steps(N, 2, a, VIS, DIFF, DT, K)
    rb.set_solver(1, 1.0)
    try:
        rb.run_steps(N, 2, b, VIS, DIFF, DT, K)
    finally:
        rb.set_solver(0)
    assert not np.array_equal(bits(a["u"]), bits(b["u"]))            # a different scheme ...
    assert np.allclose(a["u"], b["u"], atol=5e-2) and np.all(np.isfinite(b["dens"]))   # ... for the same equations
