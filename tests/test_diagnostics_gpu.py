"""The warp-shuffle reductions behind the diagnostics entry points (csrc/sf_stages.cu: max_abs_kernel,
residual_kernel) through the C ABI: sf_reduce_max_abs, sf_reduce_max_abs_async, sf_residual_l2 -- on
full-grid contexts and on row-slab contexts (owned rows only, ghost rows excluded).

The reference has no diagnostics (it prints fields, FluidSequential.c:19-52); these are the north star's
"warp-shuffle reductions for any residual or diagnostic".  max|x| is exact (a maximum does not round);
the residual sum of squares is accumulated in binary64 and is held to 1e-11 relative against numpy
float64 (tolerance = summation order only)."""
import numpy as np
import pytest
import torch

from gpu_util import dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def SF():
    from fluidsimulationcuda_b200 import solver
    return solver


def residual_sumsq_f64(x, x0, alpha, beta, row_lo, row_hi):
    """numpy float64 statement: sum over interior cells of rows [row_lo, row_hi) of (x0 - (beta*x - alpha*nb))^2."""
    X, X0 = x.astype(np.float64), x0.astype(np.float64)
    G = X.shape[0]
    lo, hi = max(row_lo, 1), min(row_hi, G - 1)
    nb = X[lo:hi, 0:-2] + X[lo:hi, 2:] + X[lo - 1:hi - 1, 1:-1] + X[lo + 1:hi + 1, 1:-1]
    r = X0[lo:hi, 1:-1] - (float(np.float32(beta)) * X[lo:hi, 1:-1] - float(np.float32(alpha)) * nb)
    return float((r * r).sum())


@pytest.mark.parametrize("G", [16, 130, 256, 1024])
def test_max_abs_exact_full_grid(SF, G):
    rng = np.random.default_rng(G)
    s = SF.StableFluids(G - 2)
    a = ((rng.random((G, G), dtype=np.float32) - np.float32(0.5)) * np.float32(3.0))
    a[rng.integers(0, G), rng.integers(0, G)] = np.float32(-7.25)       # the maximum magnitude is a negative value
    assert s.reduce_max_abs(dev(a)) == float(np.abs(a).max())
    a[0, 0] = np.float32(9.5)                                             # ... or sits on a corner of the wall ring
    assert s.reduce_max_abs(dev(a)) == 9.5
    z = s.new_field()
    assert s.reduce_max_abs(z) == 0.0
    tiny = np.zeros((G, G), np.float32); tiny[3, 5] = np.float32(1e-42)   # subnormals are not flushed
    assert s.reduce_max_abs(dev(tiny)) == float(np.float32(1e-42))
    s.close()


def test_max_abs_propagates_nan_and_inf(SF):
    G = 64
    s = SF.StableFluids(G - 2)
    a = np.ones((G, G), np.float32)
    a[10, 20] = np.inf
    assert s.reduce_max_abs(dev(a)) == float("inf")
    a[30, 40] = np.nan
    assert np.isnan(s.reduce_max_abs(dev(a)))       # fmaxf would have dropped it
    s.close()


def test_max_abs_async_accumulates(SF):
    G = 128
    rng = np.random.default_rng(1)
    s = SF.StableFluids(G - 2)
    a, b = rng.random((G, G), dtype=np.float32), rng.random((G, G), dtype=np.float32) * np.float32(2.0)
    m = torch.zeros(1, dtype=torch.float32, device="cuda")
    s.reduce_max_abs_async(dev(a), m)
    s.reduce_max_abs_async(dev(b), m)
    torch.cuda.synchronize()
    assert float(m.item()) == float(max(a.max(), b.max()))
    s.close()


@pytest.mark.parametrize("G,lo,hi,halo", [(64, 0, 32, 8), (64, 32, 64, 8), (256, 64, 192, 4)])
def test_max_abs_and_residual_on_slab_contexts(SF, G, lo, hi, halo):
    """A slab context reduces over its OWNED rows only: values planted in the ghost rows must not count."""
    rng = np.random.default_rng(lo + hi)
    N = G - 2
    full = (rng.random((G, G), dtype=np.float32) - np.float32(0.5))
    full0 = (rng.random((G, G), dtype=np.float32) - np.float32(0.5))
    s = SF.StableFluids(N, row_lo=lo, row_hi=hi, halo=halo)
    base = lo - halo

    def local(a, poison):
        out = np.full((hi - lo + 2 * halo, G), poison, np.float32)
        r0, r1 = max(base, 0), min(hi + halo, G)
        out[r0 - base:r1 - base] = a[r0:r1]
        return out
    x = local(full, 0.0)
    ghost = np.ones(x.shape[0], bool); ghost[halo:halo + hi - lo] = False
    xm = x.copy(); xm[ghost] = np.float32(100.0)                    # larger than anything owned
    assert s.reduce_max_abs(dev(xm)) == float(np.abs(full[lo:hi]).max())
    al, be = 2.5, 11.0
    got = s.residual_sumsq(dev(x), dev(local(full0, 0.0)), al, be)
    want = residual_sumsq_f64(full, full0, al, be, lo, hi)
    assert abs(got - want) <= 1e-11 * want, (got, want)
    s.close()


@pytest.mark.parametrize("G", [16, 130, 1024])
@pytest.mark.parametrize("alpha,beta", [(1.0, 4.0), (2683.2, 10733.8), (0.635, 3.54)])
def test_residual_against_float64_numpy(SF, oracle, G, alpha, beta):
    rng = np.random.default_rng(G)
    N = G - 2
    s = SF.StableFluids(N)
    x = (rng.random((G, G), dtype=np.float32) - np.float32(0.5))
    x0 = (rng.random((G, G), dtype=np.float32) - np.float32(0.5))
    got = s.residual_sumsq(dev(x), dev(x0), alpha, beta)
    want = residual_sumsq_f64(x, x0, alpha, beta, 0, G)
    assert abs(got - want) <= 1e-11 * want, (got, want)
    # and it does what a residual is for: relaxing reduces it (oracle = the reference's Jacobi sweeps)
    xs = x.copy()
    oracle.diffuse(N, 0, xs, x0, alpha, beta, 20)
    after = s.residual_sumsq(dev(xs), dev(x0), alpha, beta)
    assert after < got
    assert abs(after - residual_sumsq_f64(xs, x0, alpha, beta, 0, G)) <= 1e-11 * max(after, 1e-300)
    s.close()
