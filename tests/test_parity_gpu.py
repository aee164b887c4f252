"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle.

Bar (north star / SURVEY.md section 8c): STRICT arithmetic is BIT-IDENTICAL to the reference's
sequential path for every field, every stage, every step.  FAST arithmetic (opt-in) is held to
rel-L2 <= 1e-5 and max-abs <= 1e-5 * max|ref| per field per step."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from gpu_util import bits_equal, dev, host, mismatch_report, rel_l2

pytestmark = pytest.mark.gpu

DT, VIS, DIFF = 0.016, 0.0025, 0.1


@pytest.fixture(scope="module")
def SF():
    from fluidsimulationcuda_b200 import solver
    return solver


def rnd(rng, G, lo=-1.0, hi=1.0):
    return rng.uniform(lo, hi, (G, G)).astype(np.float32)


def assert_same(got, want, name):
    assert bits_equal(got, want), mismatch_report(got, want, name)


# ---- stage functions --------------------------------------------------------------------------
# widths: 16/32 (one band), 128, 224 (= exactly two 112-column bands), 228, 256, 512 and the
# literal-N sizes 13, 128 (G = 15 / 130: not a multiple of 4 -> generic kernels)
STAGE_N = [14, 30, 126, 222, 226, 254, 510, 13, 128]


@pytest.mark.parametrize("N", STAGE_N)
def test_set_bnd_add_source(SF, oracle, N):
    G = N + 2
    rng = np.random.default_rng(N)
    s = SF.StableFluids(N)
    for b in (0, 1, 2):
        x = rnd(rng, G); want = x.copy(); oracle.set_bnd(N, b, want)
        dx = dev(x); s.set_bnd(b, dx); assert_same(host(dx), want, f"set_bnd b={b}")
    x, src = rnd(rng, G), rnd(rng, G)
    want = x.copy(); oracle.add_source(N, want, src, DT)
    dx = dev(x); s.add_source(dx, dev(src), DT); assert_same(host(dx), want, "add_source")


@pytest.mark.parametrize("N", STAGE_N)
def test_advect_divergence_project(SF, oracle, N):
    G = N + 2
    rng = np.random.default_rng(100 + N)
    s = SF.StableFluids(N)
    amp = 3.0 / (DT * N)                      # back-traces of up to ~3 cells, clamps hit at the walls
    for b in (0, 1, 2):
        d0, u, v = rnd(rng, G), rnd(rng, G, -amp, amp), rnd(rng, G, -amp, amp)
        want = np.zeros((G, G), np.float32); oracle.advect(N, b, want, d0, u, v, DT)
        dd = dev(rnd(rng, G)); s.advect(b, dd, dev(d0), dev(u), dev(v), DT)
        assert_same(host(dd), want, f"advect b={b}")
    u, v = rnd(rng, G), rnd(rng, G)
    p, div = rnd(rng, G), rnd(rng, G)
    wp, wd = p.copy(), div.copy(); oracle.computeDivergenceAndPressure(N, u, v, wp, wd)
    dp, dd = dev(p), dev(div); s.computeDivergenceAndPressure(dev(u), dev(v), dp, dd)
    assert_same(host(dd), wd, "divergence"); assert_same(host(dp), wp, "pressure zero")
    u, v, p = rnd(rng, G), rnd(rng, G), rnd(rng, G)
    wu, wv = u.copy(), v.copy(); oracle.lastProject(N, wu, wv, p, np.zeros_like(p))
    du, dv = dev(u), dev(v); s.lastProject(du, dv, dev(p), dev(p))
    assert_same(host(du), wu, "lastProject u"); assert_same(host(dv), wv, "lastProject v")


# ---- lin_solve: every temporal-blocking depth, boundary kind, arithmetic path -----------------
@pytest.mark.parametrize("T", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("N,chunk", [(30, 0), (126, 0), (222, 16), (254, 24), (510, 0)])
def test_diffuse_all_depths(SF, oracle, N, chunk, T):
    G = N + 2
    rng = np.random.default_rng(1000 * T + N)
    s = SF.StableFluids(N, sweeps_per_launch=T)
    s.set_option(SF.SF_OPT_CHUNK_ROWS, chunk)
    for b, (alpha, beta), iters in ((0, (1.0, 4.0), 2 * T + 1), (1, (0.635, 3.54), 3 * T), (2, (2683.2, 10733.8), T + 2),
                                    (0, (107322.0, 429289.0), 40)):
        x, x0 = rnd(rng, G), rnd(rng, G)
        want = x.copy(); oracle.diffuse(N, b, want, x0, alpha, beta, iters)
        dx = dev(x); s.diffuse(b, dx, dev(x0), alpha, beta, iters)
        assert_same(host(dx), want, f"diffuse N={N} T={T} b={b} alpha={alpha} iters={iters}")


@pytest.mark.parametrize("N", [13, 128, 30])
def test_diffuse_generic_path(SF, oracle, N):
    G = N + 2
    rng = np.random.default_rng(N)
    s = SF.StableFluids(N)
    s.set_option(SF.SF_OPT_FORCE_GENERIC, 1)
    for b, (alpha, beta), iters in ((0, (1.0, 4.0), 5), (1, (0.635, 3.54), 20), (2, (41.8, 168.2), 1)):
        x, x0 = rnd(rng, G), rnd(rng, G)
        want = x.copy(); oracle.diffuse(N, b, want, x0, alpha, beta, iters)
        dx = dev(x); s.diffuse(b, dx, dev(x0), alpha, beta, iters)
        assert_same(host(dx), want, f"generic diffuse N={N} b={b}")


def test_diffuse_subnormal_and_zero_fields(SF, oracle):
    """Density decays into the subnormal range within ~50 reference steps: no FTZ."""
    N = 62; G = N + 2
    rng = np.random.default_rng(5)
    s = SF.StableFluids(N)
    x = (rnd(rng, G, 0, 1) * 1e-41).astype(np.float32); x0 = (rnd(rng, G, 0, 1) * 3e-39).astype(np.float32)
    x[10:20, :] = 0.0; x0[:, 30:40] = 0.0
    want = x.copy(); oracle.diffuse(N, 0, want, x0, 6.15, 25.6, 20)
    dx = dev(x); s.diffuse(0, dx, dev(x0), 6.15, 25.6, 20)
    assert np.any((want != 0) & (np.abs(want) < 1.17e-38)), "test must exercise subnormals"
    assert_same(host(dx), want, "subnormal diffuse")


@pytest.mark.parametrize("N,K", [(126, 40), (254, 20), (510, 40), (222, 21)])
def test_diffuse_bulk_copy_staging(SF, oracle, N, K):
    """SF_OPT_STAGING=1: rows staged by cp.async.bulk (TMA unit) + mbarrier instead of per-lane cp.async."""
    G = N + 2
    rng = np.random.default_rng(N * K)
    s = SF.StableFluids(N)
    s.set_option(SF.SF_OPT_STAGING, 1)
    for b, (alpha, beta) in ((0, (1.0, 4.0)), (1, (2683.2, 10733.8)), (2, (0.635, 3.54))):
        x, x0 = rnd(rng, G), rnd(rng, G)
        want = x.copy(); oracle.diffuse(N, b, want, x0, alpha, beta, K)
        dx = dev(x); s.diffuse(b, dx, dev(x0), alpha, beta, K)
        assert_same(host(dx), want, f"bulk staging N={N} K={K} b={b}")


def test_diffuse_huge_values_take_the_guarded_path(SF, oracle):
    """Magnitudes beyond the fast division's validated range (|a| > 1e30) must still be exact: the
    streaming kernel detects them when a row is fetched and switches to the fully guarded tick."""
    N = 254; G = N + 2
    rng = np.random.default_rng(9)
    s = SF.StableFluids(N)
    for spot in ((100, 100), (3, 250), (200, 7)):
        x, x0 = rnd(rng, G), rnd(rng, G)
        x[spot] = 3.0e31; x0[spot[1], spot[0]] = -2.5e32; x[spot[0] + 20, spot[1] - 2] = 1.0e33
        want = x.copy(); oracle.diffuse(N, 1, want, x0, 2683.2, 10733.8, 14)
        dx = dev(x); s.diffuse(1, dx, dev(x0), 2683.2, 10733.8, 14)
        assert np.isfinite(want).all()
        assert_same(host(dx), want, f"huge values at {spot}")


@pytest.mark.parametrize("T", [3, 6, 7])
@pytest.mark.parametrize("b,chunk", [(0, 0), (1, 0), (0, 400), (1, 300)])
def test_diffuse_tiny_patches_restart_the_pipeline(SF, oracle, b, chunk, T):
    """Numerators below the fast division's range in the MIDDLE of a chunk: the strict kernel votes once per
    group of three ticks, and a failed vote restarts the warp's pipeline above the group with guarded ticks
    (b = 0 runs the work-stealing variant, b = 1 the plain one).  Patches of 1e-33 .. 1e-44 (subnormals
    included) and of zeros inside an O(1) field, several per band so that retries fail and back off."""
    N = 1022; G = N + 2
    rng = np.random.default_rng(77 + T)
    s = SF.StableFluids(N, sweeps_per_launch=T)
    s.set_option(SF.SF_OPT_CHUNK_ROWS, chunk)      # 0: chunks of 2T rows at this size; 300/400: long chunks, retries back off
    x, x0 = rnd(rng, G), rnd(rng, G)
    for (r0, r1, c0, c1, scale) in ((100, 140, 50, 400, 1e-33), (300, 700, 600, 640, 3e-38), (500, 520, 0, G, 1e-41),
                                    (800, 1000, 900, 1000, 0.0), (130, 180, 380, 420, 1e-44)):
        x[r0:r1, c0:c1] = (x[r0:r1, c0:c1] * scale).astype(np.float32)
        x0[r0:r1, c0:c1] = (x0[r0:r1, c0:c1] * scale).astype(np.float32)
    for alpha, beta, iters in ((6.15, 25.6, 2 * T), (2683.2, 10733.8, 3 * T + 1)):
        want = x.copy(); oracle.diffuse(N, b, want, x0, alpha, beta, iters)
        dx = dev(x); s.diffuse(b, dx, dev(x0), alpha, beta, iters)
        assert np.any((want != 0) & (np.abs(want) < 1.17e-38)), "test must exercise subnormals"
        assert_same(host(dx), want, f"tiny patches b={b} chunk={chunk} T={T} alpha={alpha}")


# ---- steps ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("fuse", [1, 0])
@pytest.mark.parametrize("N,K", [(30, 4), (62, 20), (126, 40), (254, 20), (130, 6), (254, 12), (510, 33)])
def test_vel_and_dens_step(SF, oracle, N, K, fuse):
    """fuse = 1 (default): add_source is formed inside the first launch of each viscosity / diffusion solve when that launch
    fuses 5-7 sweeps (K = 20: 5, K = 12: 6, K = 40 and 33: 7; K = 4 and 6 take the separate add_source kernel either way)."""
    G = N + 2
    rng = np.random.default_rng(N + K)
    s = SF.StableFluids(N)
    s.set_option(SF.SF_OPT_FUSE_SOURCES, fuse)
    assert s.get_option(SF.SF_OPT_FUSE_SOURCES) == fuse
    u, v, u0, v0 = rnd(rng, G, -.1, .1), rnd(rng, G, -.1, .1), rnd(rng, G, 0, 1), rnd(rng, G, 0, 1)
    wu, wv, wu0, wv0 = (a.copy() for a in (u, v, u0, v0))
    oracle.vel_step(N, wu, wv, wu0, wv0, VIS, DT, K)
    du, dv, du0, dv0 = (dev(a) for a in (u, v, u0, v0))
    for rep in range(2):   # second round replays the captured graph on restored inputs
        du.copy_(dev(u)); dv.copy_(dev(v)); du0.copy_(dev(u0)); dv0.copy_(dev(v0))
        s.vel_step(du, dv, du0, dv0, VIS, DT, K)
        for g, w, nm in ((du, wu, "u"), (dv, wv, "v"), (du0, wu0, "u0=p"), (dv0, wv0, "v0=div")):
            assert_same(host(g), w, f"vel_step {nm} rep{rep}")
    x, x0 = rnd(rng, G, 0, 1), rnd(rng, G, 0, 1)
    wx, wx0 = x.copy(), x0.copy(); oracle.dens_step(N, wx, wx0, wu, wv, DIFF, DT, K)
    dx, dx0 = dev(x), dev(x0); s.dens_step(dx, dx0, du, dv, DIFF, DT, K)
    assert_same(host(dx), wx, "dens_step x"); assert_same(host(dx0), wx0, "dens_step x0")


@pytest.mark.parametrize("name", ["run_N14_K40", "run_N62_K20", "run_N126_K20", "run_N126_K40"])
def test_reference_run_fixture(SF, name):
    """The reference program's own run (glibc rand() initial condition, sources zeroed after step 0)
    -- BASELINE config 1 is run_N126_K20 (G=128, 20 iterations, 100 steps) -- straight against the
    fixtures the reference build produced."""
    import torch
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    N, K = (int(v) for v in g["meta_N_K"])
    s = SF.StableFluids(N)
    dens, u, v = s.new_field(), s.new_field(), s.new_field()
    dens_prev, u_prev, v_prev = dev(g["ic_dens_prev"]), dev(g["ic_u_prev"]), dev(g["ic_v_prev"])
    done = 0
    for upto in (int(x) for x in g["steps"]):
        for k in range(done, upto):
            if k > 0:
                dens_prev.zero_(); u_prev.zero_(); v_prev.zero_()
            s.step(dens, dens_prev, u, u_prev, v, v_prev, VIS, DIFF, DT, K)
        done = upto
        torch.cuda.synchronize()
        for f, t in (("dens", dens), ("u", u), ("v", v)):
            assert_same(host(t), g[f"{f}_step{upto}"], f"{name} {f} step {upto}")


def test_step_1024_against_oracle(SF, oracle_mt):
    """BASELINE config 2 size (G=1024, 20 iterations), synthetic hash initial condition generated
    on the device and on the host from the same formula."""
    N, K = 1022, 20
    s = SF.StableFluids(N)
    f = {k: s.new_field() for k in ("dens", "dens_prev", "u", "u_prev", "v", "v_prev")}
    s.init_synthetic(3, f["dens"], f["dens_prev"], f["u"], f["u_prev"], f["v"], f["v_prev"])
    w = oracle_mt.init_synthetic(N, 3)
    for k in w:
        assert_same(host(f[k]), w[k], f"synthetic IC {k}")
    for step in range(3):
        if step > 0:
            for k in ("dens_prev", "u_prev", "v_prev"):
                f[k].zero_()
        s.step(f["dens"], f["dens_prev"], f["u"], f["u_prev"], f["v"], f["v_prev"], VIS, DIFF, DT, K)
        oracle_mt.run_steps(N, 1, w, VIS, DIFF, DT, K, first_step=step)
        for k in w:
            assert_same(host(f[k]), w[k], f"step {step} {k}")


def test_fast_mode_tolerance(SF, oracle):
    N, K = 254, 20
    s = SF.StableFluids(N, arithmetic=SF.FAST)
    w = oracle.init_synthetic(N, 11)
    f = {k: dev(a) for k, a in w.items()}
    for step in range(5):
        if step > 0:
            for k in ("dens_prev", "u_prev", "v_prev"):
                f[k].zero_()
        s.step(f["dens"], f["dens_prev"], f["u"], f["u_prev"], f["v"], f["v_prev"], VIS, DIFF, DT, K)
        oracle.run_steps(N, 1, w, VIS, DIFF, DT, K, first_step=step)
        for k in ("dens", "u", "v"):
            got, want = host(f[k]), w[k]
            e2, emax = rel_l2(got, want), float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))
            assert e2 <= 1e-5 and emax <= 1e-5, (k, step, e2, emax)


def test_step_host_matches_device_step(SF, oracle):
    N, K = 126, 8
    s = SF.StableFluids(N)
    w = oracle.init_synthetic(N, 5)
    h = {k: a.copy() for k, a in w.items()}
    for step in range(2):
        s.step_host(h["dens"], h["dens_prev"], h["u"], h["u_prev"], h["v"], h["v_prev"], VIS, DIFF, DT, K,
                    download_scratch=True)
        oracle.run_steps(N, 1, w, VIS, DIFF, DT, K, first_step=0)   # sources stay live: host passes them in
        for k in w:
            assert_same(h[k], w[k], f"step_host {k} step {step}")


def test_exact_division_is_validated_for_bench_betas(SF):
    """The 3-instruction division must pass its exhaustive 2^32-numerator check for the betas of the
    BASELINE configs (so the STRICT path is the fast one there), and for awkward ones."""
    import numpy as np
    s = SF.StableFluids(30)
    f32 = np.float32
    for N in (126, 1022, 8190, 16382, 32766):
        for coef in (0.0025, 0.1):
            a = f32(0.016) * f32(coef); a = a * f32(N); a = a * f32(N)
            beta = f32(1) + f32(4) * a
            assert s.division_check(float(beta)), (N, coef, float(beta))
    for beta in (3.0, 1.9999999, 1.0000001, 0.3, 123456.7):
        assert s.division_check(beta), beta


def test_errors_are_reported_not_fatal(SF):
    s = SF.StableFluids(30)
    x = s.new_field()
    with pytest.raises(SF.StableFluidsError):
        s.diffuse(0, x, x, 1.0, 4.0, 4)            # aliased
    with pytest.raises(SF.StableFluidsError):
        s.diffuse(3, x, s.new_field(), 1.0, 4.0, 4)  # bad b
    with pytest.raises(SF.StableFluidsError):
        s.diffuse(0, x, s.new_field(), 1.0, 4.0, 0)  # iters < 1
    s.diffuse(0, x, s.new_field(), 1.0, 4.0, 1)      # context still usable


def test_launch_counter_counts_kernels(SF):
    s = SF.StableFluids(126)
    f = [s.new_field() for _ in range(6)]
    n0 = s.launch_count
    s.step(*f, VIS, DIFF, DT, 40)
    per_step = s.launch_count - n0
    # 3 viscosity / diffusion solves of 6 launches (40 sweeps as 7,7,7,7,6,6; add_source rides in the first launch) + 2 pressure
    # solves of 5 launches (8 sweeps each, from the implicit zero guess) + div(2) + grad(2) + advect(2)
    assert per_step == 3 * 6 + 2 * 5 + 6, per_step
    s.set_option(SF.SF_OPT_FUSE_SOURCES, 0)
    n1 = s.launch_count
    s.step(*f, VIS, DIFF, DT, 40)
    assert s.launch_count - n1 == 3 * 6 + 2 * 5 + 6 + 3        # three separate add_source kernels (u, v, dens)
    s.set_option(SF.SF_OPT_FUSE_SOURCES, 1)
    n0 = s.launch_count
    s.step(*f, VIS, DIFF, DT, 40); s.step(*f, VIS, DIFF, DT, 40)   # captured + replayed
    assert s.launch_count - n0 == 2 * per_step


def test_c_example_with_reference_names(oracle, tmp_path):
    """examples/fluid_main.c drives the library through stablefluids_compat.h (the reference's own
    function names) from plain C; its density checksum must equal the oracle's."""
    import shutil, subprocess
    from conftest import ROOT
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "fluid_main")
    lib = os.path.join(ROOT, "fluidsimulationcuda_b200")
    subprocess.check_call([cc if not os.path.exists("/usr/bin/gcc") else "/usr/bin/gcc", "-O2", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "fluid_main.c"), "-L" + lib, "-lstablefluids_b200",
                           "-Wl,-rpath," + lib, "-o", exe])
    N, steps, K = 62, 3, 40
    out = subprocess.run([exe, str(N), str(steps)], stdout=subprocess.PIPE, text=True, check=True).stdout
    got = float(out.split("sum(dens)")[1])
    # same schedule on the oracle: synthetic IC seed 1, sources re-seeded with 1+z before step z>0
    w = oracle.init_synthetic(N, 1)
    for z in range(steps):
        if z > 0:
            fresh = oracle.init_synthetic(N, 1 + z)
            for k in ("dens_prev", "u_prev", "v_prev"):
                w[k][...] = fresh[k]
        oracle.run_steps(N, 1, w, VIS, DIFF, DT, K, first_step=0)
    want = float(w["dens"].astype(np.float64).sum())
    assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), (got, want)


@pytest.mark.parametrize("G,K", [(4096, 40)] + ([] if os.environ.get("SF_TEST_SKIP_FULL_SIZE") else [(8192, 40)]))
def test_full_size_step_against_threaded_oracle(SF, oracle_mt, G, K):
    """One whole step at a BASELINE-sized grid, bit for bit against the (threaded, identical) oracle.
    G=8192, K=40 is the headline configuration (BASELINE configs[2], what bench.py times): ~15 s of CPU on the
    GPU box's host cores with the OpenMP build of the oracle; SF_TEST_SKIP_FULL_SIZE=1 leaves it out."""
    N = G - 2
    s = SF.StableFluids(N)
    if G == 4096:
        s.set_option(SF.SF_OPT_WAVE_SKEW, 140110)  # unequal chunks in CTA start order (a full single-wave grid at this size): same bits
    names = ("dens", "dens_prev", "u", "u_prev", "v", "v_prev")
    f = {k: s.new_field() for k in names}
    s.init_synthetic(2, *[f[k] for k in names])
    w = oracle_mt.init_synthetic(N, 2)
    s.step(*[f[k] for k in names], VIS, DIFF, DT, K)
    oracle_mt.run_steps(N, 1, w, VIS, DIFF, DT, K)
    for k in names:
        assert_same(host(f[k]), w[k], f"G={G} {k}")


def test_full_size_properties_8192(SF):
    """Size-independent properties at the headline size (G=8192, K=40), no oracle needed:
    (1) determinism / graph replay: the same step from the same state twice gives identical bits;
    (2) the temporal-blocking depth does not change a single bit (T = 1 launch-per-sweep vs T = 8);
    (3) set_bnd is idempotent on the solver's outputs and the walls obey the mirror rule."""
    import torch
    N, K = 8190, 40
    names = ("dens", "dens_prev", "u", "u_prev", "v", "v_prev")
    outs = []
    for T in (8, 8, 3):
        s = SF.StableFluids(N, sweeps_per_launch=T)
        f = {k: s.new_field() for k in names}
        s.init_synthetic(4, *[f[k] for k in names])
        for rep in range(2):          # second call replays the captured graph
            s.init_sources(4 + rep, f["dens_prev"], f["u_prev"], f["v_prev"])
            s.step(*[f[k] for k in names], VIS, DIFF, DT, K)
        torch.cuda.synchronize()
        outs.append({k: f[k].clone() for k in ("dens", "u", "v")})
        if T == 3:
            u = f["u"]
            assert torch.equal(u[1:-1, 0], -u[1:-1, 1]) and torch.equal(u[1:-1, -1], -u[1:-1, -2])   # b = 1
            assert torch.equal(u[0, 1:-1], u[1, 1:-1])
            before = u.clone(); s.set_bnd(1, u); assert torch.equal(before.view(torch.int32), u.view(torch.int32))
        s.close(); del f
    for k in ("dens", "u", "v"):
        assert torch.equal(outs[0][k].view(torch.int32), outs[1][k].view(torch.int32)), f"non-deterministic {k}"
        assert torch.equal(outs[0][k].view(torch.int32), outs[2][k].view(torch.int32)), f"blocking depth changed {k}"


# ---- device-resident driver loop (sf_run_steps) and the binary field dump -----------------------------
def test_run_steps_reference_schedule_matches_the_reference_loop(SF, oracle):
    """SF_SOURCES_REFERENCE = the reference's main loop (FluidSequential.c:289-312): sources act in step 0,
    are zeroed before every later step -- all on the device, no host round trip between steps."""
    N, K, steps = 254, 20, 5
    s = SF.StableFluids(N)
    names = ("dens", "dens_prev", "u", "u_prev", "v", "v_prev")
    w = oracle.init_synthetic(N, 11)
    f = {k: dev(w[k]) for k in names}
    s.run_steps(*[f[k] for k in names], VIS, DIFF, DT, K, steps, SF.SOURCES_REFERENCE)
    oracle.run_steps(N, steps, w, VIS, DIFF, DT, K)
    import torch
    torch.cuda.synchronize()
    for k in names:
        assert bits_equal(host(f[k]), w[k]), mismatch_report(host(f[k]), w[k], k)


def test_run_steps_live_source_fields_and_dump(SF, oracle, tmp_path):
    N, K, steps = 126, 8, 3
    s = SF.StableFluids(N)
    names = ("dens", "dens_prev", "u", "u_prev", "v", "v_prev")
    w = oracle.init_synthetic(N, 4)
    src = {k: w[k].copy() for k in ("dens_prev", "u_prev", "v_prev")}
    f = {k: dev(w[k]) for k in names}
    dsrc = {k: dev(a) for k, a in src.items()}
    s.run_steps(*[f[k] for k in names], VIS, DIFF, DT, K, steps, SF.SOURCES_FIELDS, 0,
                dsrc["dens_prev"], dsrc["u_prev"], dsrc["v_prev"])
    for _ in range(steps):
        for k in src:
            w[k][...] = src[k]
        oracle.vel_step(N, w["u"], w["v"], w["u_prev"], w["v_prev"], VIS, DT, K)
        oracle.dens_step(N, w["dens"], w["dens_prev"], w["u"], w["v"], DIFF, DT, K)
    import torch
    torch.cuda.synchronize()
    for k in names:
        assert bits_equal(host(f[k]), w[k]), mismatch_report(host(f[k]), w[k], k)
    # binary dump: 32-byte header + the field
    path = tmp_path / "dens.sfld"
    s.dump_field(f["dens"], path)
    raw = path.read_bytes()
    hdr = np.frombuffer(raw[:32], dtype=np.int32)
    assert raw[:4] == b"SFLD" and hdr[1] == 1 and hdr[2] == N and hdr[3] == 0 and hdr[4] == N + 2
    assert bits_equal(np.frombuffer(raw[32:], dtype=np.float32).reshape(N + 2, N + 2), w["dens"])


def test_run_steps_synthetic_schedule_equals_explicit_refresh(SF):
    N, K, steps = 254, 6, 3
    names = ("dens", "dens_prev", "u", "u_prev", "v", "v_prev")
    a, b = SF.StableFluids(N), SF.StableFluids(N)
    fa = [a.new_field() for _ in names]
    fb = [b.new_field() for _ in names]
    a.init_synthetic(2, *fa); b.init_synthetic(2, *fb)
    a.run_steps(*fa, VIS, DIFF, DT, K, steps, SF.SOURCES_SYNTHETIC, 40)
    for k in range(steps):
        b.init_sources(40 + k, fb[1], fb[3], fb[5])
        b.step(*fb, VIS, DIFF, DT, K)
    import torch
    torch.cuda.synchronize()
    for x, y, n in zip(fa, fb, names):
        assert bits_equal(host(x), host(y)), n


# ---- row-level work stealing (SF_OPT_WORK_STEALING): same bits, whatever the warps steal ---------------
@pytest.mark.parametrize("N,K", [(254, 20), (1022, 20), (2046, 13)])
def test_work_stealing_is_bit_identical(SF, oracle_mt, N, K):
    import torch
    rng = np.random.default_rng(N)
    G = N + 2
    x = rnd(rng, G, 0.0, 1.0)
    x0 = rnd(rng, G, 0.0, 1.0)
    # a block of tiny values (down to subnormals): the guarded binary64 ticks make the warps that own it
    # slower than the rest, which is what the stealing evens out
    x[G // 4: G // 2, G // 3: G // 2] *= np.float32(1e-37)
    x0[G // 4: G // 2, G // 3: G // 2] *= np.float32(1e-38)
    x0[G // 2:, :] = 0.0
    x[G // 2:, :] = 0.0
    al, be = 2683.2, 10733.8
    for chunk in (0, 32):
        s = SF.StableFluids(N, use_graph=False)
        s.set_option(SF.SF_OPT_WORK_STEALING, 1 if chunk else 30)
        s.set_option(SF.SF_OPT_CHUNK_ROWS, chunk)
        dx, dx0 = dev(x), dev(x0)
        want = x.copy()
        oracle_mt.diffuse(N, 0, want, x0, al, be, K)
        for rep in range(3):          # epochs advance launch after launch
            dx.copy_(torch.from_numpy(x).cuda())
            s.diffuse(0, dx, dx0, al, be, K)
            torch.cuda.synchronize()
            assert_same(host(dx), want, f"stealing chunk={chunk} rep={rep}")
        s.close()


def test_work_stealing_full_step(SF, oracle_mt):
    import torch
    N, K = 1022, 20
    names = ("dens", "dens_prev", "u", "u_prev", "v", "v_prev")
    s = SF.StableFluids(N)
    s.set_option(SF.SF_OPT_WORK_STEALING, 1)
    w = oracle_mt.init_synthetic(N, 3)
    f = {k: dev(w[k]) for k in names}
    s.run_steps(*[f[k] for k in names], VIS, DIFF, DT, K, 4, SF.SOURCES_REFERENCE)
    oracle_mt.run_steps(N, 4, w, VIS, DIFF, DT, K)
    torch.cuda.synchronize()
    for k in names:
        assert_same(host(f[k]), w[k], k)


# ---- out-of-bounds guard: canaries around every field stay untouched ------------------------------------
@pytest.mark.parametrize("N,K", [(254, 20), (1022, 9), (128, 6)])
def test_no_write_outside_the_fields(SF, oracle_mt, N, K):
    """The six fields are carved out of one buffer with canary bands between them (compute-sanitizer is not
    available on the GPU pool): two whole steps -- graphs, work stealing, fused boundaries -- must leave
    every canary word intact and still match the oracle."""
    import torch
    G = N + 2
    cells, pad = G * G, 4096 + 4 * ((G * 3) // 4)     # pads keep every field 16-byte aligned
    names = ("dens", "dens_prev", "u", "u_prev", "v", "v_prev")
    buf = torch.full((pad + len(names) * (cells + pad),), float("nan"), dtype=torch.float32, device="cuda")
    canary = buf.view(torch.int32)
    canary.fill_(0x7FC0DEAD)
    f = {}
    for i, k in enumerate(names):
        lo = pad + i * (cells + pad)
        f[k] = buf[lo: lo + cells].view(G, G)
    w = oracle_mt.init_synthetic(N, 21)
    for k in names:
        f[k].copy_(torch.from_numpy(w[k]))
    s = SF.StableFluids(N)
    s.run_steps(*[f[k] for k in names], VIS, DIFF, DT, K, 3, SF.SOURCES_REFERENCE)
    oracle_mt.run_steps(N, 3, w, VIS, DIFF, DT, K)
    torch.cuda.synchronize()
    for k in names:
        assert_same(host(f[k]), w[k], k)
    flat = canary.cpu().numpy()
    for i in range(len(names) + 1):
        lo = i * (cells + pad)
        band = flat[lo: lo + pad]
        assert (band == 0x7FC0DEAD).all(), f"canary band {i} was written ({int((band != 0x7FC0DEAD).sum())} words)"


@pytest.mark.parametrize("use_graph", [True, False])
@pytest.mark.parametrize("N,K", [(254, 20), (510, 33), (1022, 12)])
def test_overlapped_solves_are_bit_identical(SF, oracle_mt, N, K, use_graph):
    """SF_OPT_OVERLAP_SOLVES: the u / v viscosity solves and the density's diffusion solve of sf_step on three streams
    (graph branches; with use_graph = False real streams forked and joined by events on every call) -- direct run, capture and
    replays against the oracle's sequential vel_step + dens_step (FluidSequential.c:305-306), and against the same context
    with the option off; vel_step alone (u || v) afterwards"""
    G = N + 2
    rng = np.random.default_rng(N + K)
    f = [rng.uniform(0.0, 1.0, (G, G)).astype(np.float32) for _ in range(6)]
    f[0][:] = 0.0; f[1][:] = 0.0; f[1][G // 3: G // 2, G // 3: G // 2] = 100.0      # a compact density source: work stealing
    want = [a.copy() for a in f]
    s = SF.StableFluids(N, use_graph=use_graph)
    assert s.get_option(SF.SF_OPT_OVERLAP_SOLVES) == 1
    s2 = SF.StableFluids(N, use_graph=use_graph)
    s2.set_option(SF.SF_OPT_OVERLAP_SOLVES, 0)
    d, d2 = [dev(a) for a in f], [dev(a) for a in f]
    names = ("dens", "dens_prev", "u", "u_prev", "v", "v_prev")
    for step in range(4):   # direct, capture, replay, replay
        oracle_mt.vel_step(N, want[2], want[4], want[3], want[5], VIS, DT, K)
        oracle_mt.dens_step(N, want[0], want[1], want[2], want[4], DIFF, DT, K)
        s.step(*d, VIS, DIFF, DT, K)
        s2.step(*d2, VIS, DIFF, DT, K)
        for name, a, a2, b in zip(names, d, d2, want):
            assert_same(host(a), b, f"overlapped step {name} step {step}")
            assert_same(host(a2), b, f"sequential step {name} step {step}")
    assert s.launch_count == s2.launch_count
    for step in range(3):   # sf_vel_step on its own: the two viscosity solves side by side
        oracle_mt.vel_step(N, want[2], want[4], want[3], want[5], VIS, DT, K)
        s.vel_step(d[2], d[4], d[3], d[5], VIS, DT, K)
        for k in (2, 3, 4, 5):
            assert_same(host(d[k]), want[k], f"overlapped vel_step {names[k]} step {step}")
