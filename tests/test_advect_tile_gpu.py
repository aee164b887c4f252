"""advect with the source tile staged in shared memory by the TMA unit (SF_OPT_ADVECT_TILE, advect_tile_kernel in
csrc/sf_stages.cu) against the CPU oracle -- bit-identical, like every STRICT path (FluidSequential.c:107-141).

The tile path only exists for (N+2) % 4 == 0 and N+2 >= 320; each case checks through the counters that the tiles it means to
exercise (TMA box / gather fallback) really ran."""
import numpy as np
import pytest

from gpu_util import bits_equal, dev, host, mismatch_report

pytestmark = pytest.mark.gpu

DT = 0.016


@pytest.fixture(scope="module")
def SF():
    from fluidsimulationcuda_b200 import solver
    return solver


def assert_same(got, want, name):
    assert bits_equal(got, want), mismatch_report(got, want, name)


def smooth_noise(rng, G, radius):
    """white noise averaged over (2 radius + 1)^2 cells: the kind of velocity an unconverged viscosity solve leaves"""
    a = rng.uniform(-1.0, 1.0, (G + 2 * radius, G + 2 * radius))
    c = np.cumsum(np.cumsum(a, 0), 1)
    c = np.pad(c, ((1, 0), (1, 0)))
    k = 2 * radius + 1
    s = c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]
    return (s[:G, :G] / (k * k)).astype(np.float64)


def velocity(rng, G, N, kind):
    """cells of back-trace: a drift plus a jitter"""
    cells = 1.0 / (DT * N)      # velocity that moves a trace by one cell
    if kind == "drift":         # ~65 cells of drift, +-3 cells of smooth jitter: every tile fits its box
        u = 65.0 + 3.0 * smooth_noise(rng, G, 4) / 0.12
        v = -40.0 + 3.0 * smooth_noise(rng, G, 4) / 0.12
    elif kind == "walls":       # traces leave the grid on every side: the clamps and the zero-filled part of the box
        yy, xx = np.mgrid[0:G, 0:G]
        u = (xx - G / 2) * 0.9 + 2.0 * smooth_noise(rng, G, 3)
        v = (yy - G / 2) * -0.9 + 2.0 * smooth_noise(rng, G, 3)
        u, v = np.clip(u, -200, 200) * 0.15, np.clip(v, -200, 200) * 0.15
    elif kind == "rough":       # white noise of +-40 cells: no tile fits, every CTA takes the gather fallback
        u = rng.uniform(-40.0, 40.0, (G, G))
        v = rng.uniform(-40.0, 40.0, (G, G))
    elif kind == "mixed":       # smooth on the left half, rough on the right
        u = 20.0 + 2.0 * smooth_noise(rng, G, 3)
        v = 20.0 + 2.0 * smooth_noise(rng, G, 3)
        u[:, G // 2:] += rng.uniform(-30.0, 30.0, (G, G - G // 2))
    elif kind == "tall":        # row jitter just around the 64-row limit of the box
        u = 5.0 + np.zeros((G, G))
        v = 12.0 * np.sign(smooth_noise(rng, G, 6)) + 3.0 * smooth_noise(rng, G, 2)
    else:
        raise ValueError(kind)
    return (u * cells).astype(np.float32), (v * cells).astype(np.float32)


def tiles_per_launch(N, tile, nf=2):
    """CTAs of one advect launch: 128 columns x 32 rows (2..8, and the default for one field) or 16 rows (12..18, and the
    default for the u, v pair)"""
    G = N + 2
    rows = 32 if (2 <= tile <= 8 or (tile == 1 and nf == 1)) else 16
    return ((G + 127) // 128) * ((N + rows - 1) // rows)


def counters(s, SF):
    return s.get_option(SF.SF_OPT_ADVECT_TILE_COUNT), s.get_option(SF.SF_OPT_ADVECT_FALLBACK_COUNT)


CASES = [(318, "drift"), (510, "drift"), (1022, "drift"), (574, "walls"), (510, "rough"), (766, "mixed"), (510, "tall"),
         (1150, "walls")]


@pytest.mark.parametrize("N,kind", CASES)
@pytest.mark.parametrize("tile", [1, 4, 8, 13])
def test_advect_scalar_field(SF, oracle_mt, N, kind, tile):
    G = N + 2
    rng = np.random.default_rng(7 * N + len(kind))
    s = SF.StableFluids(N)
    s.set_option(SF.SF_OPT_ADVECT_TILE, tile)
    s.set_option(SF.SF_OPT_ADVECT_TILE_COUNT, 0)
    u, v = velocity(rng, G, N, kind)
    for b in (0, 1, 2):
        d0 = rng.uniform(-1.0, 1.0, (G, G)).astype(np.float32)
        want = np.zeros((G, G), np.float32); oracle_mt.advect(N, b, want, d0, u, v, DT)
        d = dev(np.full((G, G), np.nan, np.float32)); s.advect(b, d, dev(d0), dev(u), dev(v), DT)
        assert_same(host(d), want, f"advect b={b} N={N} {kind} tile={tile}")
    tma, fallback = counters(s, SF)
    ntiles = 3 * tiles_per_launch(N, tile, 1)
    assert tma + fallback == ntiles
    if kind == "drift" and tile == 1: assert fallback == 0, (tma, fallback)
    if kind == "rough": assert tma == 0, (tma, fallback)
    if kind in ("mixed", "walls"): assert tma > 0 and (fallback > 0 or kind == "walls"), (tma, fallback)
    # the same call with the tile path off: same bits, no tile counted
    s.set_option(SF.SF_OPT_ADVECT_TILE, 0)
    d2 = dev(np.full((G, G), np.nan, np.float32)); s.advect(2, d2, dev(d0), dev(u), dev(v), DT)
    assert_same(host(d2), want, "advect, tile off")
    assert counters(s, SF) == (tma, fallback)


@pytest.mark.parametrize("N,kind", CASES)
def test_advect_velocity_pair(SF, oracle_mt, N, kind):
    """sf_advect_velocity = advect(1, u, u0, u0, v0); advect(2, v, v0, u0, v0) in one pass (FluidSequential.c:228-237)"""
    G = N + 2
    rng = np.random.default_rng(11 * N + len(kind))
    u0, v0 = velocity(rng, G, N, kind)
    wu = np.zeros((G, G), np.float32); oracle_mt.advect(N, 1, wu, u0, u0, v0, DT)
    wv = np.zeros((G, G), np.float32); oracle_mt.advect(N, 2, wv, v0, u0, v0, DT)
    for tile in (1, 3, 18, 0):
        s = SF.StableFluids(N)
        s.set_option(SF.SF_OPT_ADVECT_TILE, tile)
        s.set_option(SF.SF_OPT_ADVECT_TILE_COUNT, 0)
        du, dv = dev(np.full((G, G), np.nan, np.float32)), dev(np.full((G, G), np.nan, np.float32))
        s.advect_velocity(du, dv, dev(u0), dev(v0), DT)
        assert_same(host(du), wu, f"advect_velocity u N={N} {kind} tile={tile}")
        assert_same(host(dv), wv, f"advect_velocity v N={N} {kind} tile={tile}")
        tma, fallback = counters(s, SF)
        assert tma + fallback == (0 if tile == 0 else tiles_per_launch(N, tile))
        if tile == 1 and kind == "drift": assert fallback == 0


def test_small_and_odd_grids_keep_the_gather_kernels(SF, oracle):
    for N in (126, 254, 509):
        G = N + 2
        rng = np.random.default_rng(N)
        s = SF.StableFluids(N)
        s.set_option(SF.SF_OPT_ADVECT_TILE_COUNT, 0)
        u, v = velocity(rng, G, N, "drift")
        d0 = rng.uniform(-1.0, 1.0, (G, G)).astype(np.float32)
        want = np.zeros((G, G), np.float32); oracle.advect(N, 0, want, d0, u, v, DT)
        d = dev(np.zeros((G, G), np.float32)); s.advect(0, d, dev(d0), dev(u), dev(v), DT)
        assert_same(host(d), want, f"advect N={N}")
        assert counters(s, SF) == (0, 0)


def test_vel_step_graph_replay_with_tiles(SF, oracle_mt):
    """the tensor maps are kernel parameters: a captured step replays with them"""
    N, K, G = 510, 8, 512
    rng = np.random.default_rng(5)
    f = [rng.uniform(0.0, 1.0, (G, G)).astype(np.float32) for _ in range(4)]
    want = [a.copy() for a in f]
    s = SF.StableFluids(N)
    s.set_option(SF.SF_OPT_ADVECT_TILE_COUNT, 0)
    d = [dev(a) for a in f]
    for step in range(3):   # direct, capture, replay
        oracle_mt.vel_step(N, want[0], want[1], want[2], want[3], 0.0025, DT, K)
        s.vel_step(d[0], d[1], d[2], d[3], 0.0025, DT, K)
        for name, a, b in zip(("u", "v", "u0", "v0"), d, want):
            assert_same(host(a), b, f"vel_step {name} step {step}")
    tma, fallback = counters(s, SF)
    assert tma + fallback == 3 * tiles_per_launch(N, 1) and tma > 0


def test_automatic_mode_drops_the_tiles_where_they_do_not_fit(SF, oracle_mt):
    """SF_OPT_ADVECT_TILE = 1: a velocity field whose traces scatter beyond the box makes the captured step use the gather
    kernel (no tile counted after the first, direct step); the bits do not depend on the choice"""
    N, K, G = 510, 2, 512
    rng = np.random.default_rng(6)
    amp = 60.0 / (DT * N)
    f = [rng.uniform(-amp, amp, (G, G)).astype(np.float32) for _ in range(4)]
    want = [a.copy() for a in f]
    s = SF.StableFluids(N)
    s.set_option(SF.SF_OPT_ADVECT_TILE_COUNT, 0)
    d = [dev(a) for a in f]
    seen = []
    for step in range(4):   # direct, capture, replay, replay
        oracle_mt.vel_step(N, want[0], want[1], want[2], want[3], 0.0025, DT, K)
        s.vel_step(d[0], d[1], d[2], d[3], 0.0025, DT, K)
        for name, a, b in zip(("u", "v", "u0", "v0"), d, want):
            assert_same(host(a), b, f"vel_step {name} step {step}")
        seen.append(counters(s, SF))
    assert seen[0][0] + seen[0][1] == tiles_per_launch(N, 1) and seen[0][1] > seen[0][0], seen
    assert seen[1] == seen[0] and seen[3] == seen[0], seen


def test_nan_velocities_do_not_index_shared_memory(SF):
    """a NaN velocity passes the clamps; its tile takes the gather path, where (int)NaN = 0 like in the other kernels
    (the reference's own (int)NaN is undefined behaviour: the library's two paths are compared with each other)"""
    N, G = 510, 512
    rng = np.random.default_rng(8)
    u, v = velocity(rng, G, N, "drift")
    u[100, 200] = np.nan; v[300, 17] = np.nan; u[5, 5] = np.inf; v[400, 400] = -np.inf
    d0 = rng.uniform(-1.0, 1.0, (G, G)).astype(np.float32)
    out = []
    for tile in (0, 1, 8):
        s = SF.StableFluids(N)
        s.set_option(SF.SF_OPT_ADVECT_TILE, tile)
        d = dev(np.zeros((G, G), np.float32)); s.advect(0, d, dev(d0), dev(u), dev(v), DT)
        du, dv = dev(np.zeros((G, G), np.float32)), dev(np.zeros((G, G), np.float32))
        s.advect_velocity(du, dv, dev(u), dev(v), DT)
        out.append((host(d), host(du), host(dv)))
        if tile: assert s.get_option(SF.SF_OPT_ADVECT_FALLBACK_COUNT) >= 4
    for o in out[1:]:
        for a, b, name in zip(o, out[0], ("d", "u", "v")):
            assert_same(a, b, f"NaN velocities, field {name}")
