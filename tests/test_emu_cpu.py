"""The Jacobi kernels' OWN SOURCE, compiled for the host and run warp by warp with 32 threads per warp (tools/emu/):
stream_rows, pipeline_tick, the exact division with its votes, guarded ticks and pipeline restarts, fused set_bnd, the
red-black levels -- everything except the inline-PTX helpers, which get host equivalents.  Checked bitwise against the
oracle (Jacobi) and against the in-place red-black scheme (SF_OPT_RBGS_BLOCKED), for the default build and for the
inner-loop / guarded-group paths it contains.  No GPU needed; the GPU suite remains the final word on the SASS."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

FP = C.POINTER(C.c_float)
# the build options of round 1 (inner-loop groups, guarded groups) were adopted or removed after their A/B on the GPU
# (profiles/r02/): the shipped source is the only variant left
VARIANTS = {"default": ()}
_libs = {}


def emu(variant):
    if variant not in _libs:
        sys.path.insert(0, os.path.join(ROOT, "tools", "emu"))
        import build_emu
        L = C.CDLL(build_emu.build("" if variant == "default" else variant, VARIANTS[variant]))
        L.emu_lin_solve.argtypes = [C.c_int, C.c_int, FP, FP, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
        L.emu_lin_solve.restype = C.c_int
        L.emu_source_lin_solve.argtypes = [C.c_int, C.c_int, FP, FP, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]
        L.emu_source_lin_solve.restype = C.c_int
        L.emu_set_steal_variant.argtypes = [C.c_int]
        L.emu_slab_lin_solve.argtypes = [C.c_int, C.c_int, C.c_int, FP, FP, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_float]
        L.emu_slab_lin_solve.restype = C.c_int
        _libs[variant] = L
    return _libs[variant]


def p(a):
    return a.ctypes.data_as(FP)


def same(a, b):
    return np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.fixture(scope="module")
def rb():
    from oracle.pyoracle import RedBlackCheck
    return RedBlackCheck()


@pytest.mark.parametrize("variant", ["default"])
def test_kernel_source_on_the_smallest_grids(oracle, variant):
    L = emu(variant)
    rng = np.random.default_rng(0)
    for N in (2, 6, 10, 30):
        G = N + 2
        for T in ((1, 2, 3, 7, 8) if variant == "default" else (3, 7)):
            for b, (al, be), K, zg in ((0, (1.0, 4.0), 5, 0), (1, (0.635, 3.54), 2 * T, 0), (2, (2683.2, 10733.8), T + 2, 0),
                                       (0, (1.0, 4.0), 7, 1)):
                x = rng.uniform(-1, 1, (G, G)).astype(np.float32); x0 = rng.uniform(-1, 1, (G, G)).astype(np.float32)
                if zg:
                    x[...] = 0
                want = x.copy(); oracle.diffuse(N, b, want, x0, al, be, K)
                got = x.copy()
                if zg:
                    got[...] = np.nan          # an implicit zero guess is never read
                assert L.emu_lin_solve(N, b, p(got), p(x0), al, be, K, T, zg, 0, 0, 1.0) == 0
                assert same(got, want), (variant, N, T, b, K, zg)


def test_fused_add_source_in_the_kernel_source(oracle):
    """jacobi_stream_kernel<T, STRICT, 6> (SF_OPT_FUSE_SOURCES): the first launch of a solve forms x0 + dt * s itself.  Against
    add_source + diffuse of the oracle; the raw field must come back untouched; small grids, two bands, chunked, a front that
    decays through the division's low range (guarded groups and pipeline restarts re-form the right-hand side)."""
    L = emu("default")
    rng = np.random.default_rng(5)
    dt = 0.016
    cases = [(N, T, K, 0) for N in (2, 6, 30) for T, K in ((5, 10), (6, 12), (7, 14))] + [(126, 7, 14, 0), (126, 7, 40, 24), (126, 5, 20, 16)]
    for N, T, K, chunk in cases:
        G = N + 2
        for b, (al, be) in ((0, (2683.2, 10733.8)), (1, (0.635, 3.54)), (2, (2683.2, 10733.8))):
            src = rng.uniform(0, 1, (G, G)).astype(np.float32); raw = rng.uniform(-1, 1, (G, G)).astype(np.float32)
            if N == 126 and b == 0:      # density-like: compact support, a front running through 1e-30 .. 1e-45 .. 0
                yy, xx = np.mgrid[0:G, 0:G]
                r = np.hypot(yy - G / 2, xx - G / 2)
                raw = (np.float32(0.05) * np.exp(-np.maximum(r - 10, 0) * 3.0)).astype(np.float32)
                src = (raw * np.float32(0.5)).astype(np.float32)
            rhs = raw.copy(); oracle.add_source(N, rhs, src, dt)
            want = src.copy(); oracle.diffuse(N, b, want, rhs, al, be, K)
            got, raw_in = src.copy(), raw.copy()
            assert L.emu_source_lin_solve(N, b, p(got), p(raw_in), dt, al, be, K, T, chunk) == 0
            assert same(got, want), (N, T, K, chunk, b)
            assert same(raw_in, raw)


def test_peer_slab_kernels_in_the_kernel_source(oracle):
    """jacobi_stream_kernel<T, MODE, 2 / 8> (connected peer slabs) on the host: 2 and 3 slabs of one grid, each with 8 ghost
    rows a side, run their launches in turn; the strip warps store into the neighbour's ghost rows and post / wait on the
    counters as on the device.  Checked against the oracle's lin_solve on the whole grid: the plain strict and pressure solves,
    the solve with add_source fused into its first launch (the strip warps form the ghost rows of the right-hand side
    themselves), with equal chunks and with the short chunks behind the strips (SF_OPT_STRIP_BALANCE)."""
    L = emu("default")
    rng = np.random.default_rng(17)
    dt = 0.016
    al, be = 2683.2, 10733.8
    #        N, world, T,  K, chunk, short chunks expected with balance = 1
    cases = [(126, 2, 6, 12, 0, False), (254, 2, 7, 14, 60, True), (382, 3, 5, 10, 40, True)]
    for N, world, T, K, chunk, expect in cases:
        G = N + 2
        for b in ((0, 1, 2) if N == 126 else (1,)):      # (the wall rules of all three field kinds on the small grid)
            src = rng.uniform(0, 1, (G, G)).astype(np.float32); raw = rng.uniform(-1, 1, (G, G)).astype(np.float32)
            rhs = raw.copy(); oracle.add_source(N, rhs, src, dt)
            want_fused = src.copy(); oracle.diffuse(N, b, want_fused, rhs, al, be, K)
            want_plain = src.copy(); oracle.diffuse(N, b, want_plain, raw, al, be, K)
            for balance in ((1, 0) if b == 1 else (1,)):
                got = src.copy()
                n = L.emu_slab_lin_solve(N, world, b, p(got), p(raw), al, be, K, T, 0, chunk, balance, 0, 0.0)
                assert n >= 0 and (n > 0) == (expect and balance == 1), (N, world, T, K, b, balance, n)
                assert same(got, want_plain), ("plain", N, world, T, K, b, balance)
                got, raw_in = src.copy(), raw.copy()
                n = L.emu_slab_lin_solve(N, world, b, p(got), p(raw_in), al, be, K, T, 0, chunk, balance, 1, dt)
                assert n >= 0 and (n > 0) == (expect and balance == 1), (N, world, T, K, b, balance, n)
                assert same(got, want_fused), ("fused", N, world, T, K, b, balance)
                assert same(raw_in, raw)
        if N == 126:   # K = 40 as in the benchmark: launches of 7,7,7,7,6,6 sweeps, strips of 7 and 6 rows
            src = rng.uniform(0, 1, (G, G)).astype(np.float32); raw = rng.uniform(-1, 1, (G, G)).astype(np.float32)
            rhs = raw.copy(); oracle.add_source(N, rhs, src, dt)
            want = src.copy(); oracle.diffuse(N, 2, want, rhs, al, be, 40)
            got = src.copy()
            assert L.emu_slab_lin_solve(N, world, 2, p(got), p(raw), al, be, 40, 7, 0, chunk, 1, 1, dt) >= 0
            assert same(got, want), ("K = 40, fused", N, world)
        if N == 254:   # the density solve's variants (work stealing, zero-row shortcut) on a compactly supported field: VAR 4 / 9
            yy, xx = np.mgrid[0:G, 0:G]
            raw = np.where(np.hypot(yy - G / 2, xx - G / 2) < 40, rng.uniform(0, 0.1, (G, G)), 0.0).astype(np.float32)
            src = (raw * np.float32(0.5)).astype(np.float32)
            rhs = raw.copy(); oracle.add_source(N, rhs, src, dt)
            want_fused = src.copy(); oracle.diffuse(N, 0, want_fused, rhs, al, be, K)
            want_plain = src.copy(); oracle.diffuse(N, 0, want_plain, raw, al, be, K)
            L.emu_set_steal_variant(1)
            try:
                got = src.copy()
                assert L.emu_slab_lin_solve(N, world, 0, p(got), p(raw), al, be, K, T, 0, chunk, 1, 0, 0.0) > 0
                assert same(got, want_plain), ("stealing variant, plain", N, world)
                got = src.copy()
                assert L.emu_slab_lin_solve(N, world, 0, p(got), p(raw), al, be, K, T, 0, chunk, 1, 1, dt) > 0
                assert same(got, want_fused), ("stealing variant, fused", N, world)
            finally:
                L.emu_set_steal_variant(0)
        # the pressure solve of project(): alpha = 1, beta = 4, implicit zero guess
        div = rng.uniform(-1, 1, (G, G)).astype(np.float32)
        want = np.zeros((G, G), np.float32); oracle.diffuse(N, 0, want, div, 1.0, 4.0, K)
        got = np.full((G, G), np.nan, np.float32)
        assert L.emu_slab_lin_solve(N, world, 0, p(got), p(div), 1.0, 4.0, K, T, 1, chunk, 1, 0, 0.0) >= 0
        assert same(got, want), ("pressure", N, world, T, K)


def test_slot_class_tickets_and_unequal_chunks_partition_the_rows(oracle):
    """SF_OPT_WAVE_SKEW: warps draw their item from the counter of their hardware-slot class and the three thirds of the
    items have chunks of different heights.  Whatever the slots are (the emulated ones put every fifth CTA out of order, so
    classes run dry and fall through), every row must be produced exactly once, and the counters must be back at zero."""
    L = emu("default")
    L.emu_set_wave_skew.argtypes = [C.c_int]
    rng = np.random.default_rng(21)
    try:
        for code in (131103, 150110, 120100):
            L.emu_set_wave_skew(code)
            for N, chunk, T, K in ((126, 14, 6, 12), (254, 28, 7, 14), (254, 42, 7, 7), (510, 85, 7, 14)):
                G = N + 2
                x = rng.uniform(-1, 1, (G, G)).astype(np.float32); x0 = rng.uniform(-1, 1, (G, G)).astype(np.float32)
                want = x.copy(); oracle.diffuse(N, 1, want, x0, 2683.2, 10733.8, K)
                got = x.copy()
                assert L.emu_lin_solve(N, 1, p(got), p(x0), 2683.2, 10733.8, K, T, 0, chunk, 0, 1.0) == 0
                assert same(got, want), (code, N, chunk, T)
                assert L.emu_ticket_words_nonzero() == 0
    finally:
        L.emu_set_wave_skew(0)


def test_zero_row_shortcut_of_the_scalar_field_variants(oracle):
    """jacobi_stream_kernel<T, STRICT, 3 / 7> (scalar fields: dens_step's solve): groups whose last 2T+3 input rows were
    all-zero bits only store zeros.  Compactly supported fields with zero margins of every width around them, a blob that
    starts right after a long zero run (the shortcut must end in time), -0.0 rows (not zero bits), the fused-source form."""
    L = emu("default")
    L.emu_set_steal_variant(1)
    try:
        rng = np.random.default_rng(9)
        N = 254; G = N + 2
        al, be = 107322.0, 429289.0
        for case in range(5):
            x = np.zeros((G, G), np.float32); x0 = np.zeros((G, G), np.float32)
            for _ in range(3):
                r0, c0 = int(rng.integers(1, G - 40)), int(rng.integers(1, G - 40))
                h, w = int(rng.integers(1, 40)), int(rng.integers(1, 40))
                x0[r0:r0 + h, c0:c0 + w] = rng.uniform(0, 0.1, (h, w)).astype(np.float32)
                if case % 2:
                    x[r0:r0 + h, c0:c0 + w] = rng.uniform(0, 0.1, (h, w)).astype(np.float32)
            if case == 3:
                x[100:103, :] = np.float32(-0.0)           # -0.0 is not "zero bits": ((-0) + (-0)) stays -0 through a sweep
                x0[100:103, :] = np.float32(-0.0)
            if case == 4:
                x0[G - 2, 5] = np.float32(1e-3)            # a single cell at the very end of a band's rows
            for T, K, chunk in ((7, 14, 0), (7, 20, 64), (5, 10, 0), (6, 12, 100)):
                want = x.copy(); oracle.diffuse(N, 0, want, x0, al, be, K)
                got = x.copy()
                assert L.emu_lin_solve(N, 0, p(got), p(x0), al, be, K, T, 0, chunk, 0, 1.0) == 0
                assert same(got, want), ("zero rows", case, T, K, chunk)
            dt = 0.016
            rhs = x0.copy(); oracle.add_source(N, rhs, x, dt)
            want = x.copy(); oracle.diffuse(N, 0, want, rhs, al, be, 14)
            got, raw = x.copy(), x0.copy()
            assert L.emu_source_lin_solve(N, 0, p(got), p(raw), dt, al, be, 14, 7, 0) == 0
            assert same(got, want), ("zero rows, fused source", case)
    finally:
        L.emu_set_steal_variant(0)


def test_unproven_right_hand_side_rows_keep_the_range_test(oracle):
    """The strict group drops the per-cell low-range test where every right-hand-side row it uses is proven (|x0| >= 2^-75,
    row_flags in csrc/sf_jacobi.cu).  Adversarial layout for the bookkeeping: an O(1) right-hand side with scattered blocks of
    zero / 1e-41 cells (single cells, single rows, blocks that straddle the band edge and chunk boundaries) over a TINY
    iterate, so that the numerators inside the blocks sit far below the fast division's range for several levels -- a tick
    that used an unproven row without the test would round a subnormal quotient wrongly.  A failing seed prints its layout."""
    L = emu("default")
    N = 126; G = N + 2
    al, be = 2683.2, 10733.8
    for seed in range(6):
        rng = np.random.default_rng(100 + seed)
        x0 = (rng.uniform(0.1, 1, (G, G)) * rng.choice([-1.0, 1.0], (G, G))).astype(np.float32)
        x = (10.0 ** rng.uniform(-43, -37, (G, G)) * rng.choice([-1.0, 1.0], (G, G))).astype(np.float32)
        blocks = []
        for k in range(10):
            # (the 3-instruction division is wrong only for a small fraction of the numerators below its range: tall, wide
            # blocks keep thousands of them tiny down to the deepest level.  What random data cannot reach is the tail of the
            # window -- the ticks r+1 .. r+T after a block's last row r: there the block's cells have proven neighbours
            # below them, whose values are far from tiny unless they cancel to within 2^-24; the window is kept anyway,
            # the proof in row_flags needs it)
            h, w = (int(rng.choice([12, 20, 30])), int(rng.choice([40, 90, 120]))) if k < 3 else (int(rng.choice([1, 1, 2, 5, 9])), int(rng.choice([1, 3, 9, 30])))
            r0, c0 = int(rng.integers(1, G - h - 1)), int(rng.choice([int(rng.integers(1, G - w - 1)), 108, 112, 2]))
            c0 = min(c0, G - w - 1)
            val = float(rng.choice([0.0, 1e-41, -3e-40, 1e-30]))
            x0[r0:r0 + h, c0:c0 + w] = np.float32(val)
            blocks.append((r0, h, c0, w, val))
        for T, K, chunk, b in ((7, 14, 0, 0), (7, 13, 16, 1), (6, 12, 24, 2), (5, 10, 20, 0)):
            want = x.copy(); oracle.diffuse(N, b, want, x0, al, be, K)
            got = x.copy()
            assert L.emu_lin_solve(N, b, p(got), p(x0), al, be, K, T, 0, chunk, 0, 1.0) == 0
            assert same(got, want), (seed, T, K, chunk, b, blocks)


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_kernel_source_two_bands_chunks_and_a_decaying_front(oracle, variant):
    """G = 128 (two bands), small chunks, and a compactly supported field whose front decays through the low end of the exact
    division's range: outlier votes, guarded binary64 ticks and pipeline restarts of the strict kernel."""
    L = emu(variant)
    rng = np.random.default_rng(1)
    N = 126; G = N + 2
    x = rng.uniform(-1, 1, (G, G)).astype(np.float32); x0 = rng.uniform(-1, 1, (G, G)).astype(np.float32)
    for T, chunk, (al, be), K in ((7, 0, (0.635, 3.54), 14), (3, 24, (1.0, 4.0), 6)):
        want = x.copy(); oracle.diffuse(N, 1, want, x0, al, be, K)
        got = x.copy()
        assert L.emu_lin_solve(N, 1, p(got), p(x0), al, be, K, T, 0, chunk, 0, 1.0) == 0
        assert same(got, want), (variant, T, chunk)
    src = np.zeros((G, G), np.float32)
    src[50:70, 40:90] = rng.uniform(0, 1, (G, G)).astype(np.float32)[50:70, 40:90] * np.float32(1e-24)
    al, be = 107322.0, 429289.0
    want = np.zeros((G, G), np.float32); oracle.diffuse(N, 0, want, src, al, be, 14)
    assert np.any((want != 0) & (np.abs(want) < 1e-30)), "the case must reach below the fast division's range"
    got = np.zeros((G, G), np.float32)
    assert L.emu_lin_solve(N, 0, p(got), p(src), al, be, 14, 7, 0, 0, 0, 1.0) == 0
    assert same(got, want), (variant, "front")
    # rows beyond the HIGH end of the validated range: detected when the row is fetched (row_is_big), the group runs guarded
    x = rng.uniform(-1, 1, (G, G)).astype(np.float32); x0 = rng.uniform(-1, 1, (G, G)).astype(np.float32)
    x[40:43, 10:20] = np.float32(3e31); x0[90, 60:64] = np.float32(-2e33)
    al, be = 0.635, 3.54
    want = x.copy(); oracle.diffuse(N, 2, want, x0, al, be, 14)
    got = x.copy()
    assert L.emu_lin_solve(N, 2, p(got), p(x0), al, be, 14, 7, 0, 0, 0, 1.0) == 0
    assert same(got, want), (variant, "huge values")


@pytest.mark.parametrize("variant", ["default"])
def test_red_black_levels_in_the_kernel_source(rb, variant):
    """SF_OPT_RBGS_BLOCKED: jacobi_stream_kernel<T, MODE, 5> against the in-place red-black scheme."""
    L = emu(variant)
    rng = np.random.default_rng(2)
    for N in (2, 6, 10, 30, 126):
        G = N + 2
        for om in (1.0, 1.5):
            for b, (al, be), K in ((0, (1.0, 4.0), 7), (1, (0.635, 3.54), 5), (2, (2683.2, 10733.8), 4), (1, (0.635, 3.54), 1)):
                if N == 126 and K > 5:
                    continue
                x = rng.uniform(-1, 1, (G, G)).astype(np.float32); x0 = rng.uniform(-1, 1, (G, G)).astype(np.float32)
                want = x.copy(); rb.rb_diffuse(N, b, want, x0, al, be, K, om)
                got = x.copy()
                assert L.emu_lin_solve(N, b, p(got), p(x0), al, be, K, 6, 0, 0, 1, om) == 0
                assert same(got, want), (variant, N, om, b, K)


def test_stage_kernels_source_on_the_smallest_grids(oracle):
    """csrc/sf_stages.cu as written (set_bnd, add_source, advect, divergence, gradient subtract; the float4 row kernels with
    their shuffles at G % 4 == 0, the per-cell kernels otherwise) against the oracle at G = 3 .. 32.  This is the check that
    reproduces the N = 1 corner bug of the old set_bnd kernel."""
    sys.path.insert(0, os.path.join(ROOT, "tools", "emu"))
    import build_emu
    L = C.CDLL(build_emu.build_stages())
    i, f = C.c_int, C.c_float
    L.emu_set_bnd.argtypes = [i, i, FP]; L.emu_add_source.argtypes = [i, FP, FP, f]; L.emu_advect.argtypes = [i, i, FP, FP, FP, FP, f]
    L.emu_advect_uv.argtypes = [i, FP, FP, FP, FP, f]; L.emu_divergence.argtypes = [i, FP, FP, FP, FP, i]
    L.emu_last_project.argtypes = [i, FP, FP, FP]
    for fn in ("emu_set_bnd", "emu_add_source", "emu_advect", "emu_advect_uv", "emu_divergence", "emu_last_project"):
        getattr(L, fn).restype = None
    rng = np.random.default_rng(0)
    DT = 0.016
    rnd = lambda G, lo=-1, hi=1: rng.uniform(lo, hi, (G, G)).astype(np.float32)
    zeros = lambda G: np.zeros((G, G), np.float32)
    for N in (1, 2, 3, 4, 5, 6, 10, 13, 14, 30):
        G = N + 2
        for b in (0, 1, 2):
            x = rnd(G); w = x.copy(); oracle.set_bnd(N, b, w); L.emu_set_bnd(N, b, p(x))
            assert same(x, w), ("set_bnd", N, b)
        x, s_ = rnd(G), rnd(G); w = x.copy(); oracle.add_source(N, w, s_, DT); L.emu_add_source(N, p(x), p(s_), DT)
        assert same(x, w), ("add_source", N)
        u, v = rnd(G, -.3, .3), rnd(G, -.3, .3)
        for b in (0, 1, 2):
            d0 = rnd(G); w = zeros(G); oracle.advect(N, b, w, d0, u, v, DT)
            d = zeros(G); L.emu_advect(N, b, p(d), p(d0), p(u), p(v), DT)
            assert same(d, w), ("advect", N, b)
        wu, wv = zeros(G), zeros(G); oracle.advect(N, 1, wu, u, u, v, DT); oracle.advect(N, 2, wv, v, u, v, DT)
        du, dv = zeros(G), zeros(G); L.emu_advect_uv(N, p(du), p(dv), p(u), p(v), DT)
        assert same(du, wu) and same(dv, wv), ("advect_uv", N)
        wp, wd = zeros(G), zeros(G); oracle.computeDivergenceAndPressure(N, u, v, wp, wd)
        gp, gd = rnd(G), rnd(G); L.emu_divergence(N, p(u), p(v), p(gp), p(gd), 1)
        assert same(gp, wp) and same(gd, wd), ("divergence", N)
        pr = rnd(G); wu, wv = u.copy(), v.copy(); oracle.lastProject(N, wu, wv, pr, wd)
        gu, gv = u.copy(), v.copy(); L.emu_last_project(N, p(gu), p(gv), p(pr))
        assert same(gu, wu) and same(gv, wv), ("last_project", N)


def test_red_black_half_sweep_kernel_source(rb):
    """csrc/sf_solvers.cu as written (the default, launch-per-half-sweep form of SF_SOLVER_RBGS) against the in-place scheme."""
    sys.path.insert(0, os.path.join(ROOT, "tools", "emu"))
    import build_emu
    L = C.CDLL(build_emu.build_stages())
    L.emu_rbgs.argtypes = [C.c_int, C.c_int, FP, FP, C.c_float, C.c_float, C.c_int, C.c_float]
    L.emu_rbgs.restype = None
    rng = np.random.default_rng(3)
    for N in (1, 2, 5, 14):
        G = N + 2
        for om in (1.0, 1.5):
            for b, (al, be), K in ((0, (1.0, 4.0), 4), (1, (0.635, 3.54), 3), (2, (2683.2, 10733.8), 2)):
                x = rng.uniform(-1, 1, (G, G)).astype(np.float32); x0 = rng.uniform(-1, 1, (G, G)).astype(np.float32)
                om32 = float(np.float32(om))
                want = x.copy(); rb.rb_diffuse(N, b, want, x0, al, be, K, om32)
                got = x.copy(); L.emu_rbgs(N, b, p(got), p(x0), al, be, K, om32)
                assert same(got, want), (N, om, b, K)
