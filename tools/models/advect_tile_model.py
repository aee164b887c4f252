"""numpy transcription of advect_tile_kernel (csrc/sf_stages.cu) checked against the CPU oracle -- no GPU needed.

What it models, tile by tile (8 * RPW rows x 128 columns per CTA): the back-traces in binary32 with the reference's operand
order, the truncation without a conversion (adding 2^23 with rounding toward zero, `trunc_pos`), the bounding box of the
traces with its first column rounded down to a multiple of 4 cells (the TMA unit's 16-byte rule), the fit test (160 columns,
8 * maxsub rows, inside the stored rows), the 2-D tensor copies as the TMA unit delivers them (8-row boxes, zeros outside the
array, only as many as the box is high), the index arithmetic of the shared-memory gathers, the whole-tile fallback to global
gathers (box too large, NaN velocity), and set_bnd fused into the stores.  Every gather of a fitted tile is served from the
modelled box ONLY (the rest of the box buffer holds NaN), so an index that leaves the box, or a box that misses a source cell,
shows up as a mismatch with the oracle's advect (FluidSequential.c:107-141)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
f32 = np.float32
BOXW, SUB = 160, 8


def trunc_pos(x):
    """(float)(int)x and (int)x for 0.5 <= x < 2^23 via RZ(x + 2^23): exact sum in binary64, rounded toward zero to the
    spacing-1 grid of [2^23, 2^24)"""
    t = np.floor(x.astype(np.float64) + 8388608.0)          # RZ to binary32: the exact sum is positive and below 2^24
    i = (t - 8388608.0).astype(np.int64)
    return (t - 8388608.0).astype(f32), i


def bilinear(a00, a10, a01, a11, wx0, wx1, wy0, wy1):
    # wx0 * (wy0 * a00 + wy1 * a10) + wx1 * (wy0 * a01 + wy1 * a11), every operation rounded to binary32
    return (wx0 * (wy0 * a00 + wy1 * a10).astype(f32)).astype(f32) + (wx1 * (wy0 * a01 + wy1 * a11).astype(f32)).astype(f32)


def advect_tiles(N, b, d0, u, v, dt, rpw=2, maxsub=5, stats=None):
    G = N + 2
    assert G % 4 == 0
    dt0 = f32(dt) * f32(N)
    hiC = f32(N) + f32(0.5)
    d = np.full((G, G), np.nan, f32)
    rows_per_tile = 8 * rpw
    for tr in range(1, N + 1, rows_per_tile):
        r_hi = min(tr + rows_per_tile, N + 1)
        for seg in range(0, G, 128):
            cols = np.minimum(np.arange(seg, seg + 128), G - 1)           # idle lanes mirror the last column
            rows = np.arange(tr, r_hi)
            R, C = np.meshgrid(rows, cols, indexing="ij")
            with np.errstate(invalid="ignore", over="ignore"):
                px = (C.astype(f32) - (dt0 * u[R, C]).astype(f32)).astype(f32)
                py = (R.astype(f32) - (dt0 * v[R, C]).astype(f32)).astype(f32)
            px = np.where(px < f32(0.5), f32(0.5), px); px = np.where(px > hiC, hiC, px)
            py = np.where(py < f32(0.5), f32(0.5), py); py = np.where(py > hiC, hiC, py)
            nan = ~(px <= hiC) | ~(py <= hiC)
            fit = not nan.any()
            if fit:
                cmin, cmax = int(np.min(px)), int(np.max(px))
                rmin, rmax = int(np.min(py)), int(np.max(py))
                cmin &= ~3
                span_r = rmax + 2 - rmin
                fit = (cmax + 2 - cmin <= BOXW) and (span_r <= SUB * maxsub) and rmin >= 0 and rmax + 1 < G
            if stats is not None:
                stats[0 if fit else 1] += 1
            with np.errstate(invalid="ignore"):
                pxs = np.where(nan, f32(0.0), px); pys = np.where(nan, f32(0.0), py)   # (int)NaN = 0 on the device
            fc, c0 = trunc_pos(np.where(nan, f32(0.5), pxs)); fr, r0 = trunc_pos(np.where(nan, f32(0.5), pys))
            c0 = np.where(nan, 0, c0); r0 = np.where(nan, 0, r0)
            fc = np.where(nan, f32(0.0), fc); fr = np.where(nan, f32(0.0), fr)
            assert np.array_equal(c0[~nan], px[~nan].astype(np.int64)) and np.array_equal(r0[~nan], py[~nan].astype(np.int64))
            with np.errstate(invalid="ignore"):
                wx1 = (px - fc).astype(f32); wx0 = (f32(1.0) - wx1).astype(f32)
                wy1 = (py - fr).astype(f32); wy0 = (f32(1.0) - wy1).astype(f32)
            if fit:
                nsub = (span_r + SUB - 1) // SUB
                box = np.full((maxsub * SUB, BOXW), np.nan, f32)
                for s in range(nsub):                                      # one tensor copy each: zeros outside the array
                    for rr in range(SUB):
                        gr = rmin + s * SUB + rr
                        for_cols = np.arange(cmin, cmin + BOXW)
                        ok = (for_cols < G) & (0 <= gr < G)
                        box[s * SUB + rr] = np.where(ok, d0[min(max(gr, 0), G - 1), np.minimum(for_cols, G - 1)], f32(0.0))
                j = (r0 - rmin) * BOXW + (c0 - cmin)
                flat = box.reshape(-1)
                assert j.min() >= 0 and (j + BOXW + 1).max() < nsub * SUB * BOXW, "gather leaves the copied part of the box"
                a00, a10, a01, a11 = flat[j], flat[j + BOXW], flat[j + 1], flat[j + BOXW + 1]
            else:
                rs = np.minimum(np.maximum(r0, 0), G - 2)
                a00, a10, a01, a11 = d0[rs, c0], d0[rs + 1, c0], d0[rs, c0 + 1], d0[rs + 1, c0 + 1]
            with np.errstate(invalid="ignore"):
                o = bilinear(a00, a10, a01, a11, wx0, wx1, wy0, wy1).astype(f32)
            valid = np.arange(seg, seg + 128) < G
            d[tr:r_hi, seg:min(seg + 128, G)] = o[:, valid]
    # set_bnd(b) as the stores fuse it (FluidSequential.c:62-75)
    sx = f32(-1.0) if b == 1 else f32(1.0)
    sy = f32(-1.0) if b == 2 else f32(1.0)
    d[1:N + 1, 0] = sx * d[1:N + 1, 1]; d[1:N + 1, N + 1] = sx * d[1:N + 1, N]
    d[0, 1:N + 1] = sy * d[1, 1:N + 1]; d[N + 1, 1:N + 1] = sy * d[N, 1:N + 1]
    for (r, c, rn, cn) in ((0, 0, 1, 1), (0, N + 1, 1, N), (N + 1, 0, N, 1), (N + 1, N + 1, N, N)):
        d[r, c] = f32(0.5) * (d[r, cn] + d[rn, c]).astype(f32)
    return d


def smooth(rng, G, radius):
    a = rng.uniform(-1.0, 1.0, (G + 2 * radius, G + 2 * radius))
    c = np.pad(np.cumsum(np.cumsum(a, 0), 1), ((1, 0), (1, 0)))
    k = 2 * radius + 1
    return ((c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k])[:G, :G] / (k * k))


def main(sizes=(318, 382), seed=3):
    from oracle.pyoracle import Oracle
    orc = Oracle()
    rng = np.random.default_rng(seed)
    DT = 0.016
    for N in sizes:
        G = N + 2
        cells = 1.0 / (DT * N)
        yy, xx = np.mgrid[0:G, 0:G]
        fields = {
            "drift": (65.0 + 25.0 * smooth(rng, G, 4), -40.0 + 25.0 * smooth(rng, G, 4)),
            "walls": ((xx - G / 2) * 0.2 + 10.0 * smooth(rng, G, 3), (yy - G / 2) * -0.2 + 10.0 * smooth(rng, G, 3)),
            "mixed": (20.0 + np.where(xx > G // 2, rng.uniform(-30, 30, (G, G)), 15.0 * smooth(rng, G, 3)), 20.0 + 15.0 * smooth(rng, G, 3)),
        }
        for name, (uu, vv) in fields.items():
            u, v = (uu * cells).astype(f32), (vv * cells).astype(f32)
            d0 = rng.uniform(-1.0, 1.0, (G, G)).astype(f32)
            for b, rpw, maxsub in ((0, 2, 5), (1, 4, 8), (2, 2, 3)):
                want = np.zeros((G, G), f32)
                orc.advect(N, b, want, d0, u, v, DT)
                stats = [0, 0]
                got = advect_tiles(N, b, d0, u, v, DT, rpw, maxsub, stats)
                bad = got.view(np.uint32) != want.view(np.uint32)
                assert not bad.any(), f"N={N} {name} b={b} rpw={rpw} maxsub={maxsub}: {int(bad.sum())} cells differ, first {tuple(np.argwhere(bad)[0])}"
                assert stats[0] > 0, (name, stats)
                if name == "mixed":
                    assert stats[1] > 0, stats
                print(f"N={N} {name:6s} b={b} rpw={rpw} maxsub={maxsub}: identical; tiles by box {stats[0]}, by fallback {stats[1]}")


if __name__ == "__main__":
    main()
