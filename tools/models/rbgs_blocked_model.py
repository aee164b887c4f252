"""Design check (no GPU) for the NEXT step of the opt-in red-black solver: temporal blocking of red-black
Gauss-Seidel / SOR on the streaming pipeline of jacobi_stream_kernel.

The shipped SF_SOLVER_RBGS (csrc/sf_solvers.cu) is one launch per half-sweep.  This model shows that the register
pipeline of the Jacobi kernel carries the red-black scheme unchanged in structure: one red-black ITERATION is two
pipeline LEVELS (level 2k+1 = after the red half-sweep, level 2k+2 = after the black one), a level updates the cells
of its colour and copies the others through, and set_bnd happens on black levels only (wall columns / wall rows copy
through on red levels).  A launch of T levels fuses T/2 iterations and is bit-identical to the in-place scheme of
oracle/rbgs_check.c -- which is what this file asserts, with everything never loaded poisoned by NaN.

It reuses tools/models/stream_model.py (the transcription of the Jacobi kernel) and overrides the tick."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import stream_model as sm   # noqa: E402

f32 = np.float32


class RBWarp(sm.Warp):
    def tick(self, s, row_in, W, rring, PH, WALLS):
        A, T = self.A, self.T
        UP, MID, DN = PH % 3, (PH + 1) % 3, (PH + 2) % 3
        W[0][DN] = row_in
        out = None
        cols = self.cc[:, None] + np.arange(4)[None, :]
        for t in range(T):
            a = s - t - 1
            black = (t & 1) == 1                       # level t+1 is the state after a black half-sweep
            colour = 1 if black else 0
            up, mid, dn = W[t][UP], W[t][MID], W[t][DN]
            r = rring(a)
            lft = np.concatenate([mid[:1, 3], mid[:-1, 3]])
            rgt = np.concatenate([mid[1:, 0], mid[-1:, 0]])
            gs = sm.jacobi4(A["mode"], lft, mid, rgt, up, dn, r, A["alpha"], A["beta"])
            if A["omega"] != 1.0:
                with np.errstate(all="ignore"):
                    gs = (mid + (f32(A["omega"]) * (gs - mid).astype(f32)).astype(f32)).astype(f32)
            upd = ((a + cols) & 1) == colour           # cells of this level's colour
            o = np.where(upd, gs, mid).astype(f32)
            if black:                                  # set_bnd(b) on the wall columns: after the black half-sweep only
                o[self.ownsL, 0] = (f32(A["sx"]) * o[self.ownsL, 1]).astype(f32)
                o[self.ownsR, 3] = (f32(A["sx"]) * o[self.ownsR, 2]).astype(f32)
            else:                                      # red level: wall columns copy through
                o[self.ownsL, 0] = mid[self.ownsL, 0]
                o[self.ownsR, 3] = mid[self.ownsR, 3]
            if WALLS and t + 1 < T:
                if a == A["N"] + 1:                    # row N+1 at level t+1
                    o = (W[t + 1][MID] * f32(A["sy"])).astype(f32) if black else mid.copy()
                if a == 1:                             # row 0 at level t+1 lives in the next level's MID slot
                    W[t + 1][MID] = (o * f32(A["sy"])).astype(f32) if black else up.copy()
            if t + 1 < T:
                W[t + 1][DN] = o
            else:
                out = o
        return out


def launch(xout, xin, rhs, N, b, mode, alpha, beta, omega, T, chunk_rows=0):
    assert T % 2 == 0, "a launch fuses whole iterations"
    G = N + 2
    A = dict(xin=xin, rhs=rhs, xout=xout, G=G, N=N, mode=mode, alpha=alpha, beta=beta, omega=omega, zero_guess=False,
             sx=-1.0 if b == 1 else 1.0, sy=-1.0 if b == 2 else 1.0, write_top=True, write_bot=True)
    nbands = (G + sm.VALID_W - 1) // sm.VALID_W
    rows = N
    chunk = chunk_rows if chunk_rows > 0 else max(rows, 1)
    chunk = min(chunk, rows)
    for ch in range((rows + chunk - 1) // chunk):
        lo = 1 + ch * chunk
        hi = min(lo + chunk, N + 1)
        for band in range(nbands):
            RBWarp(A, band, T).stream_rows(lo, hi)


def rb_lin_solve(N, b, x, x0, alpha, beta, iters, omega, T=8, chunk_rows=0):
    mode = "pressure" if (alpha == 1.0 and beta == 4.0) else "strict"
    per = T // 2
    plan = []
    left = iters
    while left > 0:
        plan.append(min(per, left)); left -= plan[-1]
    scratch = np.full_like(x, np.nan)
    cur, nxt = x, scratch
    for k in plan:
        launch(nxt, cur, x0, N, b, mode, alpha, beta, omega, 2 * k, chunk_rows)
        cur, nxt = nxt, cur
    if cur is not x:
        x[...] = cur


def main(sizes=(2, 6, 10, 14, 30, 62, 114, 222)):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
    from oracle.pyoracle import RedBlackCheck
    rb = RedBlackCheck()
    rng = np.random.default_rng(0)
    cases = 0
    for N in sizes:
        G = N + 2
        for T in ((2, 4, 6, 8) if N <= 30 else (8,)):
            for omega in (1.0, 1.5):
                for b, (alpha, beta), iters, chunk in ((0, (1.0, 4.0), 7, 0), (1, (0.635, 3.54), 5, 0), (2, (2683.2, 10733.8), 4, 0),
                                                       (1, (0.635, 3.54), 6, 16)):
                    x = rng.uniform(-1, 1, (G, G)).astype(f32); x0 = rng.uniform(-1, 1, (G, G)).astype(f32)
                    want = x.copy(); rb.rb_diffuse(N, b, want, x0, alpha, beta, iters, omega)
                    got = x.copy(); rb_lin_solve(N, b, got, x0, alpha, beta, iters, omega, T, chunk)
                    # corners: the in-place scheme rewrites them every iteration, the pipeline with the last level -- same values
                    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), \
                        f"blocked red-black model differs: N={N} T={T} omega={omega} b={b} alpha={alpha} iters={iters} chunk={chunk}"
                    cases += 1
    print(f"rbgs_blocked_model: {cases} cases bit-identical to the in-place red-black scheme (G = 4 .. 224, 1..4 iterations per launch)")


if __name__ == "__main__":
    main()
