"""Warp-level model of jacobi_stream_kernel (csrc/sf_jacobi.cu) in numpy binary32 -- no GPU.

A statement-by-statement transcription of stream_rows / pipeline_tick / launch_jacobi_stream / lin_solve for the
single-GPU variant: 32 lanes x float4 per band, T-level register pipeline with three rotating window slots per
level, per-warp rings, fused set_bnd (wall columns in registers, wall rows patched into the next level's window,
corners with the last level), fast groups of three wall-free ticks, chunking and the launch plan.  Everything the
kernel never loads is NaN here (ring slots of rows beyond the grid, lanes outside the grid read zero like a
zero-byte cp.async), so a value that must not be used poisons the result if it is.

Use: the GPU suite proves the kernel bit-identical to the oracle at the sizes it runs; this model, checked against
the same oracle, extends the argument to geometries that have not been on a GPU yet (the smallest grids, odd
chunkings) and documents the algorithm in executable form.  `python tools/models/stream_model.py` runs the check."""
import numpy as np

f32 = np.float32
BAND_W, HALO_X, VALID_W = 128, 8, 112


def jacobi4(mode, lft, mid, rgt, up, dn, r, alpha, beta):
    """Four adjacent cells per lane; mid/up/dn/r: (32, 4).  ((l + r) + up) + dn ; x0 + alpha*sum ; / beta."""
    left = np.concatenate([lft[:, None], mid[:, :3]], axis=1)
    right = np.concatenate([mid[:, 1:], rgt[:, None]], axis=1)
    with np.errstate(all="ignore"):
        s = ((left + right).astype(f32) + up).astype(f32)
        s = (s + dn).astype(f32)
        if mode == "pressure":
            return ((r + s).astype(f32) * f32(0.25)).astype(f32)
        num = (r + (f32(alpha) * s).astype(f32)).astype(f32)
        return (num / f32(beta)).astype(f32)      # the kernel's exact constant division == IEEE division


class Warp:
    def __init__(self, A, band, T):
        self.A, self.T = A, T
        lane = np.arange(32)
        c = band * VALID_W - HALO_X + 4 * lane
        self.indom = (c >= 0) & (c + 4 <= A["G"])
        self.ownsL, self.ownsR = (c == 0), (c + 4 == A["G"])
        self.st_ok = self.indom & (lane >= HALO_X // 4) & (lane < 32 - HALO_X // 4)
        self.cc = np.where(self.indom, c, 0)

    def load(self, field, row):
        """cp.async of one row piece per lane: 16 bytes, or zero-fill for lanes outside the grid."""
        out = np.zeros((32, 4), f32)
        for l in range(32):
            if self.indom[l]:
                out[l] = field[row, self.cc[l]:self.cc[l] + 4]
        return out

    def tick(self, s, row_in, W, rring, PH, WALLS):
        A, T = self.A, self.T
        UP, MID, DN = PH % 3, (PH + 1) % 3, (PH + 2) % 3
        W[0][DN] = row_in
        out = None
        for t in range(T):
            a = s - t - 1
            up, mid, dn = W[t][UP], W[t][MID], W[t][DN]
            r = rring(a)
            lft = np.concatenate([mid[:1, 3], mid[:-1, 3]])      # __shfl_up(mid.w, 1): lane 0 keeps its own
            rgt = np.concatenate([mid[1:, 0], mid[-1:, 0]])      # __shfl_down(mid.x, 1): lane 31 keeps its own
            o = jacobi4(A["mode"], lft, mid, rgt, up, dn, r, A["alpha"], A["beta"])
            o[self.ownsL, 0] = (f32(A["sx"]) * o[self.ownsL, 1]).astype(f32)
            o[self.ownsR, 3] = (f32(A["sx"]) * o[self.ownsR, 2]).astype(f32)
            if WALLS and t + 1 < T:
                if a == A["N"] + 1:
                    o = (W[t + 1][MID] * f32(A["sy"])).astype(f32)
                if a == 1:
                    W[t + 1][MID] = (o * f32(A["sy"])).astype(f32)
            if t + 1 < T:
                W[t + 1][DN] = o
            else:
                out = o
        return out

    def stream_rows(self, a_lo, a_hi):
        A, T = self.A, self.T
        G, N = A["G"], A["N"]
        first = a_lo
        s_lo = max(first - T, 0)
        s_hi = a_hi - 1 + T
        load_hi = min(s_hi, G - 1)
        nan_row = np.full((32, 4), np.nan, f32)
        xr, rr = {}, {}

        def issue(row):
            if row <= load_hi:
                if not A["zero_guess"]:
                    xr[row] = self.load(A["xin"], row)
                rr[row] = self.load(A["rhs"], row)

        def rring(a):
            return rr.get(a, nan_row)            # a ring slot the kernel never filled

        def xrow(row):
            if A["zero_guess"] or row > load_hi:
                return np.zeros((32, 4), f32)
            return xr[row]

        def store(row, o, mask):
            for l in range(32):
                if mask[l]:
                    A["xout"][row, self.cc[l]:self.cc[l] + 4] = o[l]

        def emit_plain(a, o):
            if first <= a < a_hi:
                store(a, o, self.st_ok)

        def emit_walls(a, o):
            if a < first or a >= a_hi:
                return
            store(a, o, self.st_ok)
            for cond, wrow in (((a == 1) and A["write_top"], 0), ((a == N) and A["write_bot"], N + 1)):
                if cond:
                    w = (o * f32(A["sy"])).astype(f32)
                    w[self.ownsL, 0] = (f32(0.5) * (w[self.ownsL, 1] + o[self.ownsL, 0]).astype(f32)).astype(f32)
                    w[self.ownsR, 3] = (f32(0.5) * (w[self.ownsR, 2] + o[self.ownsR, 3]).astype(f32)).astype(f32)
                    store(wrow, w, self.st_ok)

        fast_lo, fast_hi = T + 2, min(s_hi, N)
        W = [[np.zeros((32, 4), f32) for _ in range(3)] for _ in range(T)]
        for k in range(5):
            issue(s_lo + k)
        s = s_lo
        while s <= s_hi:
            if s >= fast_lo and s + 2 <= fast_hi:
                for k in range(3):
                    issue(s + 5 + k)
                for ph in range(3):
                    o = self.tick(s + ph, xrow(s + ph), W, rring, ph, False)
                    emit_plain(s + ph - T, o)
                s += 3
                continue
            issue(s + 5)
            o = self.tick(s, xrow(s), W, rring, 0, True)
            emit_walls(s - T, o)
            for t in range(T):
                W[t][0] = W[t][1]; W[t][1] = W[t][2]
            s += 1


def launch(xout, xin, rhs, N, b, mode, alpha, beta, T, zero_guess, chunk_rows=0, sm_count=148):
    """launch_jacobi_stream for a full-grid context (no strips, no stealing)."""
    G = N + 2
    A = dict(xin=xin, rhs=rhs, xout=xout, G=G, N=N, mode=mode, alpha=alpha, beta=beta, zero_guess=zero_guess,
             sx=-1.0 if b == 1 else 1.0, sy=-1.0 if b == 2 else 1.0, write_top=True, write_bot=True)
    a_lo, a_hi = 1, N + 1
    nbands = (G + VALID_W - 1) // VALID_W
    rows = a_hi - a_lo
    chunk = chunk_rows
    if chunk <= 0:
        heavy = mode == "strict" and T >= 6
        want = max((sm_count * (3 if heavy else 4) * 4) // nbands, 1)
        chunk = (max(rows, 1) + want - 1) // want
        chunk = max(chunk, max(2 * T, 8))
    if chunk > rows:
        chunk = max(rows, 1)
    nchunks = (rows + chunk - 1) // chunk
    for ch in range(nchunks):
        lo = a_lo + ch * chunk
        hi = min(lo + chunk, a_hi)
        for band in range(nbands):
            Warp(A, band, T).stream_rows(lo, hi)


def plan_launches(iters, T, odd_ok):
    L = (iters + T - 1) // T
    if not odd_ok and (L & 1) and L + 1 <= iters:
        L += 1
    plan = [iters // L] * L
    for k in range(iters % L):
        plan[k] += 1
    return plan


def lin_solve(N, b, x, x0, alpha, beta, iters, T=7, zero_guess=False, chunk_rows=0):
    """sf_api.cu lin_solve: ping-pong between x and a scratch field; the result ends in x."""
    mode = "pressure" if (alpha == 1.0 and beta == 4.0) else "strict"
    plan = plan_launches(iters, T, zero_guess)
    scratch = np.full_like(x, np.nan)            # never read before it is written
    cur, nxt = x, scratch
    if zero_guess and (len(plan) & 1):
        cur, nxt = scratch, x
    for k, sweeps in enumerate(plan):
        launch(nxt, cur, x0, N, b, mode, alpha, beta, sweeps, zero_guess and k == 0, chunk_rows)
        cur, nxt = nxt, cur
    if cur is not x:
        x[...] = cur


def main(sizes=(2, 6, 10, 14, 30, 62, 110, 114, 222, 226)):
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
    from oracle.pyoracle import Oracle
    orc = Oracle()
    rng = np.random.default_rng(0)
    cases = 0
    for N in sizes:
        G = N + 2
        for T in ((1, 2, 3, 5, 7, 8) if N <= 30 else (3, 7)):
            for b, (alpha, beta), iters, zg, chunk in ((0, (1.0, 4.0), 2 * T + 1, False, 0), (1, (0.635, 3.54), 3 * T, False, 0),
                                                       (2, (2683.2, 10733.8), T + 2, False, 0), (0, (1.0, 4.0), 20, True, 0),
                                                       (1, (0.635, 3.54), 2 * T, False, 16)):
                if N > 62 and chunk == 0 and iters > 12:
                    iters = 2 * T                 # keep the pure-Python model quick
                x = rng.uniform(-1, 1, (G, G)).astype(f32); x0 = rng.uniform(-1, 1, (G, G)).astype(f32)
                if zg:
                    x[...] = 0.0
                want = x.copy(); orc.diffuse(N, b, want, x0, alpha, beta, iters)
                got = x.copy()
                if zg:
                    got[...] = np.nan            # an implicit zero guess must never be read
                lin_solve(N, b, got, x0, alpha, beta, iters, T, zg, chunk)
                same = np.array_equal(got.view(np.uint32), want.view(np.uint32))
                assert same, f"model differs from the oracle: N={N} T={T} b={b} alpha={alpha} iters={iters} zero_guess={zg} chunk={chunk}"
                cases += 1
    print(f"stream_model: {cases} lin_solve cases bit-identical to the oracle (G = 4 .. 228, depths 1..8, chunked, zero guess)")


if __name__ == "__main__":
    main()
