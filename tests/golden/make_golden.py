"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Runs only where /root/reference exists (the build container): it loads
oracle/_ref/libref_seq_N<N>_K<K>.so -- /root/reference/project/sequential/FluidSequential.c compiled
from where it lies by oracle/Makefile, with only its `#define N` (:6) and `k < 40` (:91) literals
set -- and records

  * stage_N30.npz : every stage function (set_bnd, add_source, diffuse, advect,
                    computeDivergenceAndPressure, lastProject, dens_step, vel_step) applied to
                    seeded random inputs at N=30 (K=4), inputs and outputs;
  * run_N<N>_K<K>.npz : the reference program's own run (FluidSequential.c:244-312: glibc rand()
                    initial condition, sources zeroed after step 0) -- dens,u,v at chosen steps.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import ReferenceSeq  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def stage_fixture(N=30, K=4, seed=1234):
    R = ReferenceSeq(N, K)
    G = N + 2
    rng = np.random.default_rng(seed)
    rnd = lambda lo=-1.0, hi=1.0: rng.uniform(lo, hi, (G, G)).astype(np.float32)
    d = {}
    for b in (0, 1, 2):
        x = rnd(); d[f"set_bnd{b}_in"] = x.copy(); R.set_bnd(b, x); d[f"set_bnd{b}_out"] = x
    x, s = rnd(), rnd()
    d["add_source_x"], d["add_source_s"] = x.copy(), s.copy(); R.add_source(x, s); d["add_source_out"] = x
    for b, (alpha, beta) in zip((0, 1, 2), ((1.0, 4.0), (0.036, 1.144), (90.0, 361.0))):
        x, x0 = rnd(), rnd()
        d[f"diffuse{b}_x"], d[f"diffuse{b}_x0"] = x.copy(), x0.copy()
        d[f"diffuse{b}_ab"] = np.array([alpha, beta], np.float32)
        R.diffuse(b, x, x0, alpha, beta); d[f"diffuse{b}_out"] = x
    for b in (0, 1, 2):
        dd, d0, u, v = rnd(), rnd(), rnd(-6, 6), rnd(-6, 6)   # dt0 = 0.48: up to ~3 cells, hits clamps
        d[f"advect{b}_d0"], d[f"advect{b}_u"], d[f"advect{b}_v"] = d0.copy(), u.copy(), v.copy()
        R.advect(b, dd, d0, u, v); d[f"advect{b}_out"] = dd
    u, v, p, div = rnd(), rnd(), rnd(), rnd()
    d["div_u"], d["div_v"] = u.copy(), v.copy()
    R.computeDivergenceAndPressure(u, v, p, div); d["div_p_out"], d["div_div_out"] = p, div
    u, v, p, div = rnd(), rnd(), rnd(), rnd()
    d["lp_u"], d["lp_v"], d["lp_p"] = u.copy(), v.copy(), p.copy()
    R.lastProject(u, v, p, div); d["lp_u_out"], d["lp_v_out"] = u, v
    x, x0, u, v = rnd(0, 1), rnd(0, 1), rnd(), rnd()
    d["dens_x"], d["dens_x0"], d["dens_u"], d["dens_v"] = x.copy(), x0.copy(), u.copy(), v.copy()
    R.dens_step(x, x0, u, v, R.DIFF); d["dens_x_out"], d["dens_x0_out"] = x, x0
    u, v, u0, v0 = rnd(), rnd(), rnd(), rnd()
    d["vel_u"], d["vel_v"], d["vel_u0"], d["vel_v0"] = u.copy(), v.copy(), u0.copy(), v0.copy()
    R.vel_step(u, v, u0, v0, R.VIS, 0)
    d["vel_u_out"], d["vel_v_out"], d["vel_u0_out"], d["vel_v0_out"] = u, v, u0, v0
    d["meta_N_K"] = np.array([N, K], np.int32)
    np.savez_compressed(os.path.join(OUT, f"stage_N{N}.npz"), **d)


def run_fixture(N, K, record_steps):
    R = ReferenceSeq(N, K)
    s = R.initializeParameters()
    d = {"meta_N_K": np.array([N, K], np.int32), "steps": np.array(record_steps, np.int32)}
    for f in ("dens_prev", "u_prev", "v_prev"):
        d[f"ic_{f}"] = s[f].copy()
    done = 0
    for upto in record_steps:
        R.run_steps(upto - done, s, first_step=done)
        done = upto
        for f in ("dens", "u", "v"):
            d[f"{f}_step{upto}"] = s[f].copy()
    np.savez_compressed(os.path.join(OUT, f"run_N{N}_K{K}.npz"), **d)


if __name__ == "__main__":
    stage_fixture()
    run_fixture(14, 40, [1, 2, 5])
    run_fixture(62, 20, [1, 3, 10])
    run_fixture(126, 20, [1, 10, 100])       # BASELINE config 1 (G=128, 20 iterations, 100 steps)
    run_fixture(126, 40, [1, 5])             # the reference's own iteration count
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
