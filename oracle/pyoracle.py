"""ctypes bindings for the CPU oracle (TEST INFRASTRUCTURE -- see stam_oracle.c header).

Two things are bound here:

* ``Oracle``      -- oracle/libstam_oracle[_omp].so, the parametrised restatement of
                     /root/reference/project/sequential/FluidSequential.c.
* ``RedBlackCheck``-- oracle/librbgs_check.so, a CPU build of the product's opt-in red-black solver (ours, not
                     the reference's; the restatement is compiled into it a second time for the step functions).
* ``ReferenceSeq``-- oracle/_ref/libref_seq_N<N>_K<K>.so, the reference's own translation unit
                     compiled from where it lies (oracle/Makefile target ``ref``); N and the
                     iteration count are compile-time literals there, so one library per (N, K).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FP = C.POINTER(C.c_float)


def _ptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(FP)


def build(force: bool = False) -> None:
    """Compile the restatement (and, when /root/reference is present, the reference builds)."""
    have = all(os.path.exists(os.path.join(HERE, n)) for n in ("libstam_oracle.so", "libstam_oracle_omp.so", "librbgs_check.so"))
    if force or not have:
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.isdir("/root/reference/project"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class Oracle:
    """Parametrised restatement.  ``threads=True`` loads the OpenMP build (identical results)."""

    LIBNAME = None    # subclasses load another build of the same entry points

    def __init__(self, threads: bool = False):
        name = self.LIBNAME or ("libstam_oracle_omp.so" if threads else "libstam_oracle.so")
        path = os.path.join(HERE, name)
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        i, f, u64 = C.c_int, C.c_float, C.c_uint64
        L.so_set_bnd.argtypes = [i, i, FP]
        L.so_add_source.argtypes = [i, FP, FP, f]
        L.so_lin_solve.argtypes = [i, i, FP, FP, f, f, i]
        L.so_advect.argtypes = [i, i, FP, FP, FP, FP, f]
        L.so_compute_divergence_and_pressure.argtypes = [i, FP, FP, FP, FP]
        L.so_last_project.argtypes = [i, FP, FP, FP, FP]
        L.so_dens_step.argtypes = [i, FP, FP, FP, FP, f, f, i]
        L.so_vel_step.argtypes = [i, FP, FP, FP, FP, f, f, i]
        L.so_init_reference_rand.argtypes = [i] + [FP] * 6
        L.so_init_synthetic.argtypes = [i, u64] + [FP] * 6
        L.so_run_steps.argtypes = [i, i, i] + [FP] * 6 + [f, f, f, i]
        for fn in ("so_set_bnd", "so_add_source", "so_lin_solve", "so_advect",
                   "so_compute_divergence_and_pressure", "so_last_project", "so_dens_step",
                   "so_vel_step", "so_init_reference_rand", "so_init_synthetic", "so_run_steps"):
            getattr(L, fn).restype = None
        self.L = L
        self.threads = threads

    # stage functions (names follow the reference)
    def set_bnd(self, N, b, x): self.L.so_set_bnd(N, b, _ptr(x))
    def add_source(self, N, x, s, dt): self.L.so_add_source(N, _ptr(x), _ptr(s), dt)
    def diffuse(self, N, b, x, x0, alpha, beta, iters): self.L.so_lin_solve(N, b, _ptr(x), _ptr(x0), alpha, beta, iters)
    def advect(self, N, b, d, d0, u, v, dt): self.L.so_advect(N, b, _ptr(d), _ptr(d0), _ptr(u), _ptr(v), dt)
    def computeDivergenceAndPressure(self, N, u, v, p, div):
        self.L.so_compute_divergence_and_pressure(N, _ptr(u), _ptr(v), _ptr(p), _ptr(div))
    def lastProject(self, N, u, v, p, div): self.L.so_last_project(N, _ptr(u), _ptr(v), _ptr(p), _ptr(div))
    def dens_step(self, N, x, x0, u, v, diff, dt, iters):
        self.L.so_dens_step(N, _ptr(x), _ptr(x0), _ptr(u), _ptr(v), diff, dt, iters)
    def vel_step(self, N, u, v, u0, v0, visc, dt, iters):
        self.L.so_vel_step(N, _ptr(u), _ptr(v), _ptr(u0), _ptr(v0), visc, dt, iters)

    def init_reference_rand(self, N):
        G = N + 2
        f = [np.empty((G, G), np.float32) for _ in range(6)]
        self.L.so_init_reference_rand(N, *[_ptr(a) for a in f])
        return dict(zip(("dens", "dens_prev", "u", "u_prev", "v", "v_prev"), f))

    def init_synthetic(self, N, seed=1):
        G = N + 2
        f = [np.empty((G, G), np.float32) for _ in range(6)]
        self.L.so_init_synthetic(N, seed, *[_ptr(a) for a in f])
        return dict(zip(("dens", "dens_prev", "u", "u_prev", "v", "v_prev"), f))

    def run_steps(self, N, steps, s, visc, diff, dt, iters, first_step=0):
        self.L.so_run_steps(N, steps, first_step, _ptr(s["dens"]), _ptr(s["dens_prev"]), _ptr(s["u"]),
                            _ptr(s["u_prev"]), _ptr(s["v"]), _ptr(s["v_prev"]), visc, diff, dt, iters)


class RedBlackCheck(Oracle):
    """CPU build of the product's OPT-IN red-black Gauss-Seidel / SOR solver (oracle/rbgs_check.c) -- not a
    reference path.  Same entry points as ``Oracle``; ``set_solver(1, omega)`` routes the solves inside
    dens_step / vel_step / run_steps through the red-black scheme, ``set_solver(0)`` back to the reference's Jacobi."""

    LIBNAME = "librbgs_check.so"

    def __init__(self):
        super().__init__(threads=True)
        i, f = C.c_int, C.c_float
        self.L.rb_set_solver.argtypes = [i, f]
        self.L.rb_set_solver.restype = None
        self.L.rb_lin_solve.argtypes = [i, i, FP, FP, f, f, i, f]
        self.L.rb_lin_solve.restype = None
        self.L.rb_residual_sumsq.argtypes = [i, FP, FP, f, f]
        self.L.rb_residual_sumsq.restype = C.c_double

    def set_solver(self, solver: int, omega: float = 1.0): self.L.rb_set_solver(solver, omega)
    def rb_diffuse(self, N, b, x, x0, alpha, beta, iters, omega=1.0):
        self.L.rb_lin_solve(N, b, _ptr(x), _ptr(x0), alpha, beta, iters, omega)
    def residual_sumsq(self, N, x, x0, alpha, beta) -> float:
        return float(self.L.rb_residual_sumsq(N, _ptr(x), _ptr(x0), alpha, beta))


class ReferenceSeq:
    """The reference's own sequential translation unit for one (N, K).

    Signatures as in FluidSequential.c:62,78,85,107,143,161,176,189,244.  DT/VIS/DIFF are the
    reference's literals (0.016f / 0.0025f / 0.1f, :7-9)."""

    DT, VIS, DIFF = 0.016, 0.0025, 0.1

    def __init__(self, N: int, K: int):
        path = os.path.join(HERE, "_ref", f"libref_seq_N{N}_K{K}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = C.CDLL(path)
        i, f = C.c_int, C.c_float
        L.set_bnd.argtypes = [i, FP]
        L.add_source.argtypes = [FP, FP]
        L.diffuse.argtypes = [i, FP, FP, f, f]
        L.advect.argtypes = [i, FP, FP, FP, FP]
        L.computeDivergenceAndPressure.argtypes = [FP] * 4
        L.lastProject.argtypes = [FP] * 4
        L.dens_step.argtypes = [FP, FP, FP, FP, f]
        L.vel_step.argtypes = [FP, FP, FP, FP, f, i]
        L.initializeParameters.argtypes = [FP] * 6
        for fn in ("set_bnd", "add_source", "diffuse", "advect", "computeDivergenceAndPressure",
                   "lastProject", "dens_step", "vel_step", "initializeParameters"):
            getattr(L, fn).restype = None
        self.L, self.N, self.K = L, N, K

    @staticmethod
    def available(N: int, K: int) -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", f"libref_seq_N{N}_K{K}.so"))

    def set_bnd(self, b, x): self.L.set_bnd(b, _ptr(x))
    def add_source(self, x, s): self.L.add_source(_ptr(x), _ptr(s))
    def diffuse(self, b, x, x0, alpha, beta): self.L.diffuse(b, _ptr(x), _ptr(x0), alpha, beta)
    def advect(self, b, d, d0, u, v): self.L.advect(b, _ptr(d), _ptr(d0), _ptr(u), _ptr(v))
    def computeDivergenceAndPressure(self, u, v, p, div):
        self.L.computeDivergenceAndPressure(_ptr(u), _ptr(v), _ptr(p), _ptr(div))
    def lastProject(self, u, v, p, div): self.L.lastProject(_ptr(u), _ptr(v), _ptr(p), _ptr(div))
    def dens_step(self, x, x0, u, v, diff): self.L.dens_step(_ptr(x), _ptr(x0), _ptr(u), _ptr(v), diff)
    def vel_step(self, u, v, u0, v0, visc, z=0): self.L.vel_step(_ptr(u), _ptr(v), _ptr(u0), _ptr(v0), visc, z)

    def initializeParameters(self):
        """Fresh process-default rand() stream is needed for the reference IC; callers that
        want it bit-for-bit must call this before anything else has consumed rand()."""
        G = self.N + 2
        f = [np.empty((G, G), np.float32) for _ in range(6)]
        C.CDLL(None).srand(1)
        self.L.initializeParameters(*[_ptr(a) for a in f])
        return dict(zip(("dens", "dens_prev", "u", "u_prev", "v", "v_prev"), f))

    def run_steps(self, steps, s, first_step=0):
        """FluidSequential.c:289-312 loop body, driven from Python with the reference's functions."""
        for k in range(steps):
            if first_step + k > 0:
                s["u_prev"][...] = 0.0
                s["v_prev"][...] = 0.0
                s["dens_prev"][...] = 0.0
            self.vel_step(s["u"], s["v"], s["u_prev"], s["v_prev"], self.VIS, 0)
            self.dens_step(s["dens"], s["dens_prev"], s["u"], s["v"], self.DIFF)
