"""Host-side mirror of the reference's solver interface over the C ABI (include/stablefluids.h).

The reference (ArbiterMob/FluidSimulationCuda) is a set of single-file C/CUDA programs whose
solver surface is the functions ``set_bnd, add_source, diffuse, advect,
computeDivergenceAndPressure, lastProject, dens_step, vel_step``
(project/sequential/FluidSequential.c:62-241).  ``StableFluids`` exposes the same names with the
same argument order and in/out semantics; what the reference bakes in as macros (N, DT, the
literal 40 iterations) are constructor / call arguments here.

PyTorch is used only for device memory and streams: fields are float32 CUDA tensors of shape
(N+2, N+2) (row-major, ``x[row, col]`` = the reference's ``x[col + row*(N+2)]``) and every kernel
is enqueued on ``torch.cuda.current_stream()`` at construction time.  All arithmetic happens in
libstablefluids_b200.so (hand-written sm_100a CUDA).  There is no CPU fallback: constructing a
solver without the library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SF_LIBRARY") or os.path.join(PKG, "libstablefluids_b200.so")   # SF_LIBRARY: experimental builds

SF_OPT_ARITHMETIC = 1
SF_OPT_SWEEPS_PER_LAUNCH = 2
SF_OPT_USE_GRAPH = 3
SF_OPT_FORCE_GENERIC = 4
SF_OPT_CHUNK_ROWS = 5
SF_OPT_STAGING = 6
SF_OPT_WORK_STEALING = 7
SF_OPT_STEAL_COUNT = 8
SF_OPT_STEAL_SCOPE = 9
SF_OPT_PRESSURE_PLAN = 10
SF_OPT_SOLVER = 11
SF_OPT_SOR_OMEGA_MILLI = 12
SF_OPT_RBGS_BLOCKED = 13
SF_OPT_FUSE_SOURCES = 14
SF_OPT_WAVE_SKEW = 15
SF_OPT_ADVECT_TILE = 16
SF_OPT_ADVECT_TILE_COUNT = 17
SF_OPT_ADVECT_FALLBACK_COUNT = 18
SF_OPT_OVERLAP_SOLVES = 19
SF_OPT_STRIP_BALANCE = 20
STRICT, FAST = 0, 1
SOLVER_JACOBI, SOLVER_RBGS = 0, 1    # SF_OPT_SOLVER: the reference's Jacobi (default) / opt-in red-black Gauss-Seidel (SOR)

# every symbol include/stablefluids.h declares (tests/test_abi.py checks the library exports them)
ABI_SYMBOLS = [
    "sf_create", "sf_create_on_stream", "sf_create_slab", "sf_create_slab_own_stream", "sf_destroy", "sf_last_error_string",
    "sf_set_option", "sf_get_option", "sf_synchronize", "sf_set_stream", "sf_get_stream", "sf_launch_count", "sf_field_bytes",
    "sf_alloc_field", "sf_free_field", "sf_upload", "sf_download",
    "sf_set_bnd", "sf_add_source", "sf_diffuse", "sf_advect", "sf_advect_velocity", "sf_compute_divergence_and_pressure",
    "sf_last_project", "sf_project", "sf_dens_step", "sf_vel_step", "sf_step", "sf_step_host",
    "sf_run_steps", "sf_dump_field", "sf_init_synthetic", "sf_init_sources", "sf_reduce_max_abs", "sf_reduce_max_abs_async", "sf_residual_l2",
    "sf_division_check", "sf_halo_rows_needed", "sf_jacobi_launch",
    "sf_slab_arena_create", "sf_slab_field", "sf_slab_ipc_handle", "sf_slab_connect_ipc", "sf_slab_connect_local",
    "sf_slab_set_timeout_ms", "sf_slab_status",
]
SOURCES_REFERENCE, SOURCES_SYNTHETIC, SOURCES_FIELDS = 0, 1, 2
SF_SLAB_UP, SF_SLAB_DOWN = 0, 1
SF_SLAB_ERR_TIMEOUT, SF_SLAB_ERR_REACH = 1, 2

_lib = None


class StableFluidsError(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """Load libstablefluids_b200.so (built in-tree by fluidsimulationcuda_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StableFluidsError(
            f"{LIB_PATH} is missing: run `python -m fluidsimulationcuda_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    if os.environ.get("SF_LIBRARY"):
        # experimental / older builds (A/B timing): symbols they lack resolve to a stub that raises
        class _Tolerant:
            def __init__(self, lib): object.__setattr__(self, "_lib", lib)
            def __getattr__(self, name):
                try:
                    return getattr(self._lib, name)
                except AttributeError:
                    def missing(*a, **k):
                        raise StableFluidsError(f"{LIB_PATH} does not export {name}")
                    return missing
        L = _Tolerant(L)
    vp, i, f, u64 = C.c_void_p, C.c_int, C.c_float, C.c_uint64
    L.sf_create.argtypes = [C.POINTER(vp), i, i]
    L.sf_create_on_stream.argtypes = [C.POINTER(vp), i, i, vp]
    L.sf_create_slab.argtypes = [C.POINTER(vp), i, i, vp, i, i, i]
    L.sf_create_slab_own_stream.argtypes = [C.POINTER(vp), i, i, i, i, i]
    L.sf_destroy.argtypes = [vp]
    L.sf_last_error_string.argtypes = [vp]
    L.sf_last_error_string.restype = C.c_char_p
    L.sf_set_option.argtypes = [vp, i, i]
    L.sf_get_option.argtypes = [vp, i, C.POINTER(i)]
    L.sf_synchronize.argtypes = [vp]
    L.sf_set_stream.argtypes = [vp, vp]
    L.sf_get_stream.argtypes = [vp, C.POINTER(vp)]
    L.sf_launch_count.argtypes = [vp, C.POINTER(C.c_ulonglong)]
    L.sf_field_bytes.argtypes = [vp]
    L.sf_field_bytes.restype = C.c_size_t
    L.sf_alloc_field.argtypes = [vp, C.POINTER(vp)]
    L.sf_free_field.argtypes = [vp, vp]
    L.sf_upload.argtypes = [vp, vp, vp]
    L.sf_download.argtypes = [vp, vp, vp]
    L.sf_set_bnd.argtypes = [vp, i, vp]
    L.sf_add_source.argtypes = [vp, vp, vp, f]
    L.sf_diffuse.argtypes = [vp, i, vp, vp, f, f, i]
    L.sf_advect.argtypes = [vp, i, vp, vp, vp, vp, f]
    L.sf_advect_velocity.argtypes = [vp, vp, vp, vp, vp, f]
    L.sf_compute_divergence_and_pressure.argtypes = [vp, vp, vp, vp, vp]
    L.sf_last_project.argtypes = [vp, vp, vp, vp, vp]
    L.sf_project.argtypes = [vp, vp, vp, vp, vp, i]
    L.sf_dens_step.argtypes = [vp, vp, vp, vp, vp, f, f, i]
    L.sf_vel_step.argtypes = [vp, vp, vp, vp, vp, f, f, i]
    L.sf_step.argtypes = [vp] + [vp] * 6 + [f, f, f, i]
    L.sf_step_host.argtypes = [vp] + [vp] * 6 + [f, f, f, i, i]
    L.sf_run_steps.argtypes = [vp] + [vp] * 6 + [f, f, f, i, i, i, u64, vp, vp, vp]
    L.sf_dump_field.argtypes = [vp, vp, C.c_char_p]
    L.sf_init_synthetic.argtypes = [vp, u64] + [vp] * 6
    L.sf_init_sources.argtypes = [vp, u64] + [vp] * 3
    L.sf_reduce_max_abs.argtypes = [vp, vp, C.POINTER(f)]
    L.sf_reduce_max_abs_async.argtypes = [vp, vp, vp]
    L.sf_residual_l2.argtypes = [vp, vp, vp, f, f, C.POINTER(C.c_double)]
    L.sf_division_check.argtypes = [vp, f, C.POINTER(i)]
    L.sf_halo_rows_needed.argtypes = [vp, C.POINTER(i)]
    L.sf_jacobi_launch.argtypes = [vp, i, vp, vp, vp, f, f, i, i, i]
    L.sf_slab_arena_create.argtypes = [vp, i]
    L.sf_slab_field.argtypes = [vp, i, C.POINTER(vp)]
    L.sf_slab_ipc_handle.argtypes = [vp, C.c_char_p]
    L.sf_slab_connect_ipc.argtypes = [vp, i, C.c_char_p, i, i]
    L.sf_slab_connect_local.argtypes = [vp, i, vp]
    L.sf_slab_set_timeout_ms.argtypes = [vp, i]
    L.sf_slab_status.argtypes = [vp, C.POINTER(C.c_uint)]
    for name in ABI_SYMBOLS:
        if name not in ("sf_last_error_string", "sf_field_bytes"):
            getattr(L, name).restype = i
    _lib = L
    return L


class _DevicePointer:
    """Minimal __cuda_array_interface__ carrier so torch can view library-owned device memory."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2,
                                         "strides": None}


def _tensor_view(torch, ptr, shape, device):
    with torch.cuda.device(device):
        return torch.as_tensor(_DevicePointer(ptr, shape), device=f"cuda:{device}")


class StableFluids:
    """One solver context = the reference's N plus a CUDA stream.

    ``row_lo/row_hi/halo`` select a row slab for domain decomposition (see
    fluidsimulationcuda_b200.slab); the defaults are the whole grid."""

    def __init__(self, N: int, device: Optional[int] = None, *, row_lo: int = 0, row_hi: Optional[int] = None,
                 halo: int = 0, arithmetic: int = STRICT, sweeps_per_launch: int = 0, use_graph: bool = True,
                 own_stream: bool = False):
        import torch
        if not torch.cuda.is_available():
            raise StableFluidsError("no CUDA device: the stable-fluids path has no CPU fallback")
        self.torch = torch
        self.L = load_library()
        self.N, self.G = int(N), int(N) + 2
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.row_lo = int(row_lo)
        self.row_hi = self.G if row_hi is None else int(row_hi)
        self.halo = int(halo)
        self.local_rows = self.row_hi - self.row_lo + 2 * self.halo
        h = C.c_void_p()
        if own_stream:
            # the context creates (and owns) a non-blocking stream; torch sees it as an ExternalStream
            rc = self.L.sf_create_slab_own_stream(C.byref(h), self.N, self.device, self.row_lo, self.row_hi, self.halo)
            if rc != 0:
                raise StableFluidsError(f"sf_create_slab_own_stream failed with status {rc}")
            sp = C.c_void_p()
            self.L.sf_get_stream(h, C.byref(sp))
            self._stream = torch.cuda.ExternalStream(sp.value, device=f"cuda:{self.device}")
        else:
            self._stream = torch.cuda.current_stream(self.device)
            rc = self.L.sf_create_slab(C.byref(h), self.N, self.device, C.c_void_p(self._stream.cuda_stream),
                                       self.row_lo, self.row_hi, self.halo)
            if rc != 0:
                raise StableFluidsError(f"sf_create_slab failed with status {rc}")
        self.h = h
        self.set_option(SF_OPT_ARITHMETIC, arithmetic)
        self.set_option(SF_OPT_SWEEPS_PER_LAUNCH, sweeps_per_launch)
        self.set_option(SF_OPT_USE_GRAPH, 1 if use_graph else 0)

    # -- plumbing ----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.L.sf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            msg = self.L.sf_last_error_string(self.h)
            raise StableFluidsError(f"status {rc}: {msg.decode() if msg else '?'}")

    def _p(self, t):
        torch = self.torch
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise StableFluidsError("fields must be contiguous float32 CUDA tensors")
        if t.numel() != self.local_rows * self.G:
            raise StableFluidsError(f"field has {t.numel()} cells, expected {self.local_rows}x{self.G}")
        return C.c_void_p(t.data_ptr())

    def set_option(self, opt: int, value: int):
        self._check(self.L.sf_set_option(self.h, opt, int(value)))

    def get_option(self, opt: int) -> int:
        v = C.c_int(0)
        self._check(self.L.sf_get_option(self.h, opt, C.byref(v)))
        return int(v.value)

    def new_field(self):
        return self.torch.zeros((self.local_rows, self.G), dtype=self.torch.float32, device=f"cuda:{self.device}")

    def synchronize(self):
        self._check(self.L.sf_synchronize(self.h))

    @property
    def launch_count(self) -> int:
        n = C.c_ulonglong(0)
        self._check(self.L.sf_launch_count(self.h, C.byref(n)))
        return int(n.value)

    # -- the reference's stage functions (FluidSequential.c:62-173) ---------------------------
    def set_bnd(self, b, x):
        self._check(self.L.sf_set_bnd(self.h, b, self._p(x)))

    def add_source(self, x, s, dt):
        self._check(self.L.sf_add_source(self.h, self._p(x), self._p(s), dt))

    def diffuse(self, b, x, x0, alpha, beta, iters):
        self._check(self.L.sf_diffuse(self.h, b, self._p(x), self._p(x0), alpha, beta, iters))

    def advect(self, b, d, d0, u, v, dt):
        self._check(self.L.sf_advect(self.h, b, self._p(d), self._p(d0), self._p(u), self._p(v), dt))

    def advect_velocity(self, u, v, u0, v0, dt):
        """advect(1, u, u0, u0, v0); advect(2, v, v0, u0, v0) in one pass (FluidSequential.c:228-237)."""
        self._check(self.L.sf_advect_velocity(self.h, self._p(u), self._p(v), self._p(u0), self._p(v0), dt))

    def computeDivergenceAndPressure(self, u, v, p, div):
        self._check(self.L.sf_compute_divergence_and_pressure(self.h, self._p(u), self._p(v), self._p(p), self._p(div)))

    def lastProject(self, u, v, p, div):
        self._check(self.L.sf_last_project(self.h, self._p(u), self._p(v), self._p(p), self._p(div)))

    def project(self, u, v, p, div, iters):
        self._check(self.L.sf_project(self.h, self._p(u), self._p(v), self._p(p), self._p(div), iters))

    # -- step drivers (FluidSequential.c:176-241, :305-306) -----------------------------------
    def dens_step(self, x, x0, u, v, diff, dt, iters):
        self._check(self.L.sf_dens_step(self.h, self._p(x), self._p(x0), self._p(u), self._p(v), diff, dt, iters))

    def vel_step(self, u, v, u0, v0, visc, dt, iters):
        self._check(self.L.sf_vel_step(self.h, self._p(u), self._p(v), self._p(u0), self._p(v0), visc, dt, iters))

    def step(self, dens, dens_prev, u, u_prev, v, v_prev, visc, diff, dt, iters):
        self._check(self.L.sf_step(self.h, self._p(dens), self._p(dens_prev), self._p(u), self._p(u_prev),
                                   self._p(v), self._p(v_prev), visc, diff, dt, iters))

    def run_steps(self, dens, dens_prev, u, u_prev, v, v_prev, visc, diff, dt, iters, steps, sources=SOURCES_REFERENCE,
                  seed=0, src_dens=None, src_u=None, src_v=None):
        """The reference's main loop (FluidSequential.c:289-312) resident on the device: `steps` iterations of
        { source schedule; vel_step; dens_step } with no host round trip in between."""
        opt = lambda t: C.c_void_p(0) if t is None else self._p(t)
        self._check(self.L.sf_run_steps(self.h, self._p(dens), self._p(dens_prev), self._p(u), self._p(u_prev), self._p(v),
                                        self._p(v_prev), visc, diff, dt, iters, steps, sources, seed, opt(src_dens),
                                        opt(src_u), opt(src_v)))

    def dump_field(self, field, path: str):
        """Binary dump (32-byte header + owned rows as float32) -- replaces the reference's printStateGrid."""
        self._check(self.L.sf_dump_field(self.h, self._p(field), str(path).encode()))

    def step_host(self, dens, dens_prev, u, u_prev, v, v_prev, visc, diff, dt, iters, download_scratch=False):
        """Same loop body with HOST fields (numpy float32 arrays or CPU tensors, ideally pinned)."""
        cells = (self.row_hi - self.row_lo) * self.G      # a slab context exchanges its owned rows only

        def hp(a):
            if hasattr(a, "data_ptr"):
                assert not a.is_cuda and a.is_contiguous() and a.numel() == cells
                return C.c_void_p(a.data_ptr())
            assert a.dtype.name == "float32" and a.flags["C_CONTIGUOUS"] and a.size == cells
            return C.c_void_p(a.ctypes.data)
        self._check(self.L.sf_step_host(self.h, hp(dens), hp(dens_prev), hp(u), hp(u_prev), hp(v), hp(v_prev),
                                        visc, diff, dt, iters, 1 if download_scratch else 0))

    # -- synthetic data and diagnostics ------------------------------------------------------
    def init_synthetic(self, seed, dens, dens_prev, u, u_prev, v, v_prev):
        self._check(self.L.sf_init_synthetic(self.h, seed, self._p(dens), self._p(dens_prev), self._p(u),
                                             self._p(u_prev), self._p(v), self._p(v_prev)))

    def init_sources(self, seed, dens_prev, u_prev, v_prev):
        self._check(self.L.sf_init_sources(self.h, seed, self._p(dens_prev), self._p(u_prev), self._p(v_prev)))

    def reduce_max_abs(self, x) -> float:
        out = C.c_float(0)
        self._check(self.L.sf_reduce_max_abs(self.h, self._p(x), C.byref(out)))
        return float(out.value)

    def reduce_max_abs_async(self, x, dev_scalar):
        """dev_scalar (1-element float32 CUDA tensor) = max(dev_scalar, max|x|); no synchronisation."""
        self._check(self.L.sf_reduce_max_abs_async(self.h, self._p(x), C.c_void_p(dev_scalar.data_ptr())))

    def set_stream(self, stream):
        """Order all later calls on `stream` (a torch.cuda.Stream)."""
        self._stream = stream
        self._check(self.L.sf_set_stream(self.h, C.c_void_p(stream.cuda_stream)))

    def residual_sumsq(self, x, x0, alpha, beta) -> float:
        out = C.c_double(0)
        self._check(self.L.sf_residual_l2(self.h, self._p(x), self._p(x0), alpha, beta, C.byref(out)))
        return float(out.value)

    def division_check(self, beta) -> bool:
        out = C.c_int(0)
        self._check(self.L.sf_division_check(self.h, beta, C.byref(out)))
        return bool(out.value)

    # -- peer-memory slabs (include/stablefluids.h, "peer-memory slabs") -------------------------
    def arena_create(self, nfields: int):
        """Allocate this slab's fields in one peer-mappable device allocation; returns the fields as
        torch tensors (views of library-owned memory, valid until close())."""
        self._check(self.L.sf_slab_arena_create(self.h, int(nfields)))
        out = []
        for k in range(nfields):
            p = C.c_void_p()
            self._check(self.L.sf_slab_field(self.h, k, C.byref(p)))
            out.append(_tensor_view(self.torch, p.value, (self.local_rows, self.G), self.device))
        return out

    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self.L.sf_slab_ipc_handle(self.h, buf))
        return bytes(buf.raw)

    def connect_ipc(self, direction: int, handle: bytes, nbr_row_lo: int, nbr_row_hi: int):
        assert len(handle) == 64
        self._check(self.L.sf_slab_connect_ipc(self.h, direction, handle, nbr_row_lo, nbr_row_hi))

    def connect_local(self, direction: int, neighbour: "StableFluids"):
        self._check(self.L.sf_slab_connect_local(self.h, direction, neighbour.h))

    def set_slab_timeout_ms(self, ms: int):
        self._check(self.L.sf_slab_set_timeout_ms(self.h, int(ms)))

    def slab_status(self) -> int:
        """Synchronise and return the sticky device-side error bits (0 = fine)."""
        bits = C.c_uint(0)
        self._check(self.L.sf_slab_status(self.h, C.byref(bits)))
        return int(bits.value)

    def jacobi_launch(self, b, xout, xin, x0, alpha, beta, sweeps, out_lo=-1, out_hi=-1):
        self._check(self.L.sf_jacobi_launch(self.h, b, self._p(xout), self._p(xin), self._p(x0), alpha, beta,
                                            sweeps, out_lo, out_hi))
