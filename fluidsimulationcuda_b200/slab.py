"""Row-slab domain decomposition of the stable-fluids step over the GPUs of one box.

The reference has no multi-GPU path; this is the north star's slab scheme:
rank r owns global rows [r*G/p, (r+1)*G/p) of every field and stores them with `halo` ghost rows
above and below.  Every stage is a radius-1 stencil or a bounded-reach gather, so the only
communication is a NEIGHBOUR exchange of a few contiguous rows:

  lin_solve  T rows of the iterate per temporally blocked launch of T sweeps (plus T rows of the
             right-hand side once per solve); the two T-row boundary strips are computed first,
             their exchange is put in flight (NCCL send/recv on NCCL's own stream) and the interior
             launch runs meanwhile;
  divergence / gradient subtract   1 row;
  advect     W = ceil(dt*N*max|vel|) + 2 rows of the advected field, W from a MAX all-reduce.

Results are independent of the partition (no reduction enters the update): every p gives the
bit-identical fields of the single-GPU path (tests/test_slab_gpu.py).

The arithmetic runs in libstablefluids_b200.so through slab contexts (sf_create_slab,
sf_jacobi_launch); this module only sequences launches and exchanges.  ``SlabSolver.step_gen`` is a
generator that yields the communication requests, so that the same code is driven by NCCL
(one process per GPU, ``TorchDistComm``) or in lock-step inside one process for tests
(``run_lockstep``: p emulated ranks on one GPU, halos copied directly).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import solver as SF


def partition_rows(G: int, world: int) -> List[Tuple[int, int]]:
    """Owned row ranges [lo, hi) per rank: contiguous, covering [0, G), sizes differ by at most 1."""
    if world < 1 or G < 2 * world:
        raise ValueError("need at least two rows per rank")
    return [(r * G // world, (r + 1) * G // world) for r in range(world)]


def plan_launches(iters: int, T: int) -> List[int]:
    """Split `iters` sweeps into launches of at most T sweeps; an even number of launches leaves
    the result of the x <-> scratch ping-pong in x (mirrors plan_launches in csrc/sf_api.cu)."""
    L = (iters + T - 1) // T
    if (L & 1) and L + 1 <= iters:
        L += 1
    plan = [iters // L] * L
    for k in range(iters % L):
        plan[k] += 1
    return plan


def f32_coeffs(dt: float, coef: float, N: int) -> Tuple[float, float]:
    """alpha = dt*coef*N*N and beta = 1 + 4*alpha evaluated left to right in binary32
    (FluidSequential.c:179-180, :199-200)."""
    f = np.float32
    a = f(dt) * f(coef)
    a = a * f(N)
    a = a * f(N)
    return float(a), float(f(1) + f(4) * a)


class HaloSpec:
    """`rows` ghost rows of `field` are to be refreshed from the neighbours' owned rows."""
    __slots__ = ("field", "rows")

    def __init__(self, field, rows: int):
        self.field, self.rows = field, int(rows)


class SlabLayout:
    """Which rows a rank owns and where its ghost rows sit in a local field of
    (own_rows + 2*halo, G) cells.  Pure index arithmetic (no device needed)."""

    def __init__(self, G: int, rank: int, world: int, halo: int):
        self.G, self.rank, self.world, self.halo = G, rank, world, halo
        self.row_lo, self.row_hi = partition_rows(G, world)[rank]

    @property
    def own_rows(self) -> int:
        return self.row_hi - self.row_lo

    def local(self, global_row: int) -> int:
        return global_row - (self.row_lo - self.halo)

    def owned(self, field):
        return field[self.halo: self.halo + self.own_rows]

    # contiguous row blocks that travel
    def send_up(self, field, h):      # my first h owned rows -> rank-1's lower ghost rows
        return field[self.halo: self.halo + h]

    def send_down(self, field, h):    # my last h owned rows -> rank+1's upper ghost rows
        e = self.halo + self.own_rows
        return field[e - h: e]

    def recv_up(self, field, h):      # ghost rows above my slab <- rank-1's last h owned rows
        return field[self.halo - h: self.halo]

    def recv_down(self, field, h):    # ghost rows below my slab <- rank+1's first h owned rows
        e = self.halo + self.own_rows
        return field[e: e + h]


class SlabSolver(SlabLayout):
    """One rank's slab.  Fields are local tensors of (own_rows + 2*halo, G)."""

    def __init__(self, N: int, rank: int, world: int, *, iters: int = 40, halo: int = 0, arithmetic: int = SF.STRICT,
                 sweeps_per_launch: int = 7, device: Optional[int] = None, comm=None, allocate: bool = True,
                 overlap: bool = True, deferred_reach: bool = True):
        SlabLayout.__init__(self, N + 2, rank, world, 0)
        self.N = N
        self.iters = iters
        self.T = sweeps_per_launch
        own = self.row_hi - self.row_lo
        if halo <= 0:
            # room for the advection reach: ~1/8 of a slab, at least 64 rows, never more than a slab
            halo = max(64, min(own, 1024, (own // 8 + 7) // 8 * 8))
        if world == 1:
            halo = 0
        if world > 1 and halo < self.T:
            raise ValueError("halo must cover the temporal-blocking depth")
        if world > 1 and own < 2 * self.T:
            raise ValueError("slab thinner than two boundary strips")
        self.halo = halo
        self.ctx = SF.StableFluids(N, device, row_lo=self.row_lo, row_hi=self.row_hi, halo=halo,
                                   arithmetic=arithmetic, sweeps_per_launch=sweeps_per_launch, use_graph=(world == 1))
        self.comm = comm
        self.names = ("dens", "dens_prev", "u", "u_prev", "v", "v_prev")
        if allocate:
            self.f = {k: self.ctx.new_field() for k in self.names}
            self.scratch = self.ctx.new_field()
        # Boundary strips run on a high-priority side stream (through a second context bound to
        # it) so that they overlap the interior launch; `overlap=False` keeps everything on one
        # stream (lock-step emulation of several ranks in one process).
        self.overlap = bool(overlap) and world > 1
        self.side = self.ctx_hi = None
        if self.overlap:
            torch = self.ctx.torch
            self.main = torch.cuda.current_stream(self.ctx.device)
            self.side = torch.cuda.Stream(device=self.ctx.device, priority=-1)
            with torch.cuda.stream(self.side):
                self.ctx_hi = SF.StableFluids(N, device, row_lo=self.row_lo, row_hi=self.row_hi, halo=halo,
                                              arithmetic=arithmetic, sweeps_per_launch=sweeps_per_launch, use_graph=False)
        # Advection reach: checked one step late so that no host synchronisation sits inside a step
        # (deferred=True); the exchange then always carries the full ghost depth.
        self.deferred_reach = bool(deferred_reach) and world > 1
        self._pending_reach = []     # [(device scalar max|vel| over all ranks, dt)] of the current step
        self._older_reach = []       # the same, per earlier step (oldest first)

    @property
    def launch_count(self) -> int:
        return self.ctx.launch_count + (self.ctx_hi.launch_count if self.ctx_hi is not None else 0)

    # ---- the step as a generator of communication requests ---------------------------------------
    # yields ("exchange", [HaloSpec...])            blocking neighbour exchange
    #        ("exchange_begin", [HaloSpec...])      start it, keep computing
    #        ("exchange_end", None)                 wait for the one in flight
    #        ("allreduce_max", value) -> send(max)  scalar MAX over ranks
    def _lin_solve(self, b, x, x0, alpha, beta):
        c, T = self.ctx, self.T
        plan = plan_launches(self.iters, T)
        lo, hi = self.row_lo, self.row_hi
        if self.world == 1:
            c.diffuse(b, x, x0, alpha, beta, self.iters)
            return
        yield ("exchange", [HaloSpec(x, plan[0]), HaloSpec(x0, T)], None)
        cur, nxt = x, self.scratch
        torch = c.torch
        for k, sweeps in enumerate(plan):
            need_next = plan[k + 1] if k + 1 < len(plan) else 1   # after the solve: 1 row for the stencils that follow
            strip = max(need_next, 1)
            # boundary strips (only the sides that have a neighbour) and their exchange, while the interior runs
            top_hi = lo + strip if self.rank > 0 else lo
            bot_lo = hi - strip if self.rank < self.world - 1 else hi
            if self.overlap:
                ev = torch.cuda.Event()
                ev.record(self.main)                 # everything this launch reads is complete on main
                self.side.wait_event(ev)
                with torch.cuda.stream(self.side):
                    if top_hi > lo:
                        self.ctx_hi.jacobi_launch(b, nxt, cur, x0, alpha, beta, sweeps, lo, top_hi)
                    if bot_lo < hi:
                        self.ctx_hi.jacobi_launch(b, nxt, cur, x0, alpha, beta, sweeps, bot_lo, hi)
                yield ("exchange_begin", [HaloSpec(nxt, strip)], self.side)
                c.jacobi_launch(b, nxt, cur, x0, alpha, beta, sweeps, top_hi, bot_lo)   # main stream, concurrently
                yield ("exchange_end", None, self.side)
                ev2 = torch.cuda.Event()
                ev2.record(self.side)                # strips written and ghost rows received
                self.main.wait_event(ev2)
            else:
                if top_hi > lo:
                    c.jacobi_launch(b, nxt, cur, x0, alpha, beta, sweeps, lo, top_hi)
                if bot_lo < hi:
                    c.jacobi_launch(b, nxt, cur, x0, alpha, beta, sweeps, bot_lo, hi)
                yield ("exchange_begin", [HaloSpec(nxt, strip)], None)
                c.jacobi_launch(b, nxt, cur, x0, alpha, beta, sweeps, top_hi, bot_lo)
                yield ("exchange_end", None, None)
            cur, nxt = nxt, cur
        if cur is not x:
            x.copy_(cur)

    def _project(self, u, v, p, div):
        c = self.ctx
        if self.world == 1:
            c.project(u, v, p, div, self.iters)
            return
        p.zero_()                                   # zero guess including the ghost rows
        c.computeDivergenceAndPressure(u, v, p, div)  # u, v ghost rows (1) are valid on entry
        yield from self._lin_solve(0, p, div, 1.0, 4.0)
        c.lastProject(u, v, p, div)                 # p ghost row valid after the solve's last exchange

    def _advect_reach(self, dt, *vel):
        """Ghost rows the gather can touch: |dt*N*v| rounded up, +2 for the bilinear footprint.
        deferred: exchange the whole ghost depth now, verify the reach one step later (no host
        synchronisation inside the step); otherwise a scalar MAX all-reduce read back on the host."""
        if self.deferred_reach:
            m = self.ctx.torch.zeros(1, dtype=self.ctx.torch.float32, device=vel[0].device)
            for t in vel:
                self.ctx.reduce_max_abs_async(t, m)
            yield ("allreduce_max_device", m, None)
            self._pending_reach.append((m, dt))
            return self.halo
        m = 0.0
        for t in vel:
            m = max(m, self.ctx.reduce_max_abs(t))
        m = yield ("allreduce_max", m, None)
        return self._reach_rows(m, dt)

    def check_reach(self):
        """Resolve the deferred advection-reach checks (synchronises); raises if a back-trace could
        have left the ghost rows."""
        steps = self._older_reach + [self._pending_reach]
        self._older_reach, self._pending_reach = [], []
        for pending in steps:
            for m, dt in pending:
                self._reach_rows(float(m.item()), dt)

    def step_gen(self, visc: float, diff: float, dt: float):
        c, f = self.ctx, self.f
        u, v, u0, v0, d, d0 = f["u"], f["v"], f["u_prev"], f["v_prev"], f["dens"], f["dens_prev"]
        multi = self.world > 1
        # ---- vel_step (FluidSequential.c:189-241) ----
        c.add_source(u, u0, dt)
        c.add_source(v, v0, dt)
        alpha, beta = f32_coeffs(dt, visc, self.N)
        yield from self._lin_solve(1, u0, u, alpha, beta)
        yield from self._lin_solve(2, v0, v, alpha, beta)
        yield from self._project(u0, v0, u, v)
        if multi:
            W = yield from self._advect_reach(dt, u0, v0)
            yield ("exchange", [HaloSpec(u0, W), HaloSpec(v0, W)], None)
        c.advect(1, u, u0, u0, v0, dt)
        c.advect(2, v, v0, u0, v0, dt)
        if multi:
            yield ("exchange", [HaloSpec(u, 1), HaloSpec(v, 1)], None)
        yield from self._project(u, v, u0, v0)
        # ---- dens_step (:176-186) ----
        c.add_source(d, d0, dt)
        alpha, beta = f32_coeffs(dt, diff, self.N)
        yield from self._lin_solve(0, d0, d, alpha, beta)
        if multi:
            W = yield from self._advect_reach(dt, u, v)
            yield ("exchange", [HaloSpec(d0, W)], None)
        c.advect(0, d, d0, u, v, dt)

    def _reach_rows(self, max_vel: float, dt: float) -> int:
        dt0 = float(np.float32(dt) * np.float32(self.N))
        if not math.isfinite(max_vel):      # sf_reduce_max_abs propagates NaN / inf velocities
            raise SF.StableFluidsError(f"advection reach: max|velocity| is {max_vel}; the fields are corrupt")
        W = int(math.ceil(dt0 * max_vel)) + 2
        if W > self.halo:
            raise SF.StableFluidsError(
                f"advection reaches {W} rows beyond the slab but only {self.halo} ghost rows are allocated; "
                f"create the SlabSolver with halo >= {W}")
        return max(W, 1)

    # ---- drivers -----------------------------------------------------------------------------
    def init_synthetic(self, seed: int):
        self.ctx.init_synthetic(seed, *[self.f[k] for k in self.names])

    def step(self, seed: Optional[int], visc: float, diff: float, dt: float):
        """One loop-body iteration driven by this rank's communicator (NCCL in production)."""
        if seed is not None:
            self.ctx.init_sources(seed, self.f["dens_prev"], self.f["u_prev"], self.f["v_prev"])
        if self.world == 1:
            for _ in self.step_gen(visc, diff, dt):
                raise AssertionError("single-slab steps do not communicate")
            return
        if self._pending_reach:
            self._older_reach.append(self._pending_reach)
            self._pending_reach = []
        while len(self._older_reach) > 1:      # two steps old: that work finished long ago, no stall
            for m, dt_old in self._older_reach.pop(0):
                self._reach_rows(float(m.item()), dt_old)
        gen = self.step_gen(visc, diff, dt)
        reply = None
        while True:
            try:
                kind, arg, stream = gen.send(reply)
            except StopIteration:
                break
            reply = self.comm.serve(self, kind, arg, stream)


class TorchDistComm:
    """Neighbour exchange over torch.distributed point-to-point ops (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.pending = None

    def _post(self, s: SlabLayout, specs: Sequence[HaloSpec]):
        dist = self.dist
        ops = []
        for sp in specs:
            h = sp.rows
            if h <= 0:
                continue
            if s.rank > 0:
                ops.append(dist.P2POp(dist.isend, s.send_up(sp.field, h), s.rank - 1, self.group))
                ops.append(dist.P2POp(dist.irecv, s.recv_up(sp.field, h), s.rank - 1, self.group))
            if s.rank < s.world - 1:
                ops.append(dist.P2POp(dist.isend, s.send_down(sp.field, h), s.rank + 1, self.group))
                ops.append(dist.P2POp(dist.irecv, s.recv_down(sp.field, h), s.rank + 1, self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    def serve(self, s: SlabLayout, kind: str, arg, stream=None):
        """`stream`: the CUDA stream the request is ordered on (None = the current stream)."""
        if stream is not None:
            import torch
            with torch.cuda.stream(stream):
                return self.serve(s, kind, arg, None)
        if kind == "exchange":
            for r in self._post(s, arg):
                r.wait()
        elif kind == "exchange_begin":
            self.pending = self._post(s, arg)
        elif kind == "exchange_end":
            for r in self.pending or []:
                r.wait()
            self.pending = None
        elif kind == "allreduce_max_device":
            self.dist.all_reduce(arg, op=self.dist.ReduceOp.MAX, group=self.group)
        elif kind == "allreduce_max":
            import torch
            t = torch.tensor([arg], dtype=torch.float32, device=s.f["u"].device if hasattr(s, "f") else "cpu")
            if self.dist.get_backend(self.group) == "nccl" and not t.is_cuda:
                t = t.cuda()
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
            return float(t.item())
        else:
            raise ValueError(kind)
        return None


def exchange_lockstep(solvers: Sequence[SlabSolver], specs_per_rank: Sequence[Sequence[HaloSpec]]):
    """All ranks exchange the same list of fields (by position); direct copies between slabs."""
    for r, s in enumerate(solvers):
        for i, sp in enumerate(specs_per_rank[r]):
            h = sp.rows
            if h <= 0:
                continue
            if r > 0:
                nb = solvers[r - 1]
                s.recv_up(sp.field, h).copy_(nb.send_down(specs_per_rank[r - 1][i].field, h))
            if r < len(solvers) - 1:
                nb = solvers[r + 1]
                s.recv_down(sp.field, h).copy_(nb.send_up(specs_per_rank[r + 1][i].field, h))


def run_lockstep(solvers: Sequence[SlabSolver], visc: float, diff: float, dt: float):
    """Advance p emulated ranks (one process, one device) through one step in lock-step, serving
    their communication requests with direct copies.  Test-only driver."""
    gens = [s.step_gen(visc, diff, dt) for s in solvers]
    replies = [None] * len(gens)
    pending = None
    while True:
        reqs = []
        for g, rep in zip(gens, replies):
            try:
                reqs.append(g.send(rep))
            except StopIteration:
                reqs.append(None)
        if all(r is None for r in reqs):
            return
        assert all(r is not None for r in reqs), "ranks left the step at different points"
        assert all(r[2] is None for r in reqs), "lock-step emulation runs on one stream (overlap=False)"
        kinds = {r[0] for r in reqs}
        assert len(kinds) == 1, f"ranks diverged: {kinds}"
        kind = kinds.pop()
        replies = [None] * len(gens)
        if kind == "exchange":
            exchange_lockstep(solvers, [r[1] for r in reqs])
        elif kind == "exchange_begin":
            # every rank has enqueued its boundary strips (same device, stream order): copy now
            exchange_lockstep(solvers, [r[1] for r in reqs])
            pending = True
        elif kind == "exchange_end":
            pending = None
        elif kind == "allreduce_max":
            m = max(r[1] for r in reqs)
            replies = [m] * len(gens)
        elif kind == "allreduce_max_device":
            m = max(float(r[1].item()) for r in reqs)
            for r in reqs:
                r[1].fill_(m)
        else:
            raise ValueError(kind)


# =================================================================================================
# Peer-memory slabs: the product path on an NVLink / NVSwitch box
# =================================================================================================
def neighbour_links(G: int, rank: int, world: int, handles: Sequence[bytes]):
    """Which arenas rank `rank` maps, given every rank's 64-byte handle:
    [(SF_SLAB_UP | SF_SLAB_DOWN, handle, neighbour_row_lo, neighbour_row_hi)]."""
    parts = partition_rows(G, world)
    links = []
    if rank > 0:
        links.append((SF.SF_SLAB_UP, handles[rank - 1], *parts[rank - 1]))
    if rank < world - 1:
        links.append((SF.SF_SLAB_DOWN, handles[rank + 1], *parts[rank + 1]))
    return links


def exchange_handles(G: int, rank: int, world: int, my_handle: bytes, group=None):
    """All-gather the handles over torch.distributed (any backend) and return this rank's links."""
    import torch.distributed as dist
    if len(my_handle) != 64:
        raise ValueError("a CUDA IPC memory handle is 64 bytes")
    handles = [None] * world
    dist.all_gather_object(handles, bytes(my_handle), group=group)
    return neighbour_links(G, rank, world, handles)


class PeerSlabSolver(SlabLayout):
    """One rank's slab with its neighbours' memory mapped (CUDA IPC between processes, peer access
    inside one process).  The whole step -- halo pushes fused into the boundary strips of every
    temporally blocked Jacobi launch, peer-memory gathers in advect, one-warp neighbour barriers --
    runs inside libstablefluids_b200.so and is replayed from one CUDA graph per GPU
    (csrc/sf_slab.cu).  This class only allocates, wires the neighbours up and calls ``sf_step``;
    torch.distributed is used once, to pass the 64-byte IPC handles around.

    ``step`` is collective: every rank must call it the same number of times."""

    NAMES = ("dens", "dens_prev", "u", "u_prev", "v", "v_prev")

    def __init__(self, N: int, rank: int, world: int, *, iters: int = 40, halo: int = 8, arithmetic: int = SF.STRICT,
                 sweeps_per_launch: int = 0, device: Optional[int] = None, stream=None, use_graph: bool = True,
                 timeout_ms: Optional[int] = None):
        if world == 1:
            halo = 0
        SlabLayout.__init__(self, N + 2, rank, world, halo)
        self.N, self.iters, self.names = N, iters, self.NAMES
        depth = sweeps_per_launch or 8            # 0 = the library's default depth (at most 8)
        if world > 1 and halo < depth:
            raise ValueError("halo must cover the temporal-blocking depth")
        if world > 1 and self.own_rows < 2 * depth:
            raise ValueError("slab thinner than two boundary strips")
        import torch
        self.torch = torch
        dev = torch.cuda.current_device() if device is None else int(device)
        if stream is None:
            # the context's own non-blocking stream: never the legacy default stream, and (several slabs in
            # one process) one hardware queue per slab, so a waiting barrier kernel never sits in front of
            # the kernel it waits for
            self.ctx = SF.StableFluids(N, dev, row_lo=self.row_lo, row_hi=self.row_hi, halo=halo, arithmetic=arithmetic,
                                       sweeps_per_launch=sweeps_per_launch, use_graph=use_graph, own_stream=True)
            self.stream = self.ctx._stream
        else:
            self.stream = stream
            with torch.cuda.stream(self.stream):
                self.ctx = SF.StableFluids(N, dev, row_lo=self.row_lo, row_hi=self.row_hi, halo=halo, arithmetic=arithmetic,
                                           sweeps_per_launch=sweeps_per_launch, use_graph=use_graph)
        self.f = dict(zip(self.NAMES, self.ctx.arena_create(len(self.NAMES))))
        if timeout_ms is not None:
            self.ctx.set_slab_timeout_ms(timeout_ms)

    @property
    def launch_count(self) -> int:
        return self.ctx.launch_count

    # ---- wiring ------------------------------------------------------------------------------
    def connect_local(self, solvers: Sequence["PeerSlabSolver"]):
        """All ranks live in this process (`solvers[r]` is rank r)."""
        if self.rank > 0:
            self.ctx.connect_local(SF.SF_SLAB_UP, solvers[self.rank - 1].ctx)
        if self.rank < self.world - 1:
            self.ctx.connect_local(SF.SF_SLAB_DOWN, solvers[self.rank + 1].ctx)

    def connect_dist(self, group=None):
        """One process per GPU: all-gather the arenas' CUDA IPC handles and map ranks r-1 / r+1."""
        import torch.distributed as dist
        if self.world == 1:
            return
        for direction, handle, lo, hi in exchange_handles(self.G, self.rank, self.world, self.ctx.ipc_handle(), group):
            self.ctx.connect_ipc(direction, handle, lo, hi)
        dist.barrier(group=group)       # every arena is mapped before anyone enqueues a push

    # ---- drivers -----------------------------------------------------------------------------
    def init_synthetic(self, seed: int):
        with self.torch.cuda.stream(self.stream):
            self.ctx.init_synthetic(seed, *[self.f[k] for k in self.NAMES])

    def zero_sources(self):
        """The reference loop's rule for steps >= 1 (FluidSequential.c:298-302)."""
        with self.torch.cuda.stream(self.stream):
            for k in ("dens_prev", "u_prev", "v_prev"):
                self.f[k].zero_()

    def step(self, seed: Optional[int], visc: float, diff: float, dt: float):
        """One loop-body iteration (optional device-side source refresh, vel_step, dens_step)."""
        f = self.f
        with self.torch.cuda.stream(self.stream):
            if seed is not None:
                self.ctx.init_sources(seed, f["dens_prev"], f["u_prev"], f["v_prev"])
            self.ctx.step(f["dens"], f["dens_prev"], f["u"], f["u_prev"], f["v"], f["v_prev"], visc, diff, dt, self.iters)

    def new_host_fields(self):
        """Six pinned host arrays holding this rank's OWNED rows (own_rows, G), in NAMES order."""
        t = self.torch
        return [t.empty((self.own_rows, self.G), dtype=t.float32, pin_memory=True).zero_() for _ in self.NAMES]

    def step_host(self, host_fields, visc: float, diff: float, dt: float):
        """The loop body with HOST fields through the C ABI (``sf_step_host`` on a connected peer slab): uploads this rank's
        owned rows of the six fields, steps (collectively), downloads dens, u, v; returns when the host arrays are valid.
        One caller per slab at the same time: one process per GPU, or one thread per slab (ctypes releases the GIL).
        ``step_host_begin`` / ``step_host_end`` below are the same schedule split in two for a single thread that drives
        several slabs."""
        if self.world == 1:
            raise SF.StableFluidsError("PeerSlabSolver.step_host: world == 1, use StableFluids.step_host")
        self.ctx.step_host(*host_fields, visc, diff, dt, self.iters)

    def step_host_end(self):
        """Returns when the host arrays of the step begun last are valid."""
        self._d2h.synchronize()

    def step_host_begin(self, host_fields, visc: float, diff: float, dt: float):
        """The loop body for a caller whose fields live in HOST memory (the reference's CPU calling
        convention, FluidSequential.c:305-306), per slab: upload the owned rows of the six fields, step,
        download dens, u, v.  Enqueues only (several slabs of one process must all have begun before anyone
        waits); ``step_host_end`` waits.  Collective like ``step``.
        Ghost rows need no upload: every stage that reads them refreshes them from the neighbour first."""
        t, f = self.torch, self.f
        if not hasattr(self, "_h2d"):
            self._h2d, self._d2h = t.cuda.Stream(device=self.ctx.device), t.cuda.Stream(device=self.ctx.device)
        h = dict(zip(self.NAMES, host_fields))
        up = lambda name: self.owned(f[name]).copy_(h[name], non_blocking=True)
        down = lambda name: h[name].copy_(self.owned(f[name]), non_blocking=True)
        self._h2d.wait_stream(self.stream)           # the previous step is done with the device fields
        with t.cuda.stream(self._h2d):               # velocity first: vel_step starts while the density fields travel
            for name in ("u", "u_prev", "v", "v_prev"):
                up(name)
            vel_in = t.cuda.Event(); vel_in.record(self._h2d)
            for name in ("dens", "dens_prev"):
                up(name)
            dens_in = t.cuda.Event(); dens_in.record(self._h2d)
        with t.cuda.stream(self.stream):
            self.stream.wait_event(vel_in)
            self.ctx.vel_step(f["u"], f["v"], f["u_prev"], f["v_prev"], visc, dt, self.iters)
            vel_out = t.cuda.Event(); vel_out.record(self.stream)
            self.stream.wait_event(dens_in)
            self.ctx.dens_step(f["dens"], f["dens_prev"], f["u"], f["v"], diff, dt, self.iters)
            dens_out = t.cuda.Event(); dens_out.record(self.stream)
        with t.cuda.stream(self._d2h):               # u, v drain while dens_step runs
            self._d2h.wait_event(vel_out)
            down("u"); down("v")
            self._d2h.wait_event(dens_out)
            down("dens")

    def status(self) -> int:
        """Synchronise; raise if a neighbour barrier timed out or a back-trace left the neighbour's slab."""
        bits = self.ctx.slab_status() if self.world > 1 else 0
        if bits:
            what = []
            if bits & SF.SF_SLAB_ERR_TIMEOUT:
                what.append("a neighbour barrier timed out")
            if bits & SF.SF_SLAB_ERR_REACH:
                what.append("an advection back-trace reached beyond the neighbouring slab")
            raise SF.StableFluidsError("peer slab: " + "; ".join(what))
        return bits

    def check_reach(self):
        self.status()

    def close(self):
        self.f = {}
        self.ctx.close()
