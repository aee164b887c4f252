"""Per-phase device times of the step on 1 GPU (full grid) or on peer slabs (torchrun, one rank per GPU):
the stage entry points are called one by one (no graph) with CUDA events around each.
  python tools/slab_phase_times.py [G] [K]      |     torchrun ... tools/slab_phase_times.py [G] [K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fluidsimulationcuda_b200 import solver as SF
from fluidsimulationcuda_b200.slab import PeerSlabSolver, f32_coeffs
G = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
N = G - 2
DT, VIS, DIFF = 0.016, 0.0025, 0.1
s = PeerSlabSolver(N, rank, world, iters=K, use_graph=False)
if world > 1:
    s.connect_dist()
if os.environ.get("SF_STEAL"):
    s.ctx.set_option(SF.SF_OPT_WORK_STEALING, int(os.environ["SF_STEAL"]))
s.init_synthetic(1)
for i in range(3):
    s.step(100 + i, VIS, DIFF, DT)
c, f, st = s.ctx, s.f, s.stream
av, bv = f32_coeffs(DT, VIS, N)
ad, bd = f32_coeffs(DT, DIFF, N)
phases = [
    ("init_sources", lambda: c.init_sources(7, f["dens_prev"], f["u_prev"], f["v_prev"])),
    ("add_source u", lambda: c.add_source(f["u"], f["u_prev"], DT)),
    ("diffuse u (strict, visc)", lambda: c.diffuse(1, f["u_prev"], f["u"], av, bv, K)),
    ("diffuse v (strict, visc)", lambda: c.diffuse(2, f["v_prev"], f["v"], av, bv, K)),
    ("project #1", lambda: c.project(f["u_prev"], f["v_prev"], f["u"], f["v"], K)),
    ("advect u", lambda: c.advect(1, f["u"], f["u_prev"], f["u_prev"], f["v_prev"], DT)),
    ("advect v", lambda: c.advect(2, f["v"], f["v_prev"], f["u_prev"], f["v_prev"], DT)),
    ("project #2", lambda: c.project(f["u"], f["v"], f["u_prev"], f["v_prev"], K)),
    ("add_source dens", lambda: c.add_source(f["dens"], f["dens_prev"], DT)),
    ("diffuse dens (strict, diff)", lambda: c.diffuse(0, f["dens_prev"], f["dens"], ad, bd, K)),
    ("advect dens", lambda: c.advect(0, f["dens"], f["dens_prev"], f["u"], f["v"], DT)),
]
tot = {n: 0.0 for n, _ in phases}
REPS = 3
with torch.cuda.stream(st):
    for rep in range(REPS + 1):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(phases) + 1)]
        evs[0].record(st)
        for i, (n, fn) in enumerate(phases):
            fn(); evs[i + 1].record(st)
        st.synchronize()
        if rep:
            for i, (n, _) in enumerate(phases):
                tot[n] += evs[i].elapsed_time(evs[i + 1]) / REPS
if world > 1:
    dist.barrier()
for r in range(world):
    if r == rank:
        print(f"--- rank {rank}/{world} G={G} K={K}: per-phase ms (eager launches)")
        for n, _ in phases:
            print(f"  {n:30s} {tot[n]:9.3f}")
        print(f"  {'sum':30s} {sum(tot.values()):9.3f}", flush=True)
    if world > 1:
        dist.barrier()
s.status()
if world > 1:
    dist.destroy_process_group()
