"""Per-stage device times of one step, stage functions called one by one (developer tool)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from fluidsimulationcuda_b200 import solver as SF
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
arith = int(sys.argv[3]) if len(sys.argv) > 3 else 0
T = int(sys.argv[4]) if len(sys.argv) > 4 else 0
N = G - 2
DT, VIS, DIFF = 0.016, 0.0025, 0.1
s = SF.StableFluids(N, arithmetic=arith, use_graph=False, sweeps_per_launch=T)
dens, dens0, u, u0, v, v0 = [s.new_field() for _ in range(6)]
f32 = np.float32
def ab(c):
    a = f32(DT) * f32(c); a = a * f32(N); a = a * f32(N); return float(a), float(f32(1) + f32(4) * a)
times = {}
def timed(name, fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    times.setdefault(name, []).append(a.elapsed_time(b))
for step in range(4):
    s.init_sources(1 + step, dens0, u0, v0)
    al, be = ab(VIS)
    timed("add_source u", lambda: s.add_source(u, u0, DT)); timed("add_source v", lambda: s.add_source(v, v0, DT))
    timed("diffuse u", lambda: s.diffuse(1, u0, u, al, be, K)); timed("diffuse v", lambda: s.diffuse(2, v0, v, al, be, K))
    timed("divergence 1", lambda: s.computeDivergenceAndPressure(u0, v0, u, v))
    timed("pressure 1", lambda: s.diffuse(0, u, v, 1.0, 4.0, K))
    timed("lastProject 1", lambda: s.lastProject(u0, v0, u, v))
    timed("advect u", lambda: s.advect(1, u, u0, u0, v0, DT)); timed("advect v", lambda: s.advect(2, v, v0, u0, v0, DT))
    timed("divergence 2", lambda: s.computeDivergenceAndPressure(u, v, u0, v0))
    timed("pressure 2", lambda: s.diffuse(0, u0, v0, 1.0, 4.0, K))
    timed("lastProject 2", lambda: s.lastProject(u, v, u0, v0))
    timed("add_source d", lambda: s.add_source(dens, dens0, DT))
    al, be = ab(DIFF)
    timed("diffuse d", lambda: s.diffuse(0, dens0, dens, al, be, K))
    timed("advect d", lambda: s.advect(0, dens, dens0, u, v, DT))
    print("step", step, "max|u|", s.reduce_max_abs(u), "max|dens|", s.reduce_max_abs(dens), flush=True)
tot = 0
for k, v_ in times.items():
    print(f"{k:16s} " + " ".join(f"{t:8.3f}" for t in v_)); tot += v_[-1]
print("sum last step", tot)
