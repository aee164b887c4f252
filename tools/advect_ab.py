"""advect A/B on the velocity / density fields a timed step really sees (developer tool):
SF_OPT_ADVECT_TILE = 0 (global gathers) against 1..8 (TMA-staged source tile), device time per launch, effective HBM
bandwidth at the 16 B/cell a launch has to move, and the share of tiles that fell back to gathers."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from fluidsimulationcuda_b200 import solver as SF
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tiles = [int(t) for t in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 6, 5, 4]
N = G - 2
DT, VIS, DIFF = 0.016, 0.0025, 0.1
s = SF.StableFluids(N, use_graph=False)
dens, dens0, u, u0, v, v0 = [s.new_field() for _ in range(6)]
s.init_synthetic(1, dens, dens0, u, u0, v, v0)
for step in range(2):     # two whole steps, then the third up to the advection
    s.init_sources(2 + step, dens0, u0, v0)
    s.step(dens, dens0, u, u0, v, v0, VIS, DIFF, DT, K)
s.init_sources(9, dens0, u0, v0)
f32 = np.float32
a = f32(DT) * f32(VIS); a = a * f32(N); a = a * f32(N); al, be = float(a), float(f32(1) + f32(4) * a)
s.add_source(u, u0, DT); s.add_source(v, v0, DT)
s.diffuse(1, u0, u, al, be, K); s.diffuse(2, v0, v, al, be, K)
s.project(u0, v0, u, v, K)
torch.cuda.synchronize()
print("max|u0|", s.reduce_max_abs(u0), "max|v0|", s.reduce_max_abs(v0), "trace cells ~", DT * N * s.reduce_max_abs(u0), flush=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]
ref_u = ref_v = ref_d = None
for t in tiles:
    s.set_option(SF.SF_OPT_ADVECT_TILE, t)
    s.set_option(SF.SF_OPT_ADVECT_TILE_COUNT, 0)
    s.advect_velocity(u, v, u0, v0, DT); s.advect(0, dens0, dens, u, v, DT)
    tma, fb = s.get_option(SF.SF_OPT_ADVECT_TILE_COUNT), s.get_option(SF.SF_OPT_ADVECT_FALLBACK_COUNT)
    if ref_u is None: ref_u, ref_v, ref_d = u.clone(), v.clone(), dens0.clone()
    same = bool(torch.equal(u.view(torch.int32), ref_u.view(torch.int32)) and torch.equal(v.view(torch.int32), ref_v.view(torch.int32))
                and torch.equal(dens0.view(torch.int32), ref_d.view(torch.int32)))
    t2 = timed(lambda: s.advect_velocity(u, v, u0, v0, DT))
    t1 = timed(lambda: s.advect(0, dens0, dens, u, v, DT))
    bw = lambda ms: 16.0 * G * G / ms / 1e9
    print(f"tile={t}: velocity pair {t2:.4f} ms ({bw(t2):.2f} TB/s)  scalar {t1:.4f} ms ({bw(t1):.2f} TB/s)  tiles tma={tma} fallback={fb}  same bits as first: {same}", flush=True)
