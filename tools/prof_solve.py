"""Profiling target: a few temporally blocked lin_solve launches at G=8192 (developer tool).
usage: prof_solve.py <pressure|strict|fast> <T> [G] [b]      (b = 0: scalar field, work stealing on; b = 1: velocity-like)"""
import sys
sys.path.insert(0, ".")
import torch
from fluidsimulationcuda_b200 import solver as SF
mode, T = sys.argv[1], int(sys.argv[2])
G = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1
N = G - 2
s = SF.StableFluids(N, sweeps_per_launch=T, arithmetic=1 if mode == "fast" else 0, use_graph=False)
x, x0 = s.new_field(), s.new_field()
x.uniform_(0, 1); x0.uniform_(0, 1)
al, be = (1.0, 4.0) if mode == "pressure" else (2683.2, 10733.8)
assert mode != "strict" or s.division_check(be)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    a.record(); s.diffuse(B, x, x0, al, be, 2 * T); b.record(); torch.cuda.synchronize()
    print(mode, T, "2 launches", a.elapsed_time(b), "ms", flush=True)
