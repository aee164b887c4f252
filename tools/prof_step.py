"""Profiling target for `ncu -k regex:<kernel>`: a few whole steps at G=8192, K=40 with graphs off (every kernel of the step is
its own launch, same kernels and launch plan as the graphed step), sources refreshed every step like bench.py."""
import sys; sys.path.insert(0, ".")
import torch
from fluidsimulationcuda_b200 import solver as SF
G, K = 8192, 40
s = SF.StableFluids(G - 2, use_graph=False)
f = [s.new_field() for _ in range(6)]
s.init_synthetic(1, *f)
for i in range(4):
    s.init_sources(10 + i, f[1], f[3], f[5]); s.step(*f, 0.0025, 0.1, 0.016, K)
torch.cuda.synchronize()
print("done", s.launch_count)
