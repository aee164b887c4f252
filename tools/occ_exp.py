"""Developer probe: lin_solve time at depth 3 on a small and a large grid (occupancy experiments; honours SF_LIBRARY)."""
import sys; sys.path.insert(0, ".")
import torch
from fluidsimulationcuda_b200 import solver as SF
import os
print("lib", os.environ.get("SF_LIBRARY", "default"))
for G in (2048, 8192):
    for T in (3,):
        for mode, (al, be) in (("pressure", (1.0, 4.0)), ("strict", (2683.2, 10733.8))):
            s = SF.StableFluids(G - 2, sweeps_per_launch=T, use_graph=False)
            x, x0 = s.new_field(), s.new_field(); x.uniform_(0, 1); x0.uniform_(0, 1)
            K = 42
            for _ in range(2): s.diffuse(0, x, x0, al, be, K)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ts = []
            for _ in range(5):
                a.record(); s.diffuse(0, x, x0, al, be, K); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
            ms = min(ts)
            print(f"G={G} T={T} {mode:8s} {ms:8.3f} ms  {K*(G-2)**2/ms/1e6:8.1f} Gupd/s", flush=True)
            s.close()
