"""Developer probe: lin_solve time at depths 5 and 7 per arithmetic mode; argv[1] = staging (0 cp.async, 1 bulk copy)."""
import sys; sys.path.insert(0, ".")
import torch, os
from fluidsimulationcuda_b200 import solver as SF
print("lib", os.environ.get("SF_LIBRARY", "default"))
G = 8192; K = 40
import sys
staging = int(sys.argv[1]) if len(sys.argv) > 1 else 0
print("staging", staging)
for T in (5, 7):
    for mode, (al, be) in (("pressure", (1.0, 4.0)), ("strict", (2683.2, 10733.8))):
        s = SF.StableFluids(G - 2, sweeps_per_launch=T, use_graph=False)
        s.set_option(SF.SF_OPT_STAGING, staging)
        x, x0 = s.new_field(), s.new_field(); x.uniform_(0, 1); x0.uniform_(0, 1)
        for _ in range(2): s.diffuse(0, x, x0, al, be, K)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(5):
            a.record(); s.diffuse(0, x, x0, al, be, K); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print(f"T={T} {mode:8s} {min(ts):8.3f} ms", flush=True)
        s.close()
