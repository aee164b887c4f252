"""A/B timing of one 40-sweep lin_solve per (T, mode) and of the full step: default library vs SF_LIBRARY."""
import os, sys; sys.path.insert(0, ".")
import torch
from fluidsimulationcuda_b200 import solver as SF
print("lib", os.environ.get("SF_LIBRARY", "default"), flush=True)
G = 8192; K = 40
for T in [int(t) for t in os.environ.get("SF_AB_T", "6,7").split(",")]:
    for mode, (al, be) in (("pressure", (1.0, 4.0)), ("strict", (2683.2, 10733.8))):
        s = SF.StableFluids(G - 2, sweeps_per_launch=T, use_graph=False)
        x, x0 = s.new_field(), s.new_field(); x.uniform_(0, 1); x0.uniform_(0, 1)
        for _ in range(2): s.diffuse(0, x, x0, al, be, K)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(7):
            a.record(); s.diffuse(0, x, x0, al, be, K); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print(f"T={T} {mode:8s} min {min(ts):8.3f} ms  median {sorted(ts)[3]:8.3f}", flush=True)
        s.close()
s = SF.StableFluids(G - 2)
if os.environ.get("SF_STEAL"):
    s.set_option(SF.SF_OPT_WORK_STEALING, int(os.environ["SF_STEAL"]))
if os.environ.get("SF_PRESSURE_PLAN"):
    s.set_option(SF.SF_OPT_PRESSURE_PLAN, int(os.environ["SF_PRESSURE_PLAN"]))
if os.environ.get("SF_STEAL_SCOPE"):
    s.set_option(SF.SF_OPT_STEAL_SCOPE, int(os.environ["SF_STEAL_SCOPE"]))
f = [s.new_field() for _ in range(6)]
s.init_synthetic(1, *f)
for i in range(4):
    s.init_sources(10 + i, f[1], f[3], f[5]); s.step(*f, 0.0025, 0.1, 0.016, K)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(30):
    s.init_sources(100 + i, f[1], f[3], f[5]); s.step(*f, 0.0025, 0.1, 0.016, K)
b.record(); torch.cuda.synchronize()
print(f"full step: {a.elapsed_time(b)/30:.3f} ms", flush=True)
