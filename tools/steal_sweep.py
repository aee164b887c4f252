"""Full-step time at G=8192, K=40 against the work-stealing threshold (SF_OPT_WORK_STEALING, percent of a chunk)."""
import sys; sys.path.insert(0, ".")
import torch
from fluidsimulationcuda_b200 import solver as SF
G, K = 8192, 40
for pct in [int(a) for a in (sys.argv[1:] or ["30", "15", "10", "6", "0"])]:
    s = SF.StableFluids(G - 2)
    s.set_option(SF.SF_OPT_WORK_STEALING, pct)
    f = [s.new_field() for _ in range(6)]
    s.init_synthetic(1, *f)
    for i in range(6):
        s.init_sources(10 + i, f[1], f[3], f[5]); s.step(*f, 0.0025, 0.1, 0.016, K)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(30):
        s.init_sources(100 + i, f[1], f[3], f[5]); s.step(*f, 0.0025, 0.1, 0.016, K)
    b.record(); torch.cuda.synchronize()
    step_ms = a.elapsed_time(b) / 30
    ts = []
    for i in range(5):
        s.init_sources(200 + i, f[1], f[3], f[5])
        a.record(); s.dens_step(f[0], f[1], f[2], f[4], 0.1, 0.016, K); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"steal pct {pct:3d}: step {step_ms:.3f} ms   dens_step {min(ts):.3f} ms   ranges taken so far {s.get_option(SF.SF_OPT_STEAL_COUNT)}", flush=True)
    s.close(); del f
