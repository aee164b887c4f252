"""How many row ranges change hands per lin_solve with work stealing on, and what it costs / gains."""
import sys, os; sys.path.insert(0, ".")
import torch
from fluidsimulationcuda_b200 import solver as SF
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = 40
for steal in [0] + [int(a) for a in sys.argv[2:]]:
    s = SF.StableFluids(G - 2, use_graph=False)
    s.set_option(SF.SF_OPT_WORK_STEALING, steal)
    f = [s.new_field() for _ in range(6)]
    s.init_synthetic(1, *f)
    for i in range(3):
        s.init_sources(10 + i, f[1], f[3], f[5]); s.step(*f, 0.0025, 0.1, 0.016, K)
    for name, (b, x, x0, al, be) in (("u-like, b=0 (uniform)", (0, f[3], f[2], 2683.2, 10733.8)), ("density", (0, f[1], f[0], 107322.0, 429289.0))):
        n0 = s.get_option(SF.SF_OPT_STEAL_COUNT)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(5):
            a.record(); s.diffuse(b, x, x0, al, be, K); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
        n1 = s.get_option(SF.SF_OPT_STEAL_COUNT)
        print(f"G={G} steal={steal} {name:24s} {min(ts):8.3f} ms   ranges taken per solve: {(n1 - n0) / 5:.0f}", flush=True)
    s.close()
