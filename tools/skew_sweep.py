"""Full-step and per-solve time at G=8192, K=40 against SF_OPT_WAVE_SKEW (percent of a chunk between successive CTA waves)."""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from fluidsimulationcuda_b200 import solver as SF
G, K = 8192, 40
f32 = np.float32
al = f32(0.016) * f32(0.0025); al = al * f32(G - 2); al = al * f32(G - 2); be = f32(1) + f32(4) * al
for pct in [int(a) for a in (sys.argv[1:] or ["0", "120102", "125102", "131103", "135104", "140105"])]:
    s = SF.StableFluids(G - 2)
    s.set_option(SF.SF_OPT_WAVE_SKEW, pct)
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
    res = []
    for (bb, x, x0, A, B) in ((1, f[3], f[2], float(al), float(be)), (0, f[5], f[4], 1.0, 4.0)):
        ts = []
        for i in range(5):
            a.record(); s.diffuse(bb, x, x0, A, B, K); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        res.append(min(ts))
    print(f"wave skew {pct:6d}: step {step_ms:.3f} ms   strict solve {res[0]:.3f} ms   pressure solve {res[1]:.3f} ms", flush=True)
    s.close(); del f
