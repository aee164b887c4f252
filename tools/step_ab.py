"""Step time against library options (developer tool): the bench's timed step (sources refreshed on the device, then sf_step
from its captured graph), CUDA events over `steps` steps.  usage: step_ab.py G K name=opt:val,opt:val ..."""
import sys
sys.path.insert(0, ".")
import torch
from fluidsimulationcuda_b200 import solver as SF
G = int(sys.argv[1]); K = int(sys.argv[2])
N = G - 2
VIS, DIFF, DT = 0.0025, 0.1, 0.016
def run(opts, steps=20, warm=5):
    s = SF.StableFluids(N)
    for o, v in opts: s.set_option(o, v)
    f = [s.new_field() for _ in range(6)]
    s.init_synthetic(1, *f)
    for k in range(warm):
        s.init_sources(2 + k, f[1], f[3], f[5]); s.step(*f, VIS, DIFF, DT, K)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(steps):
        s.init_sources(2 + warm + k, f[1], f[3], f[5]); s.step(*f, VIS, DIFF, DT, K)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    chk = [float(t.double().sum()) for t in (f[0], f[2], f[4])]
    s.close()
    return ms, chk
base = None
for spec in sys.argv[3:]:
    name, _, rest = spec.partition("=")
    opts = [(int(p.split(":")[0]), int(p.split(":")[1])) for p in rest.split(",") if p]
    ms, chk = run(opts)
    if base is None: base = chk
    print(f"{name:28s} {ms:8.4f} ms/step   same sums as first: {chk == base}", flush=True)
