"""Developer probe: step time on small grids (G = 128 .. 4096), where launches, not bandwidth, bound the step."""
import sys; sys.path.insert(0, ".")
import torch
from fluidsimulationcuda_b200 import solver as SF
for G, K in ((128, 20), (256, 20), (512, 20), (1024, 20), (2048, 20), (4096, 40)):
    N = G - 2
    s = SF.StableFluids(N)
    f = [s.new_field() for _ in range(6)]
    s.init_synthetic(1, *f)
    for i in range(10):
        s.init_sources(10 + i, f[1], f[3], f[5]); s.step(*f, 0.0025, 0.1, 0.016, K)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 200
    a.record()
    for i in range(steps):
        s.init_sources(100 + i, f[1], f[3], f[5]); s.step(*f, 0.0025, 0.1, 0.016, K)
    b.record(); torch.cuda.synchronize()
    print(f"G={G} K={K}: {a.elapsed_time(b)/steps:.4f} ms/step", flush=True)
    s.close()
