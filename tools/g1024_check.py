import sys; sys.path.insert(0, ".")
import torch
from fluidsimulationcuda_b200 import solver as SF
for fuse in (1, 0):
    N, K = 1022, 20
    s = SF.StableFluids(N); s.set_option(SF.SF_OPT_FUSE_SOURCES, fuse)
    f = [s.new_field() for _ in range(6)]
    s.init_synthetic(1, *f)
    s.run_steps(*f, 0.0025, 0.1, 0.016, K, 20, SF.SOURCES_SYNTHETIC, 10); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); s.run_steps(*f, 0.0025, 0.1, 0.016, K, 1000, SF.SOURCES_SYNTHETIC, 100); b.record(); torch.cuda.synchronize()
    print(f"G=1024 K=20 fuse_sources={fuse}: {a.elapsed_time(b) / 1000:.4f} ms/step", flush=True)
