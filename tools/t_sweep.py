"""lin_solve time vs the temporal-blocking depth cap (SF_OPT_SWEEPS_PER_LAUNCH) for a given K.
usage: t_sweep.py [G] [K] [T,T,...]   (default depths 6,7,8; BASELINE config 3's sweep: t_sweep.py 8192 40 1,2,3,4,5,6,7,8)
Prints per depth the solve time, the launches it took and the effective bandwidth 12 B x G^2 x K / time."""
import sys; sys.path.insert(0, ".")
import torch
from fluidsimulationcuda_b200 import solver as SF
G = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
DEPTHS = tuple(int(t) for t in sys.argv[3].split(",")) if len(sys.argv) > 3 else (6, 7, 8)
for T in DEPTHS:
    for mode, (al, be) in (("pressure", (1.0, 4.0)), ("strict", (2683.2, 10733.8))):
        s = SF.StableFluids(G - 2, sweeps_per_launch=T, use_graph=False)
        x, x0 = s.new_field(), s.new_field(); x.uniform_(0, 1); x0.uniform_(0, 1)
        s.diffuse(1, x, x0, al, be, K); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = s.launch_count
        ts = []
        for _ in range(3):
            a.record(); s.diffuse(1, x, x0, al, be, K); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print(f"G={G} K={K} T<={T} {mode:8s} {min(ts):9.3f} ms  ({(s.launch_count - n0) // 3} launches)  "
              f"effective {12.0 * G * G * K / (min(ts) * 1e-3) / 1e12:6.2f} TB/s", flush=True)
        s.close()
