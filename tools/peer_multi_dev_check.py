"""One process, one peer slab per real GPU (connect_local across devices), bit-compared with the oracle.
usage: peer_multi_dev_check.py [world] [N] [K] [steps] [graph]"""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from fluidsimulationcuda_b200.slab import PeerSlabSolver
from oracle.pyoracle import Oracle
world = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N = int(sys.argv[2]) if len(sys.argv) > 2 else 510
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
graph = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
solvers = [PeerSlabSolver(N, r, world, iters=K, timeout_ms=60000, device=r, use_graph=graph) for r in range(world)]
import os
from fluidsimulationcuda_b200 import solver as SF
if os.environ.get("SF_STEAL") is not None:
    for s in solvers: s.ctx.set_option(SF.SF_OPT_WORK_STEALING, int(os.environ["SF_STEAL"]))
if os.environ.get("SF_ALL_PEERS"):
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    for i in range(world):
        rt.cudaSetDevice(i)
        for j in range(world):
            mode = os.environ["SF_ALL_PEERS"]      # all | nbr (|i-j| == 1 only) | far (|i-j| > 1 only)
            if i != j and (mode == "all" or (mode == "nbr" and abs(i - j) == 1) or (mode == "far" and abs(i - j) > 1)):
                print("enable", i, "->", j, rt.cudaDeviceEnablePeerAccess(j, 0))
    rt.cudaGetLastError(); rt.cudaSetDevice(0)
for s in solvers: s.connect_local(solvers)
for s in solvers: s.init_synthetic(5)
o = Oracle(threads=True); w = o.init_synthetic(N, 5)
for st in range(steps):
    if st > 0:
        for s in solvers: s.zero_sources()
    for s in solvers: s.step(None, 0.0025, 0.1, 0.016)
    for r, s in enumerate(solvers):
        try:
            s.ctx.synchronize(); print("rank", r, "stream ok", flush=True)
        except Exception as e:
            print("rank", r, "stream error:", e, flush=True)
    for d in range(world): torch.cuda.synchronize(d)
    print("step", st, "done", flush=True)
    o.run_steps(N, 1, w, 0.0025, 0.1, 0.016, K, first_step=st)
    for k in w:
        got = torch.cat([s.owned(s.f[k]).cpu() for s in solvers], dim=0).numpy()
        bad = got.view(np.uint32) != w[k].view(np.uint32)
        print(f"  {k}: {'identical' if not bad.any() else str(int(bad.sum())) + ' cells differ, rows ' + str(sorted(set(np.argwhere(bad)[:, 0].tolist()))[:12])}", flush=True)
for s in solvers: s.status(); s.close()
print("OK")
