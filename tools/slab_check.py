"""Multi-GPU slab run checked BITWISE against the CPU oracle (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/slab_check.py [G] [K] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from fluidsimulationcuda_b200.slab import SlabSolver, TorchDistComm, PeerSlabSolver
MODE = os.environ.get("SF_SLAB_COMM", "peer")     # peer = device-side peer-memory driver, nccl = NCCL send/recv driver
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
N = G - 2
if MODE == "peer":
    s = PeerSlabSolver(N, rank, world, iters=K)
    s.connect_dist()
else:
    s = SlabSolver(N, rank, world, iters=K, comm=TorchDistComm())
s.init_synthetic(3)
for st in range(steps):
    if st > 0:
        if MODE == "peer":
            s.zero_sources()          # on the slab's own stream
        else:
            for k in ("dens_prev", "u_prev", "v_prev"): s.f[k].zero_()
    s.step(None, 0.0025, 0.1, 0.016)
s.check_reach()
torch.cuda.synchronize()
ok = True
for k in s.names:
    mine = s.owned(s.f[k]).contiguous()
    parts = [torch.empty((hi - lo, G), device="cuda", dtype=torch.float32) for lo, hi in
             [(r * G // world, (r + 1) * G // world) for r in range(world)]] if rank == 0 else None
    dist.gather(mine, parts, dst=0)
    if rank == 0:
        got = torch.cat(parts, 0).cpu().numpy()
        if k == s.names[0]:
            from oracle.pyoracle import Oracle
            o = Oracle(threads=True); w = o.init_synthetic(N, 3); o.run_steps(N, steps, w, 0.0025, 0.1, 0.016, K)
        same = np.array_equal(got.view(np.uint32), w[k].view(np.uint32))
        ok &= same
        print(f"[{MODE}] world={world} G={G} K={K} steps={steps} field {k}: {'bit-identical' if same else 'MISMATCH'}", flush=True)
if rank == 0:
    print("SLAB CHECK", "PASSED" if ok else "FAILED", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
