"""Developer A/B on real GPUs (torchrun, one rank per GPU): SF_OPT_STRIP_BALANCE on / off on connected peer slabs.
usage: python -m torch.distributed.run --nproc-per-node P tools/strip_balance_ab.py G K[,K2...] [steps]
Prints ms per step (CUDA events on the slab's stream, max over ranks) for balance = 1, 0, 1, 0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

DT, VIS, DIFF = 0.016, 0.0025, 0.1


def main():
    G = int(sys.argv[1]); Ks = [int(k) for k in sys.argv[2].split(",")]
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from fluidsimulationcuda_b200 import build, solver as SF
    from fluidsimulationcuda_b200.slab import PeerSlabSolver
    build.build()
    for K in Ks:
        sim = PeerSlabSolver(G - 2, rank, world, iters=K, arithmetic=SF.STRICT)
        sim.connect_dist()
        sim.init_synthetic(1)
        sync = lambda: (torch.cuda.synchronize(), dist.barrier())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep, balance in enumerate((1, 0, 1, 0)):
            sim.ctx.set_option(SF.SF_OPT_STRIP_BALANCE, balance)
            for i in range(3):
                sim.step(100 + i, VIS, DIFF, DT)
            sync()
            a.record(sim.stream)
            for i in range(steps):
                sim.step(1000 + i, VIS, DIFF, DT)
            b.record(sim.stream)
            sync()
            t = torch.tensor([a.elapsed_time(b) / steps], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sim.status()
            if rank == 0:
                print(f"G={G} K={K} world={world} strip_balance={balance}: {float(t.item()):.3f} ms per step", flush=True)
        sim.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
