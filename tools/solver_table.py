"""Opt-in solvers against the reference's Jacobi at the headline size (G=8192, K=40): ms per lin_solve and the residual
sum of squares ||x0 - (beta*x - alpha*nbrs)||^2 after K iterations, for the viscosity system (alpha = 2683) and the
pressure system (alpha = 1, beta = 4).  Same right-hand side and initial guess for every solver.
usage: solver_table.py [G] [K] [out.json]"""
import json, sys
sys.path.insert(0, ".")
import numpy as np, torch
from fluidsimulationcuda_b200 import solver as SF
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
N = G - 2
f32 = np.float32
al = f32(0.016) * f32(0.0025); al = al * f32(N); al = al * f32(N); be = f32(1) + f32(4) * al
systems = {"viscosity (alpha=%.1f)" % float(al): (1, float(al), float(be)), "pressure (alpha=1, beta=4)": (0, 1.0, 4.0)}
solvers = [("Jacobi (the reference's scheme, temporally blocked)", dict()),
           ("red-black Gauss-Seidel, launch per half-sweep", dict(solver=1)),
           ("red-black Gauss-Seidel, blocked (3 iterations per launch)", dict(solver=1, blocked=1)),
           ("red-black SOR omega=1.7, blocked", dict(solver=1, blocked=1, omega=1700))]
torch.manual_seed(1)
rhs = (torch.rand((G, G), device="cuda") - 0.5) * 0.02
guess = (torch.rand((G, G), device="cuda") - 0.5) * 0.02
rows = []
for sysname, (b, A, B) in systems.items():
    for name, opt in solvers:
        s = SF.StableFluids(N, use_graph=False)
        if opt.get("solver"):
            s.set_option(SF.SF_OPT_SOLVER, SF.SOLVER_RBGS)
            s.set_option(SF.SF_OPT_RBGS_BLOCKED, opt.get("blocked", 0))
            s.set_option(SF.SF_OPT_SOR_OMEGA_MILLI, opt.get("omega", 1000))
        x = guess.clone()
        r0 = s.residual_sumsq(x, rhs, A, B)
        s.diffuse(b, x, rhs, A, B, K); torch.cuda.synchronize()          # warm (loads kernels, validates the divisor)
        ts = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(4):
            x.copy_(guess)
            e0.record(); s.diffuse(b, x, rhs, A, B, K); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        r1 = s.residual_sumsq(x, rhs, A, B)
        rows.append({"system": sysname, "solver": name, "iterations": K, "ms_per_solve": min(ts), "residual_sumsq_before": r0,
                     "residual_sumsq_after": r1, "reduction": r0 / r1 if r1 > 0 else None})
        print(f"{sysname:28s} {name:58s} {min(ts):8.3f} ms   residual^2 {r0:.4e} -> {r1:.4e}  (x{r0 / r1:.1f})", flush=True)
        s.close()
if len(sys.argv) > 3:
    json.dump({"G": G, "K": K, "gpu": torch.cuda.get_device_name(0), "rows": rows}, open(sys.argv[3], "w"), indent=1)
