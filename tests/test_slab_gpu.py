"""Slab decomposition on one GPU: p emulated ranks advanced in lock-step (halos copied directly)
must reproduce the CPU oracle -- and therefore the single-GPU path -- bit for bit."""
import numpy as np
import pytest
import torch

from gpu_util import bits_equal, mismatch_report

pytestmark = pytest.mark.gpu

DT, VIS, DIFF = 0.016, 0.0025, 0.1


def gather(solvers, name):
    return torch.cat([s.owned(s.f[name]) for s in solvers], dim=0).cpu().numpy()


@pytest.mark.parametrize("deferred", [False, True])
@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("N,K", [(254, 20), (126, 40)])
def test_slabs_bit_identical_to_oracle(oracle, world, N, K, deferred):
    from fluidsimulationcuda_b200.slab import SlabSolver, run_lockstep
    if (N + 2) // world < 16:
        pytest.skip("slab thinner than two boundary strips")
    solvers = [SlabSolver(N, r, world, iters=K, halo=16 if N < 200 else 24, overlap=False, deferred_reach=deferred)
               for r in range(world)]
    for s in solvers:
        s.init_synthetic(5)
    w = oracle.init_synthetic(N, 5)
    for k in w:
        assert bits_equal(gather(solvers, k), w[k]), f"IC {k}"
    for step in range(3):
        if step > 0:
            for s in solvers:
                for k in ("dens_prev", "u_prev", "v_prev"):
                    s.f[k].zero_()
        run_lockstep(solvers, VIS, DIFF, DT)
        oracle.run_steps(N, 1, w, VIS, DIFF, DT, K, first_step=step)
        torch.cuda.synchronize()
        for k in w:
            got = gather(solvers, k)
            assert bits_equal(got, w[k]), mismatch_report(got, w[k], f"world={world} step={step} {k}")
    for s in solvers:
        s.check_reach()


def test_single_slab_solver_equals_plain_solver(oracle):
    from fluidsimulationcuda_b200.slab import SlabSolver
    N, K = 126, 20
    s = SlabSolver(N, 0, 1, iters=K)
    s.init_synthetic(9)
    w = oracle.init_synthetic(N, 9)
    s.step(None, VIS, DIFF, DT)
    oracle.run_steps(N, 1, w, VIS, DIFF, DT, K)
    for k in w:
        assert bits_equal(s.f[k].cpu().numpy(), w[k]), k


def test_advection_reach_larger_than_halo_is_an_error():
    from fluidsimulationcuda_b200.slab import SlabSolver, run_lockstep
    from fluidsimulationcuda_b200.solver import StableFluidsError
    N = 254
    for deferred in (False, True):
        solvers = [SlabSolver(N, r, 2, iters=4, halo=8, overlap=False, deferred_reach=deferred) for r in range(2)]
        for s in solvers:
            s.init_synthetic(1)
            s.f["u_prev"].fill_(400.0)      # dt*N*v ~ 26 rows per step after add_source
            s.f["v_prev"].fill_(400.0)
        with pytest.raises(StableFluidsError):
            run_lockstep(solvers, VIS, DIFF, DT)
            for s in solvers:               # deferred mode reports at the check, not inside the step
                s.check_reach()
