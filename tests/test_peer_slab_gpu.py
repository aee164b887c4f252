"""Peer-memory slabs (csrc/sf_slab.cu): p slabs in ONE process on one GPU, each on its own stream,
wired to each other's arenas.  They synchronise on the device exactly as p GPUs would (neighbour
barrier kernels, fused strip pushes, peer-memory advect) and must reproduce the CPU oracle -- and
therefore the single-GPU path -- bit for bit, directly launched and replayed from CUDA graphs."""
import numpy as np
import pytest
import torch

from gpu_util import bits_equal, mismatch_report

pytestmark = pytest.mark.gpu

DT, VIS, DIFF = 0.016, 0.0025, 0.1


def make(N, world, K, devices=None, **kw):
    from fluidsimulationcuda_b200.slab import PeerSlabSolver
    solvers = [PeerSlabSolver(N, r, world, iters=K, timeout_ms=4000, device=None if devices is None else devices[r], **kw)
               for r in range(world)]
    for s in solvers:
        s.connect_local(solvers)
    torch.cuda.synchronize()
    return solvers


def gather(solvers, name):
    torch.cuda.synchronize()
    return torch.cat([s.owned(s.f[name]) for s in solvers], dim=0).cpu().numpy()


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("N,K", [(254, 20), (126, 40), (510, 7)])
def test_peer_slabs_bit_identical_to_oracle(oracle, world, N, K, use_graph):
    solvers = make(N, world, K, use_graph=use_graph)
    for s in solvers:
        s.init_synthetic(5)
    w = oracle.init_synthetic(N, 5)
    for k in w:
        assert bits_equal(gather(solvers, k), w[k]), f"IC {k}"
    for step in range(4):           # with graphs: direct, capture + launch, replay, replay
        if step > 0:
            for s in solvers:
                s.zero_sources()
        for s in solvers:
            s.step(None, VIS, DIFF, DT)
        oracle.run_steps(N, 1, w, VIS, DIFF, DT, K, first_step=step)
        for k in w:
            got = gather(solvers, k)
            assert bits_equal(got, w[k]), mismatch_report(got, w[k], f"world={world} step={step} {k}")
    for s in solvers:
        s.status()
        s.close()


def test_peer_slab_stage_entry_points(oracle):
    """sf_diffuse / sf_project / sf_advect on connected slabs are collective and bit-identical."""
    N, K, world = 254, 12, 2
    solvers = make(N, world, K, use_graph=False)
    rng = np.random.default_rng(3)
    G = N + 2
    full = {k: (rng.random((G, G), dtype=np.float32) - np.float32(0.5)) * np.float32(0.02) for k in ("dens", "dens_prev", "u", "v")}
    for s in solvers:
        with torch.cuda.stream(s.stream):
            for k, a in full.items():
                s.f[k].zero_()
                s.owned(s.f[k]).copy_(torch.from_numpy(a[s.row_lo:s.row_hi]).cuda())
    torch.cuda.synchronize()
    al, be = 2.5, 11.0
    want = {k: a.copy() for k, a in full.items()}
    oracle.diffuse(N, 1, want["dens"], want["dens_prev"], al, be, K)
    oracle.advect(N, 0, want["dens_prev"], want["dens"], want["u"], want["v"], DT)
    for s in solvers:
        with torch.cuda.stream(s.stream):
            s.ctx.diffuse(1, s.f["dens"], s.f["dens_prev"], al, be, K)
    for s in solvers:
        with torch.cuda.stream(s.stream):
            s.ctx.advect(0, s.f["dens_prev"], s.f["dens"], s.f["u"], s.f["v"], DT)
    for k in ("dens", "dens_prev"):
        got = gather(solvers, k)
        assert bits_equal(got, want[k]), mismatch_report(got, want[k], k)
    for s in solvers:
        s.status()
        s.close()


def test_peer_slab_long_reach_advect(oracle):
    """Back-traces that leave the slab by many rows are served from the neighbour's memory."""
    N, K, world = 254, 4, 2
    solvers = make(N, world, K, use_graph=False)
    for s in solvers:
        s.init_synthetic(1)
        with torch.cuda.stream(s.stream):
            s.f["u_prev"].mul_(40.0)      # dt*N*v of tens of rows after add_source
            s.f["v_prev"].mul_(40.0)
    w = oracle.init_synthetic(N, 1)
    w["u_prev"] *= np.float32(40.0); w["v_prev"] *= np.float32(40.0)
    for s in solvers:
        s.step(None, VIS, DIFF, DT)
    oracle.run_steps(N, 1, w, VIS, DIFF, DT, K)
    for k in w:
        got = gather(solvers, k)
        assert bits_equal(got, w[k]), mismatch_report(got, w[k], k)
    for s in solvers:
        s.status()
        s.close()


def test_peer_slab_host_field_step(oracle):
    """step_host: every slab uploads its rows of the six pinned host fields, steps, downloads dens/u/v."""
    N, K, world = 254, 12, 2
    solvers = make(N, world, K)
    w = oracle.init_synthetic(N, 8)
    hosts = []
    for s in solvers:
        hf = s.new_host_fields()
        for h, name in zip(hf, s.names):
            h.copy_(torch.from_numpy(w[name][s.row_lo:s.row_hi]))
        hosts.append(hf)
    for step in range(3):
        for s, hf in zip(solvers, hosts):
            s.step_host_begin(hf, VIS, DIFF, DT)
        for s in solvers:
            s.step_host_end()
        oracle.run_steps(N, 1, w, VIS, DIFF, DT, K, first_step=0)
        for i, name in enumerate(solvers[0].names):
            if name in ("dens", "u", "v"):
                got = torch.cat([hf[i] for hf in hosts], dim=0).numpy()
                assert bits_equal(got, w[name]), mismatch_report(got, w[name], f"step {step} {name}")
        # the *_prev host arrays were not downloaded: they still hold the sources; mirror that in the oracle
        for i, name in enumerate(solvers[0].names):
            if name.endswith("_prev"):
                w[name][...] = torch.cat([hf[i] for hf in hosts], dim=0).numpy()
    for s in solvers:
        s.status()
        s.close()


def test_peer_slab_step_host_through_the_c_abi(oracle):
    """sf_step_host on connected slabs (collective): one host thread per slab calls the C entry point with its owned rows."""
    import threading
    N, K, world = 254, 12, 3
    solvers = make(N, world, K)
    w = oracle.init_synthetic(N, 9)
    hosts = []
    for s in solvers:
        hf = s.new_host_fields()
        for h, name in zip(hf, s.names):
            h.copy_(torch.from_numpy(w[name][s.row_lo:s.row_hi]))
        hosts.append(hf)
    errors = []

    def run(s, hf):
        try:
            s.step_host(hf, VIS, DIFF, DT)
        except Exception as e:      # noqa: BLE001 -- reported by the main thread
            errors.append(e)
    for step in range(3):
        threads = [threading.Thread(target=run, args=(s, hf)) for s, hf in zip(solvers, hosts)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors
        oracle.run_steps(N, 1, w, VIS, DIFF, DT, K, first_step=0)
        for i, name in enumerate(solvers[0].names):
            if name in ("dens", "u", "v"):
                got = torch.cat([hf[i] for hf in hosts], dim=0).numpy()
                assert bits_equal(got, w[name]), mismatch_report(got, w[name], f"C ABI step_host, step {step} {name}")
        for i, name in enumerate(solvers[0].names):
            if name.endswith("_prev"):
                w[name][...] = torch.cat([hf[i] for hf in hosts], dim=0).numpy()
    for s in solvers:
        s.status()
        s.close()


def test_missing_neighbour_times_out_instead_of_hanging():
    from fluidsimulationcuda_b200.slab import PeerSlabSolver
    from fluidsimulationcuda_b200.solver import StableFluidsError
    N = 126
    solvers = [PeerSlabSolver(N, r, 2, iters=4, timeout_ms=200, use_graph=False) for r in range(2)]
    for s in solvers:
        s.connect_local(solvers)
    solvers[0].init_synthetic(1)
    solvers[0].step(None, VIS, DIFF, DT)          # rank 1 never steps
    with pytest.raises(StableFluidsError, match="timed out"):
        solvers[0].status()
    for s in solvers:
        s.close()


# ---- REAL devices: one slab per GPU in one process (peer access over NVLink), skipped on 1-GPU boxes ----------------
# What the emulated slabs above cannot exercise: st.release.sys / ld.acquire.sys ordering between two memory systems,
# peer stores of the strip rows and peer loads of the advect gathers over NVLink.  (The CUDA-IPC mapping between
# PROCESSES is exercised by bench.py at N > 1, which bit-checks a small problem against the oracle before it times.)
def _devices(world):
    n = torch.cuda.device_count()
    if n < world:
        pytest.skip(f"needs {world} GPUs, this box has {n}")
    return list(range(world))


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("N,K", [(510, 20), (1022, 40)])
def test_peer_slabs_on_real_devices_bit_identical_to_oracle(oracle_mt, world, N, K, use_graph):
    devices = _devices(world)
    solvers = make(N, world, K, devices=devices, use_graph=use_graph)
    for s in solvers:
        s.init_synthetic(5)
    w = oracle_mt.init_synthetic(N, 5)
    for step in range(3):
        if step > 0:
            for s in solvers:
                s.zero_sources()
        for s in solvers:
            s.step(None, VIS, DIFF, DT)
        oracle_mt.run_steps(N, 1, w, VIS, DIFF, DT, K, first_step=step)
        for k in w:
            for d in devices:
                torch.cuda.synchronize(d)
            got = torch.cat([s.owned(s.f[k]).cpu() for s in solvers], dim=0).numpy()
            assert bits_equal(got, w[k]), mismatch_report(got, w[k], f"real devices world={world} step={step} {k}")
    for s in solvers:
        s.status()
        s.close()


@pytest.mark.parametrize("world,tile", [(2, 15), (3, 8), (2, 4)])
def test_peer_slabs_with_tma_staged_advect(oracle_mt, world, tile):
    """SF_OPT_ADVECT_TILE on request on connected slabs (advect_tile_kernel<NF, PEER = true>): tiles whose traces stay inside
    the slab's own rows take the TMA box, tiles near a slab edge gather from the neighbour's memory; same bits as the oracle"""
    from fluidsimulationcuda_b200 import solver as SF
    N, K = 638, 6
    solvers = make(N, world, K, use_graph=True)
    for s in solvers:
        s.ctx.set_option(SF.SF_OPT_ADVECT_TILE, tile)
        s.init_synthetic(5)
        with torch.cuda.stream(s.stream):
            s.f["u_prev"].mul_(25.0)      # traces of several rows: some leave the slab
            s.f["v_prev"].mul_(25.0)
    w = oracle_mt.init_synthetic(N, 5)
    w["u_prev"] *= np.float32(25.0); w["v_prev"] *= np.float32(25.0)
    for step in range(3):           # direct, capture + launch, replay
        if step > 0:
            for s in solvers:
                s.zero_sources()
        for s in solvers:
            s.step(None, VIS, DIFF, DT)
        oracle_mt.run_steps(N, 1, w, VIS, DIFF, DT, K, first_step=step)
        for k in w:
            got = gather(solvers, k)
            assert bits_equal(got, w[k]), mismatch_report(got, w[k], f"world={world} tile={tile} step={step} {k}")
    for s in solvers:
        s.status()
        tma = s.ctx.get_option(SF.SF_OPT_ADVECT_TILE_COUNT); fb = s.ctx.get_option(SF.SF_OPT_ADVECT_FALLBACK_COUNT)
        assert tma > 0 and fb > 0, (tma, fb)      # interior tiles by TMA, tiles near the slab edges by peer gathers
        s.close()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("N,K", [(254, 20), (510, 40), (382, 12)])      # first launches of depth 5, 7, 6
def test_peer_slabs_fused_add_source_matches_separate_pass(oracle_mt, world, N, K):
    """SF_OPT_FUSE_SOURCES on connected slabs (jacobi_stream_kernel<T, MODE, 8 / 9>): the first launch of the viscosity and
    diffusion solves forms x + dt * s itself, its strip warps also write the right-hand side's ghost rows.  Same bits as the
    separate add_source pass and as the oracle, two launches fewer per slab and step."""
    from fluidsimulationcuda_b200 import solver as SF
    launches = {}
    for fuse in (1, 0):
        solvers = make(N, world, K, use_graph=True)
        for s in solvers:
            s.ctx.set_option(SF.SF_OPT_FUSE_SOURCES, fuse)
            s.init_synthetic(9)
        w = oracle_mt.init_synthetic(N, 9)
        n0 = [s.launch_count for s in solvers]
        for step in range(3):           # direct, capture + launch, replay
            if step == 2:
                for s in solvers:
                    s.zero_sources()
            for s in solvers:
                s.step(None, VIS, DIFF, DT)
            if step == 2:
                for k in ("dens_prev", "u_prev", "v_prev"):
                    w[k][...] = 0
            oracle_mt.run_steps(N, 1, w, VIS, DIFF, DT, K, first_step=0)
            for k in w:
                got = gather(solvers, k)
                assert bits_equal(got, w[k]), mismatch_report(got, w[k], f"world={world} fuse={fuse} step={step} {k}")
            if step == 0:
                launches[fuse] = [s.launch_count - a for s, a in zip(solvers, n0)]
        for s in solvers:
            s.status()
            s.close()
    assert all(a == b - 2 for a, b in zip(launches[1], launches[0])), launches


@pytest.mark.parametrize("world", [2, 3])
def test_peer_slabs_short_chunks_behind_the_strips(oracle_mt, world):
    """SF_OPT_STRIP_BALANCE: the warps that computed a boundary strip take a shorter interior chunk.  With a mean chunk of 60
    rows on slabs of 170-256 rows (K = 20: launches of 5 sweeps, strips of 5 rows, cost 30 rows) the short chunks are 26-44 rows
    (one per neighbour), the others 56-74: same bits as equal chunks and as the oracle."""
    from fluidsimulationcuda_b200 import solver as SF
    N, K = 510, 20
    for balance in (1, 0):
        solvers = make(N, world, K, use_graph=True)
        for s in solvers:
            s.ctx.set_option(SF.SF_OPT_CHUNK_ROWS, 60)
            s.ctx.set_option(SF.SF_OPT_STRIP_BALANCE, balance)
            assert s.ctx.get_option(SF.SF_OPT_STRIP_BALANCE) == balance
            s.init_synthetic(11)
        w = oracle_mt.init_synthetic(N, 11)
        for step in range(3):           # direct, capture + launch, replay
            for s in solvers:
                s.step(None, VIS, DIFF, DT)
            oracle_mt.run_steps(N, 1, w, VIS, DIFF, DT, K, first_step=0)
            for k in w:
                got = gather(solvers, k)
                assert bits_equal(got, w[k]), mismatch_report(got, w[k], f"world={world} balance={balance} step={step} {k}")
        for s in solvers:
            s.status()
            s.close()
