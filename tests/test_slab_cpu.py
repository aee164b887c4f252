"""Host-side logic of the slab decomposition on CPU: partitioning, launch planning, and the
neighbour exchange under torch.distributed (gloo, world_size 2 and 3)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fluidsimulationcuda_b200.slab import (HaloSpec, SlabLayout, TorchDistComm, exchange_handles, f32_coeffs,
                                            neighbour_links, partition_rows, plan_launches)
from fluidsimulationcuda_b200 import solver as SF


def test_partition_covers_grid():
    for G in (16, 130, 1024, 8192, 32768):
        for p in (1, 2, 3, 4, 8):
            parts = partition_rows(G, p)
            assert parts[0][0] == 0 and parts[-1][1] == G
            for (a, b), (c, d) in zip(parts, parts[1:]):
                assert b == c and b > a
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_launch_plan():
    assert plan_launches(40, 8) == [7, 7, 7, 7, 6, 6]
    assert plan_launches(20, 8) == [5, 5, 5, 5]
    assert plan_launches(200, 8) == [8] * 18 + [7] * 8 or sum(plan_launches(200, 8)) == 200
    for iters in range(1, 90):
        for T in range(1, 9):
            plan = plan_launches(iters, T)
            assert sum(plan) == iters and max(plan) <= T and min(plan) >= 1
            assert len(plan) % 2 == 0 or iters == 1 or len(plan) == iters, (iters, T, plan)


def test_coefficients_match_reference_rounding():
    # FluidSequential.c:179-180 evaluated in binary32, left to right
    f = np.float32
    a = f(f(f(0.016) * f(0.1)) * f(8190)) * f(8190)
    assert f32_coeffs(0.016, 0.1, 8190) == (float(a), float(f(1) + f(4) * a))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, G, halo, h):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lay = SlabLayout(G, rank, world, halo)
        rows = lay.own_rows + 2 * halo
        # every cell holds its GLOBAL row number (+ a per-field offset); ghosts start as -1
        fields = []
        for k in range(2):
            f = torch.full((rows, G), -1.0)
            for r in range(lay.row_lo, lay.row_hi):
                f[lay.local(r)] = r + 1000.0 * k
            fields.append(f)
        comm = TorchDistComm()
        comm.serve(lay, "exchange", [HaloSpec(fields[0], h), HaloSpec(fields[1], 1)])
        for k, hh in ((0, h), (1, 1)):
            f = fields[k]
            for r in range(lay.row_lo - hh, lay.row_hi + hh):
                if 0 <= r < G:
                    assert torch.all(f[lay.local(r)] == r + 1000.0 * k), (rank, k, r)
            if rank > 0 and halo > hh:       # rows further out were not requested: untouched
                assert torch.all(f[lay.local(lay.row_lo - hh - 1)] == -1.0)
        # overlapped form + scalar MAX
        fields[0][lay.local(lay.row_lo)] = 7.0 + rank
        comm.serve(lay, "exchange_begin", [HaloSpec(fields[0], 1)])
        comm.serve(lay, "exchange_end", None)
        if rank < world - 1:
            assert torch.all(fields[0][lay.local(lay.row_hi)] == 7.0 + rank + 1)
        m = comm.serve(lay, "allreduce_max", float(rank) * 0.5)
        assert m == (world - 1) * 0.5
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, 48, 6, 4), nprocs=world, join=True)


# ---- peer-memory slabs: the only host-side communication is the exchange of the arenas' IPC handles ----
def test_neighbour_links():
    G = 1024
    hs = [bytes([r]) * 64 for r in range(4)]
    assert neighbour_links(G, 0, 4, hs) == [(SF.SF_SLAB_DOWN, hs[1], 256, 512)]
    assert neighbour_links(G, 2, 4, hs) == [(SF.SF_SLAB_UP, hs[1], 256, 512), (SF.SF_SLAB_DOWN, hs[3], 768, 1024)]
    assert neighbour_links(G, 3, 4, hs) == [(SF.SF_SLAB_UP, hs[2], 512, 768)]
    assert neighbour_links(G, 0, 1, hs[:1]) == []


def _handle_worker(rank, world, port, G):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = bytes([rank + 1]) * 64                     # stands in for cudaIpcGetMemHandle's 64 bytes
        links = exchange_handles(G, rank, world, mine)
        parts = partition_rows(G, world)
        want = []
        if rank > 0:
            want.append((SF.SF_SLAB_UP, bytes([rank]) * 64, *parts[rank - 1]))
        if rank < world - 1:
            want.append((SF.SF_SLAB_DOWN, bytes([rank + 2]) * 64, *parts[rank + 1]))
        assert links == want, (rank, links, want)
        with pytest.raises(ValueError):
            exchange_handles(G, rank, world, b"short")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ipc_handle_exchange_gloo(world):
    port = _free_port()
    mp.spawn(_handle_worker, args=(world, port, 130), nprocs=world, join=True)
