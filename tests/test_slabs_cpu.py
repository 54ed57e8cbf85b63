"""Host-side logic of the z-slab partition (no GPU): the reference's np_/k0_/zp_ formulas, scatter/gather round trips,
and the neighbour-blob exchange over torch.distributed (gloo, world_size 2) that bench.py uses under torchrun."""
import os
import socket

import numpy as np
import pytest

from mithra_b200 import slabs
from tests import helpers


@pytest.mark.parametrize("N2,size", [(242, 2), (242, 3), (8252, 8), (33335, 8), (502, 5), (64, 4)])
def test_partition_tiles_the_mesh(N2, size):
    seen = np.zeros(N2, dtype=int)
    for r in range(size):
        npl, k0 = slabs.slab_extent(N2, r, size)
        assert npl >= 4 and k0 >= 0 and k0 + npl <= N2
        if r + 1 < size:
            n2, k2 = slabs.slab_extent(N2, r + 1, size)
            assert k0 + npl - 2 == k2, "two planes overlap: local (np-2, np-1) of r == local (0, 1) of r+1"
        seen[slabs.owned_planes(N2, r, size)] += 1
    assert (seen == 1).all()
    assert slabs.slab_extent(N2, size - 1, size)[0] + slabs.slab_extent(N2, size - 1, size)[1] == N2


def test_ownership_intervals_are_contiguous_and_exclusive():
    p, _, g = helpers.params_for("micro-nsfd")
    for size in (2, 3, 4):
        q = [slabs.slab_params(p, r, size) for r in range(size)]
        assert q[0].zp[0] == p.zmin + 0 * p.dz
        for r in range(size - 1):
            assert q[r].zp[1] == q[r + 1].zp[0]
        z = g["p0"][:, 3]
        own = slabs.owner_of(p, z, size)
        assert (own >= 0).all()
        for r in range(size):
            sel = own == r
            assert ((z[sel] >= q[r].zp[0]) & (z[sel] < q[r].zp[1])).all()


def test_scatter_gather_round_trip():
    p, _, _ = helpers.params_for("micro-nsfd")
    rng = np.random.RandomState(1)
    a = rng.rand(p.N2 * p.N0 * p.N1 * 3)
    for size in (2, 3):
        parts = [slabs.scatter_field(p, a, 3, r, size) for r in range(size)]
        for r in range(size):
            assert parts[r].size == slabs.slab_extent(p.N2, r, size)[0] * p.N0 * p.N1 * 3
        np.testing.assert_array_equal(slabs.gather_field(p, parts, 3, size), a)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class Fake:                                   # stands in for abi.GpuSolver: same two methods, no GPU
        def __init__(self):
            self.got = None

        def export_blob(self):
            return b"blob-of-rank-%d\0tail" % rank

        def connect(self, prev, nxt):
            self.got = (prev, nxt)

    from mithra_b200 import abi
    f = Fake()
    abi.GpuSolver.connect_neighbours(f, dist, rank, world)
    out.put((rank, f.got))
    dist.destroy_process_group()


def test_blob_exchange_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert got[0] == (b"blob-of-rank-1\0tail", b"blob-of-rank-1\0tail")      # ring of two: prev == next
    assert got[1] == (b"blob-of-rank-0\0tail", b"blob-of-rank-0\0tail")
