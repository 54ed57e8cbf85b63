"""z-slab partition of the mesh over the GPUs of one box -- the reference's own scheme with size_ := number of GPUs
(Solver::initializeMesh, solver.cpp:619-641 for np_/k0_, solver.cpp:677-680 for the ownership interval zp_), and the
host-side helpers that scatter a global state over the slabs and gather it back (harness code: tests, bench)."""
import copy

import numpy as np


def slab_extent(N2, rank, size):
    """(np, k0) of slab `rank`: two planes are shared with each neighbour."""
    if size == 1:
        return N2, 0
    q = N2 // size
    if rank == 0:
        return q + 1, 0
    if rank == size - 1:
        return N2 - (size - 1) * q + 1, (size - 1) * q - 1
    return q + 2, rank * q - 1


def slab_params(p, rank, size):
    """Parameter block of one slab from the single-slab block `p` (which must describe the whole mesh)."""
    assert p.size == 1 and p.k0 == 0 and p.np == p.N2
    q = copy.copy(p)
    q.np, q.k0 = slab_extent(p.N2, rank, size)
    q.rank, q.size = rank, size
    if size > 1:
        # z of local plane 0 and of local plane np-2 (np-1 on the last slab), as Solver::rc computes them
        q.zp[0] = p.zmin + (0 + q.k0) * p.dz
        q.zp[1] = p.zmin + ((q.np - (1 if rank == size - 1 else 2)) + q.k0) * p.dz
    return q


def owned_planes(N2, rank, size):
    """Global plane indices whose potentials slab `rank` computes itself (every plane belongs to exactly one slab)."""
    npl, k0 = slab_extent(N2, rank, size)
    lo = 0 if rank == 0 else 1
    hi = npl if rank == size - 1 else npl - 1
    return np.arange(k0 + lo, k0 + hi)


def owner_of(p, z, size):
    """Slab that owns a particle at z: wrapped position inside [zp0, zp1) (solver.cpp:1440-1441)."""
    zr = np.mod(z - p.zmin, p.Lz) + p.zmin
    out = np.full(zr.shape, -1, dtype=np.int64)
    for r in range(size):
        q = slab_params(p, r, size)
        out[(zr >= q.zp[0]) & (zr < q.zp[1])] = r
    return out


def scatter_field(p, arr, ncomp, rank, size):
    """Local part (reference slab numbering, ghosts included) of a global array double[N2*N0*N1][ncomp]."""
    npl, k0 = slab_extent(p.N2, rank, size)
    a = np.asarray(arr).reshape(p.N2, p.N0 * p.N1 * ncomp)
    return np.ascontiguousarray(a[k0:k0 + npl]).reshape(-1)


def gather_field(p, parts, ncomp, size):
    """Global array from the slabs' local arrays, taking every plane from the slab that computes it."""
    out = np.zeros((p.N2, p.N0 * p.N1 * ncomp))
    for r in range(size):
        npl, k0 = slab_extent(p.N2, r, size)
        loc = np.asarray(parts[r]).reshape(npl, -1)
        g = owned_planes(p.N2, r, size)
        out[g] = loc[g - k0]
    return out.reshape(-1)
