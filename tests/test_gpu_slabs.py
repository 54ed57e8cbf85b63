"""Multi-slab self-consistency on the GPU: the mesh cut into S z-slabs (one handle per slab, all on the test GPU,
exchanging ghost planes, boundary currents and migrating particles through the same peer-memory path that connects
the GPUs of a box) must reproduce the single-slab run -- up to the summation order of the deposited current.
This is the k-GPU parity requirement of SURVEY.md section 8(e): the SINGLE-rank reference result, including both
half-segments of a particle that crosses a slab boundary (reference quirk Q13)."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from mithra_b200 import abi, slabs
from oracle import binding
from tests import helpers

pytestmark = pytest.mark.gpu

PHASES = ("fieldUpdate", "bunchUpdate", "screenProfile", "powerSample", "fieldShift", "currentReset", "currentUpdate",
          "currentCommunicate", "migrateBegin", "migrateEnd", "advanceTime")


def make_slabs(p, g, size, sort_interval=0):
    p.sort_interval = sort_interval
    parts = [abi.GpuSolver(slabs.slab_params(p, r, size)) for r in range(size)]
    blobs = [s.export_blob() for s in parts]
    for r, s in enumerate(parts):
        s.connect(blobs[(r - 1) % size], blobs[(r + 1) % size])
    own = slabs.owner_of(p, g["p0"][:, 3], size)
    t = g["t0"]
    for r, s in enumerate(parts):
        s.set_time(float(t[0]), float(t[1]), int(t[2]))
        s.upload_particles(g["p0"][own == r])
        if p.seed_enabled:
            s.seedInitial()
    return parts


def step_all(parts, n):
    for _ in range(n):
        for ph in PHASES:
            for s in parts:
                getattr(s, ph)()


def match_particles(a, b):
    """Pair the particles of two runs (the order differs after migration) by nearest neighbour in (r, gb)."""
    assert a.shape == b.shape
    scale = np.abs(b[:, [1, 2, 3, 7, 8, 9]]).max(axis=0) + 1e-300
    ta = a[:, [1, 2, 3, 7, 8, 9]] / scale
    tb = b[:, [1, 2, 3, 7, 8, 9]] / scale
    d, idx = cKDTree(tb).query(ta)
    assert np.unique(idx).size == idx.size, "not a bijection"
    return d.max(), idx


@pytest.mark.parametrize("job,size,sort", [("micro-nsfd", 2, 0), ("micro-nsfd", 3, 3), ("micro-sc", 2, 1), ("micro-seeded", 3, 0), ("micro-fd", 4, 0)])
def test_slabs_reproduce_single_slab(job, size, sort):
    p, meta, g = helpers.params_for(job)
    nsteps = 100
    cpu = binding.Oracle(p)
    helpers.start_from_golden(cpu, g)
    for _ in range(nsteps):
        helpers.solve_step(cpu)

    parts = make_slabs(p, g, size, sort)
    n0 = sum(s.num_particles() for s in parts)
    assert n0 == g["p0"].shape[0]
    step_all(parts, nsteps)
    for s in parts:
        s.synchronize()
    assert sum(s.num_particles() for s in parts) == n0, "particles were lost or duplicated in the migration"

    names = ("an", "anm1", "anp1") + (("fn", "fnm1", "fnp1") if p.space_charge else ())
    loc = [s.download_fields(names) for s in parts]
    ref = cpu.download_fields(names)
    for k in names:
        nc = 3 if k.startswith("a") else 1
        glob = slabs.gather_field(p, [l[k] for l in loc], nc, size)
        if k in ("anp1", "fnp1"):
            # the deposited current of a plane shared by two slabs is complete on the slab that stencils it
            pass
        assert helpers.rel_l2(glob, ref[k]) < 1e-9, k

    allp = np.concatenate([s.download_particles() for s in parts])
    dmax, idx = match_particles(allp, cpu.download_particles())
    assert dmax < 1e-8
    # every slab holds exactly the particles it owns
    for r, s in enumerate(parts):
        z = s.download_particles()[:, 3]
        assert (slabs.owner_of(p, z, size) == r).all()

    pw = sum(s.fetch_power() for s in parts)
    ref_pw = cpu.fetch_power()
    np.testing.assert_allclose(pw, ref_pw, rtol=1e-8, atol=1e-12 * np.abs(ref_pw).max())
    np.testing.assert_allclose(pw, g["power"], rtol=1e-8, atol=1e-12 * np.abs(g["power"]).max())

    if p.screens.enabled:
        for sc in range(p.screens.N):
            rec = np.concatenate([s.fetch_screen(sc) for s in parts])
            want = cpu.fetch_screen(sc)
            assert rec.shape == want.shape
            o1, o2 = np.lexsort((rec[:, 0], rec[:, 2])), np.lexsort((want[:, 0], want[:, 2]))
            np.testing.assert_allclose(rec[o1], want[o2], rtol=1e-8, atol=1e-10)


def test_particles_do_cross_slab_boundaries():
    """The case above is only a migration test if particles really change slabs."""
    p, meta, g = helpers.params_for("micro-nsfd")
    parts = make_slabs(p, g, 3)
    before = [s.num_particles() for s in parts]
    step_all(parts, 100)
    after = [s.num_particles() for s in parts]
    assert sum(before) == sum(after)
    assert before != after


def test_unconnected_slab_fails_loudly():
    p, meta, g = helpers.params_for("micro-nsfd")
    s = abi.GpuSolver(slabs.slab_params(p, 0, 2))
    with pytest.raises(RuntimeError, match="not connected"):
        s.fieldUpdate()
