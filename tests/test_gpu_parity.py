"""GPU parity tests proper: the CUDA time-march (through the C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): cell indexing / particle-to-cell assignment bit-exact; potentials within 1e-9
relative L2 after 100 steps; radiated power within 1 %.  What is actually achieved is much tighter and asserted so:
  * one fieldUpdate from identical state: potentials BIT-identical (same association order, no FMA contraction);
  * E/B at the nodes the reference evaluates: bit-identical floats;
  * push from identical state: positions/momenta to 1e-12 (CUDA libm vs glibc last-ulp differences in sin/cosh/exp);
  * deposit from identical particles: J to 1e-12 relative L2 (FP64 atomics reorder the sums);
  * 100 coupled steps: A, phi, particles, power to 1e-9.
"""
import numpy as np
import pytest

from mithra_b200 import abi
from oracle import binding
from tests import helpers

pytestmark = pytest.mark.gpu


def _pair(job):
    p, meta, g = helpers.params_for(job)
    gpu, cpu = abi.GpuSolver(p), binding.Oracle(p)
    for s in (gpu, cpu):
        helpers.start_from_golden(s, g)
    return p, g, gpu, cpu


def _field_names(p):
    return ("anp1", "an", "anm1") + (("fnp1", "fn", "fnm1") if p.space_charge else ())


def _sync_gpu_to_cpu(p, gpu, cpu):
    """Copy the oracle's complete state (start of a field step) to the GPU."""
    f = cpu.download_fields(_field_names(p))
    gpu.upload_fields(an=f["an"], anm1=f["anm1"], jn=f["anp1"], fn=f.get("fn"), fnm1=f.get("fnm1"), rho=f.get("fnp1"))
    gpu.upload_particles(cpu.download_particles())
    gpu.set_time(cpu.lib.oracle_time(cpu.o), cpu.lib.oracle_time_bunch(cpu.o), gpu.get_time()[2])


@pytest.mark.parametrize("job", helpers.JOBS)
def test_initial_state(job):
    p, g, gpu, cpu = _pair(job)
    a, b = gpu.download_fields(("an", "anm1")), cpu.download_fields(("an", "anm1"))
    for k in a:
        # seed potential: cos/exp/atan of CUDA libm vs glibc
        assert helpers.rel_l2(a[k], b[k]) < 1e-12 if p.seed_enabled else np.array_equal(a[k], b[k])
    np.testing.assert_array_equal(gpu.download_particles(), cpu.download_particles())
    np.testing.assert_array_equal(gpu.push_cells(), cpu.push_cells())
    np.testing.assert_array_equal(gpu.deposit_cells(), cpu.deposit_cells())


@pytest.mark.parametrize("job", helpers.JOBS)
def test_phases_from_identical_state(job):
    """Advance the oracle 70 steps (bunch inside the undulator fringe, J != 0, e = 1), copy its state to the GPU and
    compare every phase of the next step separately."""
    p, g, gpu, cpu = _pair(job)
    for _ in range(70):
        helpers.solve_step(cpu)
    _sync_gpu_to_cpu(p, gpu, cpu)
    np1 = ("anp1", "fnp1") if p.space_charge else ("anp1",)

    # 1. fieldUpdate: bit-identical potentials (the seeded job differs by the libm of the injected seed only)
    gpu.fieldUpdate(); cpu.fieldUpdate()
    a, b = gpu.download_fields(np1), cpu.download_fields(np1)
    for k in np1:
        if p.seed_enabled:
            assert helpers.rel_l2(a[k], b[k]) < 1e-13, k
        else:
            np.testing.assert_array_equal(a[k], b[k], err_msg=k)

    # 2. push (same E/B source); cell indices before the push are those of identical positions
    np.testing.assert_array_equal(gpu.push_cells(), cpu.push_cells())
    gpu.bunchUpdate(); cpu.bunchUpdate()
    pg, pc = gpu.download_particles(), cpu.download_particles()
    assert pg.shape == pc.shape
    np.testing.assert_array_equal(pg[:, 0], pc[:, 0])
    np.testing.assert_array_equal(pg[:, 4:7], pc[:, 4:7])          # rnm = start-of-step position
    np.testing.assert_array_equal(pg[:, 10], pc[:, 10])            # entrance flag
    assert helpers.rel_l2(pg[:, 1:4], pc[:, 1:4]) < 1e-12
    assert helpers.rel_l2(pg[:, 7:10], pc[:, 7:10]) < 1e-12

    # 3. E/B on every node the reference evaluated lazily: bit-identical floats
    en_g, bn_g, mask_g = gpu.download_eb()
    en_c, bn_c, pic = cpu.download_eb()
    # the reference also evaluates the two boundary planes of every slab inside fieldUpdate (fdtd.cpp:742-774); the GPU
    # evaluates the padded particle box only (the power kernel evaluates its own plane), so look at what particles touched
    kk = np.arange(pic.size) // (p.N0 * p.N1)
    idx = np.flatnonzero((pic != 0) & (kk > 1) & (kk < p.np - 2))
    assert idx.size > 0
    assert mask_g[idx].all(), "GPU did not evaluate E/B on a node the reference used"
    if p.seed_enabled:
        np.testing.assert_allclose(en_g.reshape(-1, 3)[idx], en_c.reshape(-1, 3)[idx], rtol=1e-5, atol=1e-30)
    else:
        np.testing.assert_array_equal(en_g.reshape(-1, 3)[idx], en_c.reshape(-1, 3)[idx])
        np.testing.assert_array_equal(bn_g.reshape(-1, 3)[idx], bn_c.reshape(-1, 3)[idx])

    # 4. power sample of this step
    gpu.screenProfile(); cpu.screenProfile()
    gpu.powerSample(); cpu.powerSample()

    # 5. deposit from IDENTICAL particles
    gpu.upload_particles(pc)
    gpu.fieldShift(); cpu.fieldShift()
    gpu.currentReset(); cpu.currentReset()
    np.testing.assert_array_equal(gpu.deposit_cells(), cpu.deposit_cells())
    gpu.currentUpdate(); cpu.currentUpdate()
    gpu.currentCommunicate()
    a, b = gpu.download_fields(np1), cpu.download_fields(np1)
    for k in np1:
        assert helpers.rel_l2(a[k], b[k]) < 1e-12, k
        # the deposit touches exactly the same nodes
        np.testing.assert_array_equal(a[k] != 0.0, b[k] != 0.0, err_msg=k + " support")


@pytest.mark.parametrize("job", helpers.JOBS)
def test_100_steps(job):
    p, g, gpu, cpu = _pair(job)
    gpu.step(100)
    for _ in range(100):
        helpers.solve_step(cpu)
    names = _field_names(p)
    a, b = gpu.download_fields(names), cpu.download_fields(names)
    for k in names:
        assert helpers.rel_l2(a[k], b[k]) < 1e-9, k
    pg, pc = gpu.download_particles(), cpu.download_particles()
    assert helpers.rel_l2(pg[:, 1:4], pc[:, 1:4]) < 1e-9
    assert helpers.rel_l2(pg[:, 7:10], pc[:, 7:10]) < 1e-9
    np.testing.assert_array_equal(pg[:, 10], pc[:, 10])
    np.testing.assert_array_equal(gpu.deposit_cells(), cpu.deposit_cells())
    np.testing.assert_array_equal(gpu.push_cells(), cpu.push_cells())
    pw_g, pw_c = gpu.fetch_power(), cpu.fetch_power()
    assert pw_g.shape == pw_c.shape == (100, p.power.N * p.power.Nl)
    np.testing.assert_allclose(pw_g, pw_c, rtol=1e-8, atol=1e-12 * np.abs(pw_c).max())
    if p.screens.enabled:
        for s in range(p.screens.N):
            rg, rc = gpu.fetch_screen(s), cpu.fetch_screen(s)
            assert rg.shape == rc.shape
            np.testing.assert_allclose(rg, rc, rtol=1e-9, atol=1e-12)
    if p.power_map.enabled:
        # Solver::powerVisualize: per-pixel map of the last step against the oracle and the reference's own
        mg, mc = gpu.fetch_power_map(), cpu.fetch_power_map()
        np.testing.assert_allclose(mg, mc, rtol=1e-7, atol=1e-10 * np.abs(mc).max())
        np.testing.assert_allclose(mg, g["pmap"], rtol=1e-7, atol=1e-10 * np.abs(mc).max())
    # and against the reference's own golden output (power curve within 1 % is the contract; we are far inside)
    np.testing.assert_allclose(pw_g, g["power"], rtol=1e-8, atol=1e-12 * np.abs(g["power"]).max())
    assert helpers.rel_l2(pg, g["p100"]) < 1e-9


@pytest.mark.parametrize("job", ["micro-sc", "micro-seeded", "micro-optical"])
def test_bunch_sorted_by_cell_every_step(job):
    """The counting sort by cell (kernels_sort.cuh) re-orders the bunch in device memory; results, the order of the
    downloaded particles and of the screen records, and the cell assignment must not change."""
    p, meta, g = helpers.params_for(job)
    p.sort_interval = 1
    gpu, cpu = abi.GpuSolver(p), binding.Oracle(p)
    for s in (gpu, cpu):
        helpers.start_from_golden(s, g)
    gpu.step(100)
    for _ in range(100):
        helpers.solve_step(cpu)
    names = _field_names(p)
    a, b = gpu.download_fields(names), cpu.download_fields(names)
    for k in names:
        assert helpers.rel_l2(a[k], b[k]) < 1e-9, k
    pg, pc = gpu.download_particles(), cpu.download_particles()
    np.testing.assert_array_equal(pg[:, 0], pc[:, 0])
    assert helpers.rel_l2(pg[:, 1:4], pc[:, 1:4]) < 1e-9
    assert np.abs(pg[:, 1:4] - pc[:, 1:4]).max() < 1e-7 * np.abs(pc[:, 1:4]).max()      # row by row, not only in the norm
    np.testing.assert_array_equal(gpu.deposit_cells(), cpu.deposit_cells())
    np.testing.assert_array_equal(gpu.push_cells(), cpu.push_cells())
    np.testing.assert_allclose(gpu.fetch_power(), cpu.fetch_power(), rtol=1e-8, atol=1e-12 * np.abs(cpu.fetch_power()).max() + 1e-300)
    if p.screens.enabled:
        for s in range(p.screens.N):
            rg, rc = gpu.fetch_screen(s), cpu.fetch_screen(s)
            assert rg.shape == rc.shape
            np.testing.assert_allclose(rg, rc, rtol=1e-9, atol=1e-12)


def test_explicit_sort_keeps_the_upload_order():
    p, g, gpu, cpu = _pair("micro-nsfd")
    before = gpu.download_particles()
    gpu.sortParticles()
    np.testing.assert_array_equal(gpu.download_particles(), before)
    np.testing.assert_array_equal(gpu.deposit_cells(), cpu.deposit_cells())
    rng = np.random.default_rng(3)
    shuffled = before[rng.permutation(before.shape[0])]
    gpu.upload_particles(shuffled)
    gpu.sortParticles(); gpu.sortParticles()
    np.testing.assert_array_equal(gpu.download_particles(), shuffled)


def test_tabulated_seed_equals_full_evaluation(monkeypatch):
    """A seed along +z is injected from the per-plane table (seed_plane_table); MITHRA_SEED_GENERIC=1 evaluates
    Seed::fields in full at every shell node.  Same carrier phase bit for bit, envelope to rounding."""
    p, meta, g = helpers.params_for("micro-seeded")
    fast = abi.GpuSolver(p)
    monkeypatch.setenv("MITHRA_SEED_GENERIC", "1")
    full = abi.GpuSolver(p)
    monkeypatch.delenv("MITHRA_SEED_GENERIC")
    for s in (fast, full):
        helpers.start_from_golden(s, g)
        s.step(50)
    a, b = fast.download_fields(("an",))["an"], full.download_fields(("an",))["an"]
    assert np.abs(b).max() > 0
    assert helpers.rel_l2(a, b) < 1e-13


def test_step_entry_points_equal_fused_step():
    """mithra_gpu_step == the nine per-method entry points in the reference's order."""
    p, g, gpu, cpu = _pair("micro-nsfd")
    cpu.close()
    gpu2 = abi.GpuSolver(p)
    helpers.start_from_golden(gpu2, g)
    gpu.step(20)
    for _ in range(20):
        helpers.solve_step(gpu2)
    a, b = gpu.download_fields(("an",)), gpu2.download_fields(("an",))
    assert helpers.rel_l2(a["an"], b["an"]) < 1e-12
    assert gpu.counters().field_steps == 20 and gpu2.counters().field_steps == 20
    assert gpu.counters().cell_updates == 20 * p.N0 * p.N1 * p.np


def test_empty_bunch_and_errors():
    p, g, gpu, cpu = _pair("micro-nsfd")
    gpu.upload_particles(np.zeros((0, 11)))
    gpu.step(3)                                  # no particles: fields only
    assert gpu.download_particles().shape == (0, 11)
    with pytest.raises(RuntimeError):
        p2, _, _ = helpers.params_for("micro-nsfd", max_particles=4)
        s = abi.GpuSolver(p2)
        s.upload_particles(np.zeros((8, 11)))


def _resized(job, N0, N1, npl):
    """The parameter block of a micro job on a larger mesh (same cell sizes and update coefficients)."""
    p, meta, g = helpers.params_for(job, max_particles=1024)
    p.N0, p.N1, p.N2, p.np = N0, N1, npl, npl
    p.xmin, p.xmax = -0.5 * (N0 - 1) * p.dx, 0.5 * (N0 - 1) * p.dx
    p.ymin, p.ymax = -0.5 * (N1 - 1) * p.dy, 0.5 * (N1 - 1) * p.dy
    p.zmax = p.zmin + (npl - 1) * p.dz
    p.zp[0], p.zp[1], p.Lz = p.zmin, p.zmax, p.zmax - p.zmin
    p.power.enabled, p.screens.enabled = 0, 0
    return p, g


@pytest.mark.parametrize("job,shape", [("micro-seeded", (14, 14, 242)), ("micro-seeded", (37, 85, 70)), ("micro-seeded", (80, 41, 9)),
                                       ("micro-seeded", (8, 200, 12)), ("micro-seeded", (31, 8, 75)), ("micro-sc", (29, 53, 66)),
                                       ("micro-o1", (33, 101, 40)), ("micro-fd", (64, 9, 130)), ("micro-nsfd", (85, 85, 70)),
                                       ("micro-nsfd", (102, 102, 40)), ("micro-nsfd", (20, 73, 30)), ("micro-fd", (40, 27, 33)),
                                       ("micro-sc", (8, 200, 12)), ("micro-o1", (31, 8, 75)), ("micro-nsfd", (14, 14, 242)),
                                       ("micro-sc", (9, 513, 20)), ("micro-nsfd", (140, 73, 12)), ("micro-fd", (40, 39, 20)),
                                       ("micro-seeded", (85, 85, 70)), ("micro-seeded", (40, 102, 30)), ("micro-seeded", (9, 513, 20)),
                                       ("micro-seeded", (140, 73, 12)), ("micro-nsfd", (402, 402, 10))])
def test_fused_stencil_equals_separate_kernels(job, shape, monkeypatch):
    """The production path -- jobs without a seed: stencil_stream with its face warp (interior value of every node and the
    y faces in one kernel, the x faces as a pass over whole rows); seeded jobs, and meshes with more than 32 face-warp
    nodes per tile (rows shorter than about 30 nodes): stencil_stream on the inner nodes + rim_update -- against the
    seeded variant of the face warp (MITHRA_SEEDWARP: y-shell seed terms in the face warp, x-shell terms as a pass over
    whole rows), against the reference's three passes as
    separate kernels (MITHRA_NO_FUSE), against stencil_stream + rim_update (MITHRA_NO_FACEWARP) and against the plain-load
    stencil (MITHRA_STENCIL_PLAIN): bit-identical potentials, on meshes whose 448- / 480-node tiles cut through rows (102:
    the FEL-LCLS row; 140 x 73: a tile starts on the node next to a y face; 39: a tile ends on one; 513: a tile inside
    row 1) and down to the smallest mesh the rim path takes (8 nodes across)."""
    p, g = _resized(job, *shape)
    rng = np.random.default_rng(11)
    n = p.N0 * p.N1 * p.np
    an, anm1, jn = (rng.standard_normal(n * 3) for _ in range(3))
    jn[rng.random(n * 3) < 0.7] = 0.0
    sc = {}
    if p.space_charge:
        sc = dict(fn=rng.standard_normal(n), fnm1=rng.standard_normal(n), rho=rng.standard_normal(n))
    # particles only to put the deposit box (the nodes where J is read) somewhere inside the mesh
    bunch = np.zeros((64, 11))
    bunch[:, 0] = 1.0
    bunch[:, 1] = rng.uniform(0.6 * p.xmin, 0.6 * p.xmax, 64)
    bunch[:, 2] = rng.uniform(0.6 * p.ymin, 0.6 * p.ymax, 64)
    bunch[:, 3] = rng.uniform(p.zmin + 2 * p.dz, p.zmax - 2 * p.dz, 64)
    bunch[:, 4:7] = bunch[:, 1:4] + 0.3 * np.array([p.dx, p.dy, p.dz])
    bunch[:, 10] = 1.0
    names = ("anp1", "an", "anm1") + (("fnp1", "fn", "fnm1") if p.space_charge else ())
    out = {}
    # MITHRA_SEEDWARP: seeded jobs through stencil_stream's face warp + seed_xshell_rows (opt-in: it loses on FEL-SEEDED)
    # MITHRA_FACEWARP: the face warp also on wide meshes (402 x 402, the FEL-ICS plane: the default there is rim_update)
    modes = ("MITHRA_NO_FUSE", "MITHRA_NO_FACEWARP", "MITHRA_STENCIL_PLAIN", "MITHRA_FACEWARP") + (("MITHRA_SEEDWARP",) if job == "micro-seeded" else ())
    for mode in ("fused",) + modes:
        if mode != "fused":
            monkeypatch.setenv(mode, "1")
        s = abi.GpuSolver(p)
        s.set_time(0.37, 0.37, 5)
        s.upload_fields(an=an, anm1=anm1, jn=jn, **sc)
        s.upload_particles(bunch)
        s.currentReset(); s.currentUpdate()            # J = deposit of the bunch; the random jn only pre-fills it
        for _ in range(3):
            # the reference's J lives in anp1_ and does not survive a field update (fdtd.cpp:244): deposit it again
            s.fieldUpdate(); s.fieldShift(); s.currentReset(); s.currentUpdate(); s.advanceTime()
        s.fieldUpdate()
        out[mode] = s.download_fields(names)
        s.close()
        if mode != "fused":
            monkeypatch.delenv(mode)
    assert np.abs(out["fused"]["anp1"]).max() > 0
    for mode in modes:
        for k in names:
            np.testing.assert_array_equal(out["fused"][k], out[mode][k], err_msg="%s %s" % (mode, k))


@pytest.mark.parametrize("job,shape", [("micro-nsfd", (14, 14, 242)), ("micro-sc", (29, 53, 66)), ("micro-seeded", (85, 85, 70)),
                                       ("micro-fd", (41, 100, 37)), ("micro-sc", (9, 9, 12))])
def test_eb_march_equals_node_kernel(job, shape, monkeypatch):
    """E/B over the particle box: the production z-marching kernel (eval_eb_march over the pencils the mask marks + the
    two copied end planes) against the node-at-a-time kernel over the whole box (MITHRA_EB_BOX) and against the march
    without the mask (MITHRA_NO_EBMASK): bit-identical floats on every evaluated node, the masked set lies inside the
    box and covers every node within a particle's reach of one field step -- boxes wider than one 32 x 8 tile,
    boxes that reach the first and the last plane, and a box that is the whole (tiny) mesh."""
    p, g = _resized(job, *shape)
    rng = np.random.default_rng(23)
    n = p.N0 * p.N1 * p.np
    an, anm1 = rng.standard_normal(n * 3), rng.standard_normal(n * 3)
    sc = dict(fn=rng.standard_normal(n), fnm1=rng.standard_normal(n)) if p.space_charge else {}
    nb = 24
    bunch = np.zeros((nb, 11))
    bunch[:, 0] = 1.0
    bunch[:, 1] = rng.uniform(0.8 * p.xmin, 0.8 * p.xmax, nb)
    bunch[:, 2] = rng.uniform(0.8 * p.ymin, 0.8 * p.ymax, nb)
    bunch[:, 3] = rng.uniform(p.zmin + 1e-3 * p.dz, p.zmax - 1e-3 * p.dz, nb)
    bunch[0, 3], bunch[1, 3] = p.zmin + 1e-3 * p.dz, p.zmax - 1e-3 * p.dz          # first and last cell in z
    bunch[:, 4:7] = bunch[:, 1:4]
    bunch[:, 10] = 1.0
    out = {}
    for mode in ("march", "MITHRA_EB_BOX", "MITHRA_NO_EBMASK"):
        if mode != "march":
            monkeypatch.setenv(mode, "1")
        s = abi.GpuSolver(p)
        s.set_time(0.37, 0.37, 5)
        s.upload_fields(an=an, anm1=anm1, **sc)
        s.upload_particles(bunch)
        s.fieldUpdate()
        out[mode] = s.download_eb()
        s.close()
        if mode != "march":
            monkeypatch.delenv(mode)
    e0, b0, m0 = out["march"]
    e1, b1, m1 = out["MITHRA_EB_BOX"]
    e2, b2, m2 = out["MITHRA_NO_EBMASK"]
    assert np.abs(e0).max() > 0 and np.abs(b0).max() > 0 and m0.sum() > 0
    # without the mask the march covers the box exactly like the node kernel
    np.testing.assert_array_equal(m2, m1)
    np.testing.assert_array_equal(e2.view(np.uint32), e1.view(np.uint32))
    np.testing.assert_array_equal(b2.view(np.uint32), b1.view(np.uint32))
    # with it: a subset of the box, same bits where evaluated
    assert not np.any(m0 & ~m1.astype(bool))
    w = np.repeat(m0.astype(bool), 3)
    np.testing.assert_array_equal(e0.view(np.uint32)[w], e1.view(np.uint32)[w])
    np.testing.assert_array_equal(b0.view(np.uint32)[w], b1.view(np.uint32)[w])
    # every node within a particle's reach of one field step (it travels less than c dt) is evaluated: per axis the nodes
    # cell(r - c dt) .. cell(r + c dt) + 1
    M = m0.reshape(p.np, p.N0, p.N1)
    cdt = p.c0 * p.dt
    lo = [np.floor((bunch[:, 1 + a] - cdt - o) / d).astype(int) for a, (o, d) in enumerate(((p.xmin, p.dx), (p.ymin, p.dy), (p.zmin, p.dz)))]
    hi = [np.floor((bunch[:, 1 + a] + cdt - o) / d).astype(int) + 1 for a, (o, d) in enumerate(((p.xmin, p.dx), (p.ymin, p.dy), (p.zmin, p.dz)))]
    i = np.floor((bunch[:, 1] - p.xmin) / p.dx).astype(int)
    j = np.floor((bunch[:, 2] - p.ymin) / p.dy).astype(int)
    for t in range(nb):
        if not (1 <= i[t] <= p.N0 - 3 and 1 <= j[t] <= p.N1 - 3):
            continue                                     # outside x/y (min + d, max - d): gathers nothing (solver.cpp:1444)
        for kk in range(lo[2][t], hi[2][t] + 1):
            for ii in range(lo[0][t], hi[0][t] + 1):
                for jj in range(lo[1][t], hi[1][t] + 1):
                    if 0 <= kk < p.np and 1 <= ii <= p.N0 - 2 and 1 <= jj <= p.N1 - 2:
                        assert M[kk, ii, jj], (t, kk, ii, jj)
    if min(shape) > 12:
        assert m0.sum() < m1.sum()                       # the mask does skip something on a sparse bunch


@pytest.mark.parametrize("job", ["micro-nsfd", "micro-sc", "micro-fd", "micro-seeded"])
def test_field_sample_matches_oracle(job):
    """mithra_gpu_field_sample (FdTd::fieldSample, fdtd.cpp:851-913) against the oracle's restatement after a field update
    from identical state: E/B at the 8 nodes are bit-identical floats and the interpolation runs in the reference's
    order, so et, bt, at are bit-identical doubles (seeded job: the libm of the injected seed -- 1e-12 on A, a float ulp
    on E and B); points anywhere
    inside the mesh, in the first and in the last cell the reference admits."""
    p, g, gpu, cpu = _pair(job)
    for _ in range(40):
        helpers.solve_step(cpu)
    _sync_gpu_to_cpu(p, gpu, cpu)
    gpu.fieldUpdate(); cpu.fieldUpdate()
    rng = np.random.default_rng(3)
    n = 200
    pos = np.empty((n, 3))
    pos[:, 0] = rng.uniform(p.xmin + 1.001 * p.dx, p.xmax - 1.001 * p.dx, n)
    pos[:, 1] = rng.uniform(p.ymin + 1.001 * p.dy, p.ymax - 1.001 * p.dy, n)
    pos[:, 2] = rng.uniform(p.zmin + 1.001 * p.dz, p.zmax - 1.001 * p.dz, n)
    pos[0] = [p.xmin + 1.001 * p.dx, p.ymin + 1.001 * p.dy, p.zmin + 1.001 * p.dz]
    pos[1] = [p.xmax - 1.001 * p.dx, p.ymax - 1.001 * p.dy, p.zmax - 1.001 * p.dz]
    got, mine = gpu.field_sample(pos)
    want = cpu.field_sample(pos)
    assert mine.all() and np.abs(want).max() > 0
    if p.seed_enabled:
        # E/B are FLOATS of differences of potentials that carry the 1e-13 of the seed's libm: the last float bit may flip
        assert np.all(np.abs(got - want) <= 1e-6 * np.abs(want).max(axis=0))
        np.testing.assert_allclose(got[:, 6:], want[:, 6:], rtol=1e-9, atol=1e-12 * np.abs(want[:, 6:]).max())
    else:
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("d", [60.0, 9.19059968, -0.01532827, 30.0, 4.59529984, -4.49297199e+08, 3.0, 1.9999999999999998,
                               1.0000000000000002, 1e-3, 6.02e23, 1.7e-19, 1e-200])
def test_constant_divisor_division_is_ieee(d):
    """div_by (reciprocal + two FMAs, device_types.cuh) against the true division on the device, bit for bit: random
    mantissas over 600 binades, the special values, and operands in the ranges that must take the fallback."""
    rng = np.random.default_rng(5)
    n = 4_000_000
    x = rng.standard_normal(n) * np.exp2(rng.integers(-300, 300, n).astype(np.float64))
    x[:16] = [0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 2.2250738585072014e-308, 1.7976931348623157e308,
              1e-160, -1e-160, 1e160, 1.0, -1.0, 1e-149, 1e151]
    x[16:1000] = rng.standard_normal(984) * 1e-305
    assert abi.selftest_divide(x, d) == 0


def test_bunch_moments_reduction():
    """mithra_gpu_bunch_moments (Solver::bunchSample's sums, reduced on the device) against the same sums in numpy."""
    p, g, gpu, cpu = _pair("micro-sc")
    gpu.step(30)
    q = gpu.download_particles()
    w = q[:, 0]
    want = np.concatenate(([w.sum()], (w[:, None] * q[:, 1:4]).sum(0), (q[:, 1:4] ** 2 * w[:, None]).sum(0),
                           (w[:, None] * q[:, 7:10]).sum(0), (q[:, 7:10] ** 2 * w[:, None]).sum(0)))
    got = gpu.bunch_moments()
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12 * np.abs(want).max())
    # deterministic: the same reduction twice gives the same bits
    np.testing.assert_array_equal(got, gpu.bunch_moments())
