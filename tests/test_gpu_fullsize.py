"""Parity at BASELINE.json's full sizes: one whole field step of the CUDA path against the CPU oracle on the same synthetic
state bench.py uses, and properties that do not depend on the size -- exact scaling of the source-free field update, fused
against separate kernels.
  configs[1] FEL-SEEDED  85 x 85 x 8252 nodes, 4,194,304 macro-particles   whole mesh against the oracle
  configs[4] fdtdSC unit 102 x 102 x 4098 nodes (A + phi), 1,048,576       whole mesh against the oracle
  configs[3] FEL-LCLS    102 x 102 x 33,335 nodes, 8,388,608               the whole mesh on the GPU; the oracle restates three
                         z windows of it (both mesh ends and the middle) as slabs of the reference's own partition -- the
                         update is local, so the window's inner planes must equal the whole-mesh result bit for bit"""
import copy
import os

import numpy as np
import pytest

import bench
from mithra_b200 import abi, meta as mmeta
from oracle import binding
from tests import helpers

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _params(workload="fel-seeded", **over):
    p = mmeta.params_from_meta(dict(np.load(os.path.join(ROOT, "bench", workload + ".meta.npz"))))
    p.max_particles = bench.WORKLOADS[workload]["particles"] + 1024
    p.max_screen_records = 1 << 16
    for k, v in over.items():
        setattr(p, k, v)
    return p


def test_one_full_size_step_against_the_oracle():
    p = _params()
    n = 4194304
    bunch = bench.synthetic_bunch(p, n)
    a_n = bench.synthetic_potential(p)
    a_nm1 = a_n * 0.999
    tb = bench.undulator_time(p)
    gpu, cpu = abi.GpuSolver(p), binding.Oracle(copy.copy(p))
    for s in (gpu, cpu):
        s.set_time(tb, tb, 0)
        s.upload_fields(an=a_n, anm1=a_nm1)
        s.upload_particles(bunch)
    del a_nm1

    # fieldUpdate: stencil + TF/SF seed + absorbing boundaries over 59.6 M nodes (the seed differs by libm only)
    gpu.fieldUpdate(); cpu.fieldUpdate()
    a, b = gpu.download_fields(("anp1",))["anp1"], cpu.download_fields(("anp1",))["anp1"]
    assert np.abs(b).max() > 0
    assert helpers.rel_l2(a, b) < 1e-13
    inner = a.reshape(p.np, p.N0, p.N1, 3)[3:-3, 3:-3, 3:-3], b.reshape(p.np, p.N0, p.N1, 3)[3:-3, 3:-3, 3:-3]
    np.testing.assert_array_equal(inner[0], inner[1])            # away from the seed shell: bit-identical
    del a, b, inner

    # cell assignment of 4.2 M particles: bit-exact; push and deposit
    np.testing.assert_array_equal(gpu.push_cells(), cpu.push_cells())
    gpu.bunchUpdate(); cpu.bunchUpdate()
    pg, pc = gpu.download_particles(), cpu.download_particles()
    np.testing.assert_array_equal(pg[:, 4:7], pc[:, 4:7])
    assert helpers.rel_l2(pg[:, 1:4], pc[:, 1:4]) < 1e-12
    assert helpers.rel_l2(pg[:, 7:10], pc[:, 7:10]) < 1e-12
    gpu.upload_particles(pc)
    for s in (gpu, cpu):
        s.fieldShift(); s.currentReset()
    np.testing.assert_array_equal(gpu.deposit_cells(), cpu.deposit_cells())
    gpu.currentUpdate(); cpu.currentUpdate()
    a, b = gpu.download_fields(("anp1",))["anp1"], cpu.download_fields(("anp1",))["anp1"]
    assert np.abs(b).max() > 0
    assert helpers.rel_l2(a, b) < 1e-12
    np.testing.assert_array_equal(a != 0.0, b != 0.0)
    gpu.close(); cpu.close()


def test_source_free_update_scales_exactly_and_fused_equals_separate(monkeypatch):
    """Without seed and bunch the update is linear with fixed coefficients: scaling the state by a power of two scales
    the result by exactly that power, bit for bit, at any size.  And the rim path equals the separate kernels."""
    p = _params(seed_enabled=0)
    p.power.enabled, p.screens.enabled = 0, 0
    a_n = bench.synthetic_potential(p)
    out = {}
    for name, scale, env in (("one", 1.0, None), ("eighth", 0.125, None), ("separate", 1.0, "MITHRA_NO_FUSE")):
        if env:
            monkeypatch.setenv(env, "1")
        s = abi.GpuSolver(p)
        s.upload_fields(an=a_n * scale, anm1=a_n * (0.999 * scale))
        s.upload_particles(np.zeros((0, 11)))
        s.step(3)
        out[name] = s.download_fields(("an",))["an"]
        s.close()
        if env:
            monkeypatch.delenv(env)
    assert np.abs(out["one"]).max() > 0
    np.testing.assert_array_equal(out["eighth"], out["one"] * 0.125)
    np.testing.assert_array_equal(out["separate"], out["one"])


def _bunch(p, workload):
    wl = bench.WORKLOADS[workload]
    return bench.synthetic_bunch(p, wl["particles"], sigma_t=wl["sigma_t"], trunc_t=wl["trunc_t"], sigma_gb=wl["sigma_gb"])


def test_one_full_size_step_of_the_space_charge_unit_against_the_oracle():
    """configs[4]: FdTdSC (A + phi, rho deposit) at the weak-scaling unit's size, whole mesh against the oracle."""
    p = _params("sc-weak")
    assert p.space_charge
    bunch = _bunch(p, "sc-weak")
    a_n = bench.synthetic_potential(p)
    f_n = np.ascontiguousarray(a_n.reshape(-1, 3)[:, 1]) * 0.5           # a smooth phi of the same shape
    tb = bench.undulator_time(p)
    gpu, cpu = abi.GpuSolver(p), binding.Oracle(copy.copy(p))
    for s in (gpu, cpu):
        s.set_time(tb, tb, 0)
        s.upload_fields(an=a_n, anm1=a_n * 0.999, fn=f_n, fnm1=f_n * 0.998)
        s.upload_particles(bunch)
    gpu.fieldUpdate(); cpu.fieldUpdate()
    a, b = gpu.download_fields(("anp1", "fnp1")), cpu.download_fields(("anp1", "fnp1"))
    for k in ("anp1", "fnp1"):
        assert np.abs(b[k]).max() > 0
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)         # no seed: bit-identical potentials everywhere
    del a, b
    np.testing.assert_array_equal(gpu.push_cells(), cpu.push_cells())
    gpu.bunchUpdate(); cpu.bunchUpdate()
    pg, pc = gpu.download_particles(), cpu.download_particles()
    np.testing.assert_array_equal(pg[:, 4:7], pc[:, 4:7])
    assert helpers.rel_l2(pg[:, 1:4], pc[:, 1:4]) < 1e-12
    assert helpers.rel_l2(pg[:, 7:10], pc[:, 7:10]) < 1e-12
    gpu.upload_particles(pc)
    for s in (gpu, cpu):
        s.fieldShift(); s.currentReset()
    np.testing.assert_array_equal(gpu.deposit_cells(), cpu.deposit_cells())
    gpu.currentUpdate(); cpu.currentUpdate()
    a, b = gpu.download_fields(("anp1", "fnp1")), cpu.download_fields(("anp1", "fnp1"))
    for k in ("anp1", "fnp1"):                                         # J and rho
        assert np.abs(b[k]).max() > 0
        assert helpers.rel_l2(a[k], b[k]) < 1e-12, k
        np.testing.assert_array_equal(a[k] != 0.0, b[k] != 0.0, err_msg=k)
    gpu.close(); cpu.close()


def test_one_full_size_step_of_fel_lcls_against_the_oracle_on_windows():
    """configs[3], the target configuration of north_star: 346.8 M nodes and 8.4 M macro-particles on one GPU."""
    from mithra_b200 import slabs
    p = _params("fel-lcls")
    assert p.N2 == 33335 and p.n_update_bunch == 1
    n = bench.WORKLOADS["fel-lcls"]["particles"]
    bunch = _bunch(p, "fel-lcls")
    a_n = bench.synthetic_potential(p)
    a_nm1 = a_n * 0.999
    tb = bench.undulator_time(p)
    gpu = abi.GpuSolver(p)
    gpu.set_time(tb, tb, 0)
    gpu.upload_fields(an=a_n, anm1=a_nm1)
    gpu.upload_particles(bunch)
    cells0 = gpu.push_cells()
    gpu.fieldUpdate()
    ap_g = gpu.download_fields(("anp1",))["anp1"]
    gpu.bunchUpdate()
    pg = gpu.download_particles()
    np.testing.assert_array_equal(pg[:, 4:7], bunch[:, 1:4])          # rnm = start-of-step position
    gpu.fieldShift(); gpu.currentReset()
    dep_g = gpu.deposit_cells()
    gpu.currentUpdate()
    j_g = gpu.download_fields(("anp1",))["anp1"]
    gpu.close()

    size = 256                                                       # slabs of ~132 planes
    plane = p.N0 * p.N1
    checked = 0
    for rank in (0, size // 2, size - 1):
        q = slabs.slab_params(p, rank, size)
        q.max_particles = n
        npl, k0 = q.np, q.k0
        cpu = binding.Oracle(q)
        cpu.set_time(tb, tb, 0)
        cpu.upload_fields(an=slabs.scatter_field(p, a_n, 3, rank, size), anm1=slabs.scatter_field(p, a_nm1, 3, rank, size))
        own = np.flatnonzero((bunch[:, 3] >= q.zp[0]) & (bunch[:, 3] < q.zp[1]))
        assert own.size > 1000 or rank in (0, size - 1)                # the synthetic bunch fills the middle 80 % of z
        cpu.upload_particles(bunch[own])
        # potentials: every plane the slab updates itself (on the end slabs that includes the mesh's z face)
        cpu.fieldUpdate()
        ap_c = cpu.download_fields(("anp1",))["anp1"].reshape(npl, -1)
        lo, hi = (0 if rank == 0 else 1), (npl if rank == size - 1 else npl - 1)
        win = ap_g.reshape(p.N2, -1)[k0 + lo:k0 + hi]
        assert np.abs(win).max() > 0
        np.testing.assert_array_equal(win, ap_c[lo:hi], err_msg="A+ of slab %d" % rank)
        if own.size == 0:
            cpu.close()
            continue
        # particles at least two cells inside the window see only E/B the slab computed itself
        zin = (bunch[own, 3] >= q.zp[0] + 2 * p.dz) & (bunch[own, 3] < q.zp[1] - 3 * p.dz)
        ref_cells = cpu.push_cells()
        loc = cells0[own] - k0 * plane                                 # whole-mesh node number -> slab numbering
        np.testing.assert_array_equal(loc[zin], ref_cells[zin])
        cpu.bunchUpdate()
        pc = cpu.download_particles()
        assert helpers.rel_l2(pg[own][zin][:, 1:4], pc[zin][:, 1:4]) < 1e-12
        assert helpers.rel_l2(pg[own][zin][:, 7:10], pc[zin][:, 7:10]) < 1e-12
        # deposit from identical particles (the GPU's), inner planes of the window
        cpu.upload_particles(pg[own])
        cpu.fieldShift(); cpu.currentReset()
        dc = cpu.deposit_cells()
        dg = dep_g[own].copy()
        np.testing.assert_array_equal(dg, dc)                           # global cell indices (ip jp kp im jm km)
        cpu.currentUpdate()
        j_c = cpu.download_fields(("anp1",))["anp1"].reshape(npl, -1)[3:npl - 3]
        j_w = j_g.reshape(p.N2, -1)[k0 + 3:k0 + npl - 3]
        assert np.abs(j_c).max() > 0
        assert helpers.rel_l2(j_w, j_c) < 1e-12
        np.testing.assert_array_equal(j_w != 0.0, j_c != 0.0)
        checked += int(zin.sum())
        cpu.close()
    assert checked > 10000
