"""Parity at BASELINE.json's full size (configs[1], FEL-SEEDED: 85 x 85 x 8252 nodes, 4,194,304 macro-particles): one
whole field step of the CUDA path against the CPU oracle on the same synthetic state bench.py uses, and properties
that do not depend on the size -- exact scaling of the source-free field update, fused against separate kernels."""
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


def _params(**over):
    p = mmeta.params_from_meta(dict(np.load(os.path.join(ROOT, "bench", "fel-seeded.meta.npz"))))
    p.max_particles = 4194304 + 1024
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
