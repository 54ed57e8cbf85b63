"""the bunch of Solver::initialize() on the device -- Bunch::initializeEllipsoid (classes.cpp:104-298),
the Lorentz boost and the ballistic back-projection of Solver::lorentzBoostBunch (solver.cpp:294-346) and the split over
the slabs (Solver::distributeParticles) as CUDA kernels (kernels_init.cuh) behind mithra_gpu_bunch_*.

Formulas and operation order are the reference's; the libm is not (CUDA's log / cos / sin against glibc's: last ulp), so the
device list equals the list the UNMODIFIED reference generated (p0 of tests/golden/*.npz) to a few ulp of the bunch size --
asserted as 1e-13 of each column's scale -- and the head particle's z, hence the time origin dt_, to 1e-14 relative.
(The host path, tests/test_host.py, stays bit-identical to the reference and is what small bunches use.)"""
import os
import subprocess

import numpy as np
import pytest

from mithra_b200 import meta as mmeta
from tests import helpers

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "mithra_b200", "host", "mithra_b200")


def _job(name):
    return os.path.join(ROOT, "tests", "jobs", name + ".job")


@pytest.mark.parametrize("job", ["micro-nsfd", "micro-sc", "micro-lcls", "micro-optical", "micro-trap"])
def test_device_generated_bunch_equals_the_references(job, tmp_path):
    """Uniform profile with Gaussian tapers, groups of four with a bunching factor, static and optical undulators (the
    bunching wavelength differs), gamma 30 ... 13089: the whole initialize() with the bunch on the device."""
    pre = str(tmp_path / "h")
    env = dict(os.environ, MITHRA_DEVICE_BUNCH="1")
    out = subprocess.check_output([EXE, _job(job), "--dump-params", pre], cwd=str(tmp_path), env=env).decode()
    assert "generated on the device" in out
    rec = mmeta.read_records(pre + ".meta.bin")
    meta, g = helpers.load_golden(job)
    got, want = rec["particles"].reshape(-1, 11), g["p0"]
    assert got.shape == want.shape                                      # same candidates accepted, same order
    scale = np.abs(want).max(axis=0) + 1e-300
    assert np.all(np.abs(got - want) <= 1e-13 * scale), (np.abs(got - want) / scale).max(axis=0)
    np.testing.assert_array_equal(got[:, [0, 4, 5, 6, 10]], want[:, [0, 4, 5, 6, 10]])      # q, rnm = 0, e = 0: no libm involved
    np.testing.assert_allclose(float(rec["dtShift"][0]), float(meta["dtShift"][0]), rtol=1e-14)
    for k in ("gamma", "beta", "dt", "dz", "zmin", "zmax"):            # nothing else of initialize() depends on the bunch
        np.testing.assert_array_equal(np.asarray(rec[k]), np.asarray(meta[k]), err_msg=k)


@pytest.mark.parametrize("gpus", [1, 3])
def test_job_with_a_device_generated_bunch_writes_the_references_power(gpus, tmp_path):
    """End to end: generate, boost, back-project and distribute on the device (1 and 3 slabs), 100 field steps, the power file
    against the unmodified reference's own rows."""
    meta, g = helpers.load_golden("micro-nsfd")
    env = dict(os.environ, MITHRA_DEVICE_BUNCH="1")
    out = subprocess.check_output([EXE, _job("micro-nsfd"), "--steps", "100", "--gpus", str(gpus)], cwd=str(tmp_path), env=env).decode()
    assert "generated on the device" in out
    got = np.loadtxt(tmp_path / "power-sampling" / "power-micro-0.txt")
    want = g["power"][:, 0]
    assert got.shape == (100, 2)
    np.testing.assert_allclose(got[:, 1], want, rtol=1e-7, atol=1e-10 * np.abs(want).max())


def test_eight_million_particles_in_milliseconds():
    """FEL-LCLS's bunch (8,388,608 macro-particles requested; uniform profile, bunching factor 0.001): generated, boosted
    and back-projected on the device in well under the 3 s the host path takes."""
    import ctypes as C
    import time
    from mithra_b200 import abi
    lib = abi.load()

    class E(C.Structure):
        _fields_ = [("number_of_particles", C.c_uint), ("index_offset", C.c_uint), ("cloud_charge", C.c_double), ("initial_gamma", C.c_double),
                    ("beta_vector", C.c_double * 3), ("position", C.c_double * 3), ("sigma_position", C.c_double * 3),
                    ("sigma_gamma_beta", C.c_double * 3), ("tran_trun", C.c_double), ("long_trun", C.c_double), ("lambda_", C.c_double),
                    ("bunching_factor", C.c_double), ("bunching_phase", C.c_double), ("distribution", C.c_int), ("device", C.c_int)]
    e = E(8388608, 0, 1.25e8, 13089.0, (0.0, 0.0, 1.0), (0.0, 0.0, 0.0), (30.0, 30.0, 0.4), (0.007, 0.007, 13.089), 180.0, 0.43,
          1.9e-6, 0.001, 0.0, 0, 0)
    lib.mithra_gpu_bunch_generate.argtypes = [C.POINTER(E), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.mithra_gpu_bunch_boost.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_double)]
    lib.mithra_gpu_bunch_backproject.argtypes = [C.c_void_p, C.c_double, C.c_double]
    lib.mithra_gpu_bunch_download.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_size_t)]
    lib.mithra_gpu_bunch_destroy.argtypes = [C.c_void_p]
    lib.mithra_gpu_bunch_destroy.restype = None
    times = []
    for rep in range(3):
        b, n, zmax = C.c_void_p(), C.c_size_t(), C.c_double()
        t0 = time.perf_counter()
        assert lib.mithra_gpu_bunch_generate(C.byref(e), C.byref(b), C.byref(n)) == 0, lib.mithra_gpu_last_error()
        assert lib.mithra_gpu_bunch_boost(b, 4903.6, 0.99999998, C.byref(zmax)) == 0
        assert lib.mithra_gpu_bunch_backproject(b, zmax.value, 0.99999998) == 0
        times.append(time.perf_counter() - t0)
        assert 8.3e6 < n.value < 8.6e6 and n.value % 4 == 0
        if rep == 2:
            a = np.empty((n.value, 11))
            assert lib.mithra_gpu_bunch_download(b, a.ctypes.data_as(C.POINTER(C.c_double)), n.value, C.byref(n)) == 0
            assert np.isfinite(a).all() and 29.0 < a[:, 1].std() < 40.0          # sigma_x = 30, widened by the back-projection
            assert a[:, 3].max() <= zmax.value * (1 + 1e-12)
        lib.mithra_gpu_bunch_destroy(b)
    print("generate + boost + back-project of %d particles: %.1f ms (first call %.1f ms)" % (n.value, 1e3 * min(times), 1e3 * times[0]))
    assert min(times) < 0.05
