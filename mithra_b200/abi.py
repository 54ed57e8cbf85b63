"""ctypes mirror of include/mithra_gpu.h (the C ABI of libmithra_gpu.so).

Python is only the test / bench harness of this project: the product is the C ABI and the C++ host classes
above it (mithra_b200/host).  Nothing here computes; there is no CPU path.  Loading fails loudly when the
shared library has not been built (python -c "import __graft_entry__ as g; g.build()").
"""
import ctypes as C
import os

import numpy as np

ABI_VERSION = 4
MAX_UNDULATORS, MAX_EXTFIELDS = 16, 8
MAX_POWER_PLANES, MAX_POWER_LAMBDAS, MAX_SCREENS = 256, 64, 64
NPHASES = 8
PHASE_NAMES = ("stencil", "boundary", "clear", "eval_eb", "push", "deposit", "power", "screens")

ROOT = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(ROOT, "libmithra_gpu.so")


class Signal(C.Structure):
    _fields_ = [("type", C.c_int), ("t0", C.c_double), ("s", C.c_double), ("f0", C.c_double), ("cep", C.c_double),
                ("nR", C.c_int), ("sigma_inv_g", C.c_double * 2)]


class Beam(C.Structure):
    _fields_ = [("seed_type", C.c_int), ("position", C.c_double * 3), ("direction", C.c_double * 3),
                ("polarization", C.c_double * 3), ("amplitude", C.c_double), ("radius", C.c_double * 2),
                ("l", C.c_double), ("zR", C.c_double * 2), ("order", C.c_int * 2), ("signal", Signal)]


class Undulator(C.Structure):
    _fields_ = [("type", C.c_int), ("k", C.c_double), ("lu", C.c_double), ("rb", C.c_double), ("theta", C.c_double),
                ("length", C.c_double), ("dist", C.c_double), ("beam", Beam)]


class Power(C.Structure):
    _fields_ = [("enabled", C.c_int), ("N", C.c_int), ("z", C.c_double * MAX_POWER_PLANES), ("Nl", C.c_int),
                ("w", C.c_double * MAX_POWER_LAMBDAS), ("Nf", C.c_int), ("pc", C.c_double)]


class PowerMap(C.Structure):
    _fields_ = [("enabled", C.c_int), ("Nf", C.c_int), ("z", C.c_double), ("w", C.c_double), ("pc", C.c_double)]


class Screens(C.Structure):
    _fields_ = [("enabled", C.c_int), ("N", C.c_int), ("pos", C.c_double * MAX_SCREENS)]


class Params(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int),
        ("N0", C.c_int), ("N1", C.c_int), ("N2", C.c_int), ("np", C.c_int), ("k0", C.c_int),
        ("rank", C.c_int), ("size", C.c_int),
        ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double), ("dt", C.c_double),
        ("xmin", C.c_double), ("xmax", C.c_double), ("ymin", C.c_double), ("ymax", C.c_double),
        ("zmin", C.c_double), ("zmax", C.c_double),
        ("zp", C.c_double * 2), ("Lz", C.c_double),
        ("solver", C.c_int), ("space_charge", C.c_int), ("truncation_order", C.c_int),
        ("a", C.c_double * 6), ("alpha", C.c_double), ("beta_nsfd", C.c_double),
        ("bB", C.c_double * 5), ("cB", C.c_double * 5), ("dB", C.c_double * 5),
        ("eE", C.c_double * 5), ("fE", C.c_double * 5), ("gE", C.c_double * 5), ("hC", C.c_double * 17),
        ("c0", C.c_double), ("gamma", C.c_double), ("beta", C.c_double), ("dt_shift", C.c_double),
        ("dt_bunch", C.c_double), ("n_update_bunch", C.c_int),
        ("r1", C.c_double), ("r2", C.c_double), ("dtb", C.c_double),
        ("n_undulators", C.c_int), ("undulator", Undulator * MAX_UNDULATORS),
        ("n_ext_fields", C.c_int), ("ext_field", Beam * MAX_EXTFIELDS),
        ("seed_enabled", C.c_int), ("seed", Beam),
        ("power", Power), ("screens", Screens),
        ("max_particles", C.c_size_t), ("max_screen_records", C.c_size_t), ("device", C.c_int),
        ("sort_interval", C.c_int),
        ("power_map", PowerMap),
    ]


class Counters(C.Structure):
    _fields_ = [("field_steps", C.c_ulonglong), ("cell_updates", C.c_ulonglong),
                ("particle_pushes", C.c_ulonglong), ("kernel_launches", C.c_ulonglong)]


# every symbol include/mithra_gpu.h declares (checked by tests/test_abi.py without touching a GPU)
SYMBOLS = (
    "mithra_gpu_last_error", "mithra_gpu_abi_version", "mithra_gpu_device_count", "mithra_gpu_create",
    "mithra_gpu_destroy", "mithra_gpu_seed_initial", "mithra_gpu_upload_fields", "mithra_gpu_download_fields", "mithra_gpu_download_eb",
    "mithra_gpu_upload_particles", "mithra_gpu_download_particles", "mithra_gpu_num_particles",
    "mithra_gpu_particle_cells", "mithra_gpu_sort_particles",
    "mithra_gpu_set_time", "mithra_gpu_get_time", "mithra_gpu_field_update", "mithra_gpu_bunch_update",
    "mithra_gpu_screen_profile", "mithra_gpu_power_sample", "mithra_gpu_field_shift", "mithra_gpu_current_reset",
    "mithra_gpu_current_update", "mithra_gpu_current_communicate", "mithra_gpu_advance_time", "mithra_gpu_step",
    "mithra_gpu_step_timed", "mithra_gpu_synchronize", "mithra_gpu_fetch_power", "mithra_gpu_fetch_screen",
    "mithra_gpu_counters", "mithra_gpu_step_profiled", "mithra_gpu_ipc_export", "mithra_gpu_ipc_connect",
    "mithra_gpu_migrate_begin", "mithra_gpu_migrate_end", "mithra_gpu_selftest_divide",
    "mithra_gpu_power_visualize", "mithra_gpu_fetch_power_map", "mithra_gpu_bunch_moments", "mithra_gpu_field_sample",
    "mithra_gpu_field_nodes",
    "mithra_gpu_bunch_generate", "mithra_gpu_bunch_boost", "mithra_gpu_bunch_backproject", "mithra_gpu_bunch_download",
    "mithra_gpu_bunch_destroy", "mithra_gpu_upload_particles_device",
)

_lib = None


def load():
    """Load libmithra_gpu.so; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    global LIB_PATH
    if os.environ.get("MITHRA_GPU_LIB"):                 # tuning experiments: another build of the same library
        LIB_PATH = os.environ["MITHRA_GPU_LIB"]
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: build it first (python -c 'import __graft_entry__ as g; g.build()'). "
                           "There is no CPU fallback." % LIB_PATH)
    # see preload_kernels() in engine.cu; harmless if CUDA is already initialised
    os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
    lib = C.CDLL(LIB_PATH)
    dp, fp, vp = C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_void_p
    lib.mithra_gpu_last_error.restype = C.c_char_p
    lib.mithra_gpu_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    lib.mithra_gpu_destroy.argtypes = [vp]
    lib.mithra_gpu_destroy.restype = None
    lib.mithra_gpu_upload_fields.argtypes = [vp] + [dp] * 6
    lib.mithra_gpu_download_fields.argtypes = [vp] + [dp] * 6
    lib.mithra_gpu_download_eb.argtypes = [vp, fp, fp, C.POINTER(C.c_ubyte)]
    lib.mithra_gpu_upload_particles.argtypes = [vp, dp, C.c_size_t]
    lib.mithra_gpu_download_particles.argtypes = [vp, dp, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.mithra_gpu_num_particles.argtypes = [vp, C.POINTER(C.c_size_t)]
    lib.mithra_gpu_particle_cells.argtypes = [vp, C.POINTER(C.c_long), C.POINTER(C.c_int), C.c_size_t]
    lib.mithra_gpu_set_time.argtypes = [vp, C.c_double, C.c_double, C.c_uint]
    lib.mithra_gpu_get_time.argtypes = [vp, dp, dp, C.POINTER(C.c_uint)]
    for name in ("field_update", "bunch_update", "screen_profile", "power_sample", "field_shift", "current_reset",
                 "current_update", "current_communicate", "advance_time", "synchronize", "seed_initial", "sort_particles",
                 "migrate_begin", "migrate_end", "power_visualize"):
        getattr(lib, "mithra_gpu_" + name).argtypes = [vp]
    lib.mithra_gpu_step.argtypes = [vp, C.c_int]
    lib.mithra_gpu_step_timed.argtypes = [vp, C.c_int, fp]
    lib.mithra_gpu_step_profiled.argtypes = [vp, C.c_int, fp]
    lib.mithra_gpu_fetch_power.argtypes = [vp, dp, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.mithra_gpu_fetch_screen.argtypes = [vp, C.c_int, dp, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.mithra_gpu_counters.argtypes = [vp, C.POINTER(Counters)]
    lib.mithra_gpu_ipc_export.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.mithra_gpu_ipc_connect.argtypes = [vp, vp, vp]
    lib.mithra_gpu_fetch_power_map.argtypes = [vp, dp, C.c_size_t, C.POINTER(C.c_int)]
    lib.mithra_gpu_bunch_moments.argtypes = [vp, dp]
    lib.mithra_gpu_field_sample.argtypes = [vp, dp, C.c_size_t, dp, C.POINTER(C.c_ubyte)]
    lib.mithra_gpu_field_nodes.argtypes = [vp, C.POINTER(C.c_int), C.c_size_t, dp, C.POINTER(C.c_ubyte)]
    lib.mithra_gpu_selftest_divide.argtypes = [dp, C.c_size_t, C.c_double, C.POINTER(C.c_ulonglong)]
    _lib = lib
    return lib


def _dptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


class GpuSolver:
    """Thin object wrapper over the C ABI; method names follow the reference's Solver/FdTd methods."""

    def __init__(self, params):
        self.lib = load()
        self.params = params
        h = C.c_void_p()
        self._check(self.lib.mithra_gpu_create(C.byref(params), C.byref(h)))
        self.h = h
        self.nodes = params.N0 * params.N1 * params.np

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.mithra_gpu_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.mithra_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state ------------------------------------------------------------------------------------------
    def upload_fields(self, an=None, anm1=None, jn=None, fn=None, fnm1=None, rho=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (an, anm1, jn, fn, fnm1, rho)]
        self._check(self.lib.mithra_gpu_upload_fields(self.h, *[_dptr(a) for a in arrs]))

    def download_fields(self, which=("anp1", "an", "anm1"), out=None):
        """`out`: optional dict name -> preallocated float64 array (e.g. pinned host memory) to receive the level."""
        names = ("anp1", "an", "anm1", "fnp1", "fn", "fnm1")
        given, out = out or {}, {}
        for n in which:
            size = self.nodes * (3 if n.startswith("a") else 1)
            out[n] = given[n] if n in given else np.empty(size, dtype=np.float64)
            assert out[n].dtype == np.float64 and out[n].size == size and out[n].flags["C_CONTIGUOUS"]
        self._check(self.lib.mithra_gpu_download_fields(self.h, *[_dptr(out.get(n)) for n in names]))
        return out

    def download_eb(self):
        en = np.empty(self.nodes * 3, dtype=np.float32)
        bn = np.empty(self.nodes * 3, dtype=np.float32)
        mask = np.empty(self.nodes, dtype=np.uint8)
        self._check(self.lib.mithra_gpu_download_eb(self.h, en.ctypes.data_as(C.POINTER(C.c_float)),
                                                    bn.ctypes.data_as(C.POINTER(C.c_float)),
                                                    mask.ctypes.data_as(C.POINTER(C.c_ubyte))))
        return en, bn, mask

    def upload_particles(self, aos11):
        a = np.ascontiguousarray(aos11, dtype=np.float64).reshape(-1, 11)
        self._check(self.lib.mithra_gpu_upload_particles(self.h, _dptr(a), a.shape[0]))

    def download_particles(self, out=None):
        """`out`: optional preallocated float64 array of at least n x 11 (e.g. pinned host memory)."""
        n = C.c_size_t()
        self._check(self.lib.mithra_gpu_num_particles(self.h, C.byref(n)))
        if out is None:
            out = np.empty((n.value, 11), dtype=np.float64)
        else:
            assert out.dtype == np.float64 and out.size >= n.value * 11 and out.flags["C_CONTIGUOUS"]
            out = out.reshape(-1)[:n.value * 11].reshape(n.value, 11)
        self._check(self.lib.mithra_gpu_download_particles(self.h, _dptr(out), n.value, C.byref(n)))
        return out

    def num_particles(self):
        n = C.c_size_t()
        self._check(self.lib.mithra_gpu_num_particles(self.h, C.byref(n)))
        return n.value

    def push_cells(self):
        n = self.num_particles()
        out = np.empty(n, dtype=np.int64)
        self._check(self.lib.mithra_gpu_particle_cells(self.h, out.ctypes.data_as(C.POINTER(C.c_long)), None, n))
        return out

    def deposit_cells(self):
        n = self.num_particles()
        out = np.empty((n, 6), dtype=np.int32)
        self._check(self.lib.mithra_gpu_particle_cells(self.h, None, out.ctypes.data_as(C.POINTER(C.c_int)), n))
        return out

    def set_time(self, time, time_bunch, n_time):
        self._check(self.lib.mithra_gpu_set_time(self.h, time, time_bunch, n_time))

    def get_time(self):
        t, tb, n = C.c_double(), C.c_double(), C.c_uint()
        self._check(self.lib.mithra_gpu_get_time(self.h, C.byref(t), C.byref(tb), C.byref(n)))
        return t.value, tb.value, n.value

    # -- the reference's methods ------------------------------------------------------------------------
    def fieldUpdate(self):
        self._check(self.lib.mithra_gpu_field_update(self.h))

    def bunchUpdate(self):
        self._check(self.lib.mithra_gpu_bunch_update(self.h))

    def screenProfile(self):
        self._check(self.lib.mithra_gpu_screen_profile(self.h))

    def powerSample(self):
        self._check(self.lib.mithra_gpu_power_sample(self.h))

    def powerVisualize(self):
        self._check(self.lib.mithra_gpu_power_visualize(self.h))

    def bunch_moments(self):
        """The 13 raw sums of Solver::bunchSample: q, q r[3], q r^2[3], q gb[3], q gb^2[3]."""
        out = np.zeros(13)
        self._check(self.lib.mithra_gpu_bunch_moments(self.h, _dptr(out)))
        return out

    def field_sample(self, pos):
        """FdTd::fieldSample at the points pos[n][3] (moving frame): (et, bt, at per point as an (n, 9) array, mine[n])."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        out = np.zeros((len(pos), 9))
        mine = np.zeros(len(pos), dtype=np.uint8)
        self._check(self.lib.mithra_gpu_field_sample(self.h, _dptr(pos), len(pos), _dptr(out), mine.ctypes.data_as(C.POINTER(C.c_ubyte))))
        return out, mine

    def fetch_power_map(self):
        """pL[i*N1 + j] of the last powerVisualize call, or None when the plane lies in another slab."""
        out = np.zeros(self.params.N0 * self.params.N1)
        mine = C.c_int(0)
        self._check(self.lib.mithra_gpu_fetch_power_map(self.h, _dptr(out), out.size, C.byref(mine)))
        return out if mine.value else None

    def fieldShift(self):
        self._check(self.lib.mithra_gpu_field_shift(self.h))

    def currentReset(self):
        self._check(self.lib.mithra_gpu_current_reset(self.h))

    def currentUpdate(self):
        self._check(self.lib.mithra_gpu_current_update(self.h))

    def currentCommunicate(self):
        self._check(self.lib.mithra_gpu_current_communicate(self.h))

    def advanceTime(self):
        self._check(self.lib.mithra_gpu_advance_time(self.h))

    def sortParticles(self):
        self._check(self.lib.mithra_gpu_sort_particles(self.h))

    def seedInitial(self):
        self._check(self.lib.mithra_gpu_seed_initial(self.h))

    def step(self, nsteps=1):
        self._check(self.lib.mithra_gpu_step(self.h, nsteps))

    def step_timed(self, nsteps):
        ms = C.c_float()
        self._check(self.lib.mithra_gpu_step_timed(self.h, nsteps, C.byref(ms)))
        return ms.value

    def step_profiled(self, nsteps):
        ms = (C.c_float * NPHASES)()
        self._check(self.lib.mithra_gpu_step_profiled(self.h, nsteps, ms))
        return dict(zip(PHASE_NAMES, list(ms)))

    def synchronize(self):
        self._check(self.lib.mithra_gpu_synchronize(self.h))

    def fetch_power(self):
        n = C.c_size_t()
        self._check(self.lib.mithra_gpu_fetch_power(self.h, None, 0, C.byref(n)))
        w = max(1, self.params.power.N * self.params.power.Nl)
        out = np.empty((n.value, w), dtype=np.float64)
        if n.value:
            self._check(self.lib.mithra_gpu_fetch_power(self.h, _dptr(out), n.value, C.byref(n)))
        return out

    def fetch_screen(self, s):
        n = C.c_size_t()
        self._check(self.lib.mithra_gpu_fetch_screen(self.h, s, None, 0, C.byref(n)))
        out = np.empty((n.value, 6), dtype=np.float64)
        if n.value:
            self._check(self.lib.mithra_gpu_fetch_screen(self.h, s, _dptr(out), n.value, C.byref(n)))
        return out

    # -- z-slabs ----------------------------------------------------------------------------------------
    def migrateBegin(self):
        self._check(self.lib.mithra_gpu_migrate_begin(self.h))

    def migrateEnd(self):
        self._check(self.lib.mithra_gpu_migrate_end(self.h))

    def export_blob(self):
        n = C.c_size_t()
        self._check(self.lib.mithra_gpu_ipc_export(self.h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        self._check(self.lib.mithra_gpu_ipc_export(self.h, buf, n.value, C.byref(n)))
        return buf.raw

    def connect(self, blob_prev, blob_next):
        self._check(self.lib.mithra_gpu_ipc_connect(self.h, C.c_char_p(blob_prev), C.c_char_p(blob_next)))

    def connect_neighbours(self, dist, rank, world):
        """One process per GPU: exchange the blobs through torch.distributed and connect the ring neighbours."""
        blobs = [None] * world
        dist.all_gather_object(blobs, self.export_blob())
        self.connect(blobs[(rank - 1) % world], blobs[(rank + 1) % world])
        dist.barrier()

    def counters(self):
        c = Counters()
        self._check(self.lib.mithra_gpu_counters(self.h, C.byref(c)))
        return c


def selftest_divide(x, d):
    """Number of elements of x for which the device's constant-divisor division differs bitwise from x / d."""
    lib = load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    bad = C.c_ulonglong(0)
    if lib.mithra_gpu_selftest_divide(_dptr(x), x.size, float(d), C.byref(bad)):
        raise RuntimeError(lib.mithra_gpu_last_error().decode())
    return bad.value
