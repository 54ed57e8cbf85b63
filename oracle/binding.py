"""ctypes binding of oracle/liboracle.so (the C restatement of the reference algorithm) and helpers to read the
state dumps of oracle/_ref/ref_dump.  TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg."""
import ctypes as C
import os
import subprocess

import numpy as np

from mithra_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_DUMP = os.path.join(HERE, "_ref", "ref_dump")
REF_MAIN = os.path.join(HERE, "_ref", "mithra_ref")

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        build()
    lib = C.CDLL(LIB)
    vp, dp = C.c_void_p, C.POINTER(C.c_double)
    lib.oracle_create.argtypes = [C.POINTER(abi.Params)]
    lib.oracle_create.restype = vp
    lib.oracle_destroy.argtypes = [vp]
    for n in ("anp1", "an", "anm1", "fnp1", "fn", "fnm1", "particles"):
        f = getattr(lib, "oracle_" + n)
        f.argtypes = [vp]
        f.restype = dp
    lib.oracle_en.argtypes = [vp]
    lib.oracle_en.restype = C.POINTER(C.c_float)
    lib.oracle_bn.argtypes = [vp]
    lib.oracle_bn.restype = C.POINTER(C.c_float)
    lib.oracle_pic.argtypes = [vp]
    lib.oracle_pic.restype = C.POINTER(C.c_ubyte)
    lib.oracle_set_particles.argtypes = [vp, dp, C.c_size_t]
    lib.oracle_num_particles.argtypes = [vp]
    lib.oracle_num_particles.restype = C.c_size_t
    lib.oracle_set_time.argtypes = [vp, C.c_double, C.c_double, C.c_uint]
    lib.oracle_time.argtypes = [vp]
    lib.oracle_time.restype = C.c_double
    lib.oracle_time_bunch.argtypes = [vp]
    lib.oracle_time_bunch.restype = C.c_double
    for n in ("field_update", "bunch_update", "screen_profile", "power_sample", "field_shift", "current_reset",
              "current_update", "advance_time", "seed_initial"):
        getattr(lib, "oracle_" + n).argtypes = [vp]
    lib.oracle_field_evaluate.argtypes = [vp, C.c_long]
    lib.oracle_field_sample.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.oracle_step.argtypes = [vp, C.c_int]
    lib.oracle_push_cells.argtypes = [vp, C.POINTER(C.c_long)]
    lib.oracle_deposit_cells.argtypes = [vp, C.POINTER(C.c_int)]
    lib.oracle_power_visualize.argtypes = [vp]
    lib.oracle_power_map.argtypes = [vp]
    lib.oracle_power_map.restype = dp
    lib.oracle_power_rows.argtypes = [vp]
    lib.oracle_power_rows.restype = C.c_size_t
    lib.oracle_power_data.argtypes = [vp]
    lib.oracle_power_data.restype = dp
    lib.oracle_screen_count.argtypes = [vp, C.c_int]
    lib.oracle_screen_count.restype = C.c_size_t
    lib.oracle_screen_data.argtypes = [vp, C.c_int]
    lib.oracle_screen_data.restype = dp
    lib.oracle_seed_fields.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_double, dp]
    _lib = lib
    return lib


class Oracle:
    """CPU oracle with the same method names as mithra_b200.abi.GpuSolver."""

    def __init__(self, params):
        self.lib = load()
        self.params = params
        self.o = self.lib.oracle_create(C.byref(params))
        self.nodes = params.N0 * params.N1 * params.np
        self.sc = bool(params.space_charge)

    def close(self):
        if self.o:
            self.lib.oracle_destroy(self.o)
            self.o = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _view(self, name, n, dtype=np.float64):
        ptr = getattr(self.lib, "oracle_" + name)(self.o)
        return np.ctypeslib.as_array(ptr, shape=(n,))

    def arr(self, name):
        n = self.nodes * (3 if name[0] in "aeb" else 1)
        return self._view(name, n)

    def upload_fields(self, an=None, anm1=None, jn=None, fn=None, fnm1=None, rho=None):
        for name, src in (("an", an), ("anm1", anm1), ("anp1", jn), ("fn", fn), ("fnm1", fnm1), ("fnp1", rho)):
            if src is not None:
                self.arr(name)[:] = np.asarray(src, dtype=np.float64).ravel()

    def download_fields(self, which=("anp1", "an", "anm1")):
        return {n: self.arr(n).copy() for n in which}

    def download_eb(self):
        return self.arr("en").copy(), self.arr("bn").copy(), self._view("pic", self.nodes).copy()

    def upload_particles(self, aos11):
        a = np.ascontiguousarray(aos11, dtype=np.float64).reshape(-1, 11)
        self.lib.oracle_set_particles(self.o, a.ctypes.data_as(C.POINTER(C.c_double)), a.shape[0])

    def download_particles(self):
        n = self.lib.oracle_num_particles(self.o)
        if n == 0:
            return np.zeros((0, 11))
        return np.ctypeslib.as_array(self.lib.oracle_particles(self.o), shape=(n, 11)).copy()

    def set_time(self, time, time_bunch, n_time):
        self.lib.oracle_set_time(self.o, time, time_bunch, n_time)

    def fieldUpdate(self):
        self.lib.oracle_field_update(self.o)

    def bunchUpdate(self):
        self.lib.oracle_bunch_update(self.o)

    def screenProfile(self):
        self.lib.oracle_screen_profile(self.o)

    def powerSample(self):
        self.lib.oracle_power_sample(self.o)

    def powerVisualize(self):
        self.lib.oracle_power_visualize(self.o)

    def field_sample(self, pos):
        """FdTd::fieldSample at the points pos[n][3] (moving frame): et, bt, at per point as an (n, 9) array."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        out = np.zeros((len(pos), 9))
        for t in range(len(pos)):
            self.lib.oracle_field_sample(self.o, pos[t].ctypes.data_as(C.POINTER(C.c_double)), out[t].ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def fetch_power_map(self):
        ptr = self.lib.oracle_power_map(self.o)
        if not ptr:
            return None
        return np.ctypeslib.as_array(ptr, shape=(self.params.N0 * self.params.N1,)).copy()

    def fieldShift(self):
        self.lib.oracle_field_shift(self.o)

    def currentReset(self):
        self.lib.oracle_current_reset(self.o)

    def currentUpdate(self):
        self.lib.oracle_current_update(self.o)

    def currentCommunicate(self):
        pass

    def advanceTime(self):
        self.lib.oracle_advance_time(self.o)

    def seedInitial(self):
        self.lib.oracle_seed_initial(self.o)

    def step(self, nsteps=1):
        self.lib.oracle_step(self.o, nsteps)

    def push_cells(self):
        n = self.lib.oracle_num_particles(self.o)
        out = np.empty(n, dtype=np.int64)
        self.lib.oracle_push_cells(self.o, out.ctypes.data_as(C.POINTER(C.c_long)))
        return out

    def deposit_cells(self):
        n = self.lib.oracle_num_particles(self.o)
        out = np.empty((n, 6), dtype=np.int32)
        self.lib.oracle_deposit_cells(self.o, out.ctypes.data_as(C.POINTER(C.c_int)))
        return out

    def fetch_power(self):
        n = self.lib.oracle_power_rows(self.o)
        w = max(1, self.params.power.N * self.params.power.Nl)
        if n == 0:
            return np.zeros((0, w))
        return np.ctypeslib.as_array(self.lib.oracle_power_data(self.o), shape=(n, w)).copy()

    def fetch_screen(self, s):
        n = self.lib.oracle_screen_count(self.o, s)
        if n == 0:
            return np.zeros((0, 6))
        return np.ctypeslib.as_array(self.lib.oracle_screen_data(self.o, s), shape=(n, 6)).copy()


from mithra_b200.meta import read_records, write_records, params_from_meta  # noqa: E402,F401


def have_reference():
    return os.path.exists(REF_DUMP)


def run_ref_dump(job, prefix, nsteps, full_at=(), phases_at=None, cwd=None):
    cmd = [REF_DUMP, job, prefix, str(nsteps), "--quiet"]
    if full_at:
        cmd += ["--full-at", ",".join(str(s) for s in full_at)]
    if phases_at is not None:
        cmd += ["--phases-at", str(phases_at)]
    subprocess.check_call(cmd, cwd=cwd, stdout=subprocess.DEVNULL)


