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
    lib.oracle_step.argtypes = [vp, C.c_int]
    lib.oracle_push_cells.argtypes = [vp, C.POINTER(C.c_long)]
    lib.oracle_deposit_cells.argtypes = [vp, C.POINTER(C.c_int)]
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


# ------------------------------------------------------------------------------------------------------
# ref_dump record files

_DT = {0: np.float64, 1: np.float32, 2: np.int32, 3: np.uint8}


def read_records(fn):
    out = {}
    with open(fn, "rb") as f:
        while True:
            h = f.read(48)
            if len(h) < 48:
                break
            name = h.split(b"\0")[0].decode()
            t = int(np.frombuffer(f.read(4), np.int32)[0])
            n = int(np.frombuffer(f.read(8), np.int64)[0])
            out[name] = np.frombuffer(f.read(n * np.dtype(_DT[t]).itemsize), _DT[t]).copy()
    return out


def write_records(fn, rec):
    code = {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.int32): 2, np.dtype(np.uint8): 3}
    with open(fn, "wb") as f:
        for name, a in rec.items():
            a = np.ascontiguousarray(a)
            f.write(name.encode().ljust(48, b"\0"))
            f.write(np.int32(code[a.dtype]).tobytes())
            f.write(np.int64(a.size).tobytes())
            f.write(a.tobytes())


def have_reference():
    return os.path.exists(REF_DUMP)


def run_ref_dump(job, prefix, nsteps, full_at=(), phases_at=None, cwd=None):
    cmd = [REF_DUMP, job, prefix, str(nsteps), "--quiet"]
    if full_at:
        cmd += ["--full-at", ",".join(str(s) for s in full_at)]
    if phases_at is not None:
        cmd += ["--phases-at", str(phases_at)]
    subprocess.check_call(cmd, cwd=cwd, stdout=subprocess.DEVNULL)


def params_from_meta(meta, max_particles=0):
    """Build the C-ABI parameter block from a ref_dump meta record (the reference's own initialize() results)."""
    g = lambda k: meta[k][0]
    p = abi.Params()
    p.abi_version = abi.ABI_VERSION
    p.N0, p.N1, p.N2, p.np, p.k0 = int(g("N0")), int(g("N1")), int(g("N2")), int(g("np")), int(g("k0"))
    p.rank, p.size = int(g("rank")), int(g("size"))
    p.dx, p.dy, p.dz, p.dt = g("dx"), g("dy"), g("dz"), g("dt")
    p.xmin, p.xmax, p.ymin, p.ymax, p.zmin, p.zmax = g("xmin"), g("xmax"), g("ymin"), g("ymax"), g("zmin"), g("zmax")
    p.zp[0], p.zp[1] = meta["zp"]
    p.Lz = g("Lz")
    p.solver, p.space_charge, p.truncation_order = int(g("solver")), int(g("spaceCharge")), int(g("truncationOrder"))
    for k in ("a", "bB", "cB", "dB", "eE", "fE", "gE", "hC"):
        for i, v in enumerate(meta[k]):
            getattr(p, k)[i] = v
    p.alpha, p.beta_nsfd = g("alpha"), g("betaNSFD")
    p.c0, p.gamma, p.beta, p.dt_shift = g("c0"), g("gamma"), g("beta"), g("dtShift")
    p.dt_bunch, p.n_update_bunch = g("dtBunch"), int(round(g("nUpdateBunch")))
    p.r1, p.r2, p.dtb = g("r1"), g("r2"), g("dtb")
    p.n_undulators = int(g("nUndulators"))
    for u in range(p.n_undulators):
        s = meta["und%d.static" % u]
        U = p.undulator[u]
        U.k, U.lu, U.rb, U.length, U.dist, U.theta, U.type = s[0], s[1], s[2], s[3], s[4], s[5], int(s[6])
        o = meta["und%d.optical" % u]
        U.beam.seed_type = int(s[7])
        for c in range(3):
            U.beam.position[c], U.beam.direction[c], U.beam.polarization[c] = o[c], o[3 + c], o[6 + c]
        U.beam.amplitude = o[9]
        U.beam.radius[0], U.beam.radius[1], U.beam.l, U.beam.zR[0], U.beam.zR[1] = o[11], o[12], o[13], o[14], o[15]
        sg = meta["und%d.signal" % u]
        U.beam.signal.type, U.beam.signal.t0, U.beam.signal.s, U.beam.signal.f0 = int(sg[0]), sg[1], sg[2], sg[3]
        U.beam.signal.nR, U.beam.signal.cep = int(sg[4]), sg[5]
    if "power0.N" in meta:
        w = p.power
        w.enabled, w.N, w.Nl, w.Nf, w.pc = 1, int(g("power0.N")), int(g("power0.Nl")), int(g("power0.Nf")), g("power0.pc")
        for i, v in enumerate(meta["power0.z"]):
            w.z[i] = v
        for i, v in enumerate(meta["power0.w"]):
            w.w[i] = v
    if "screen0.pos" in meta:
        s = p.screens
        s.enabled, s.N = 1, len(meta["screen0.pos"])
        for i, v in enumerate(meta["screen0.pos"]):
            s.pos[i] = v
    p.max_particles = max_particles
    p.device = -1
    return p
