"""Parameter-block helpers of the test / bench harness: the record files written by the state dumper of the
unmodified reference (oracle/ref_dump.cpp) and the conversion of its meta record -- the reference's own
Solver::initialize() results -- into the C-ABI parameter block (include/mithra_gpu.h MithraGpuParams).
Nothing here computes physics."""
import numpy as np

from . import abi

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
    def beam(dst, key):
        o, sg = meta[key + "beam"], meta[key + "sig"]
        dst.seed_type = int(o[0])
        for c in range(3):
            dst.position[c], dst.direction[c], dst.polarization[c] = o[1 + c], o[4 + c], o[7 + c]
        dst.amplitude = o[10]
        dst.radius[0], dst.radius[1], dst.l, dst.zR[0], dst.zR[1] = o[11], o[12], o[13], o[14], o[15]
        dst.order[0], dst.order[1] = int(o[16]), int(o[17])
        dst.signal.type, dst.signal.t0, dst.signal.s, dst.signal.f0 = int(sg[0]), sg[1], sg[2], sg[3]
        dst.signal.nR, dst.signal.cep = int(sg[4]), sg[5]
        dst.signal.sigma_inv_g[0], dst.signal.sigma_inv_g[1] = sg[6], sg[7]

    p.n_undulators = int(g("nUndulators"))
    for u in range(p.n_undulators):
        s = meta["und%d.static" % u]
        U = p.undulator[u]
        U.k, U.lu, U.rb, U.length, U.dist, U.theta, U.type = s[0], s[1], s[2], s[3], s[4], s[5], int(s[6])
        beam(U.beam, "und%d." % u)
    p.n_ext_fields = int(g("nExtFields"))
    for u in range(p.n_ext_fields):
        beam(p.ext_field[u], "ext%d." % u)
    p.seed_enabled = 1 if abs(g("seedAmplitude")) > 1.0e-50 else 0      # fdtd.cpp:307
    beam(p.seed, "seed.")
    # every sub-group of FEL-OUTPUT is its own FreeElectronLaser entry (datainput.cpp:632-751); the C ABI carries
    # one power group and one screen group: the first of each
    pk = sorted(k for k in meta if k.startswith("power") and k.endswith(".N"))
    if pk:
        key = pk[0][:-1]
        w = p.power
        w.enabled, w.N, w.Nl, w.Nf, w.pc = 1, int(g(key + "N")), int(g(key + "Nl")), int(g(key + "Nf")), g(key + "pc")
        for i, v in enumerate(meta[key + "z"]):
            w.z[i] = v
        for i, v in enumerate(meta[key + "w"]):
            w.w[i] = v
    mk = sorted(k for k in meta if k.startswith("pmap") and k.endswith(".Nf"))
    if mk:
        key = mk[0][:-2]
        m = p.power_map
        m.enabled, m.Nf, m.z, m.w, m.pc = 1, int(g(key + "Nf")), g(key + "z"), g(key + "w"), g(key + "pc")
    sk = sorted(k for k in meta if k.startswith("screen") and k.endswith(".pos"))
    if sk:
        s = p.screens
        s.enabled, s.N = 1, len(meta[sk[0]])
        for i, v in enumerate(meta[sk[0]]):
            s.pos[i] = v
    p.max_particles = max_particles
    p.device = -1
    return p
