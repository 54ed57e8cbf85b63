#!/usr/bin/env python
"""bench.py -- throughput of the B200-native MITHRA time-march on BASELINE.json's workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload fel-lcls|fel-seeded|sc-weak]
                    [--scaling weak|strong]

One "step" = one field time step of Solver::solve (solver.cpp:1300-1399): fieldUpdate (stencil + absorbing boundaries
+ TF/SF seed + E/B evaluation), nUpdateBunch Boris sub-pushes of every macro-particle, screens, radiated power,
fieldShift, currentReset, ZigZag deposit.  Workload at N=1: BASELINE.json configs[3], the configuration north_star's
target is quoted on -- FEL-LCLS (jobs/fel-lcls.job: 102 x 102 x 33,335 nodes, 8,388,608 macro-particles, 1 sub-push
per step, 46 GB resident on one B200); configs[1] FEL-SEEDED and configs[4] (fdtdSC unit) are --workload options.  The
mesh constants and update coefficients are the reference's own Solver::initialize() results for the job
(bench/<workload>.meta.npz), the bunch and the potentials are synthetic (Halton bunch inside the undulator, smooth
wave packet) as SURVEY.md section 8(d) specifies.

The JSON line (rank 0): metric = cell-updates/s of the whole job (median of the timed blocks of K steps; the blocks
are repeated until at least --min-seconds of device time have been measured), with the particle pushes/s of the same
timed region under "pushes"; "roofline" = the WHOLE field step against the HBM roofline (every algorithmic byte of
SURVEY 8(d) over ms_per_step), with the per-phase table and the dominant kernel beside it; "cpu_baseline" = the
unmodified reference (oracle/_ref/ref_dump --bench on the mini-MPI ranks of all host cores) on a bounded z-slice of the
same job; "e2e" = the same job through the C ABI with host buffers (pinned upload of potentials + bunch, per-step power
read-back, final download inside the timed region).  With N > 1 the line also carries "check": the N-slab run of a
reduced mesh against the 1-slab run of the same problem.  At N = 1 "host_binary" is one more leg: jobs/<workload>.job through
the C++ host (mithra_b200/host/mithra_b200 --steps K+W: the reference's Solver surface over the C ABI, its own initialize()),
step time from the host's own clock around the K steps -- the number for the path a user of the reference would run.

--impl reference times the reference's own CPU implementation alone with the same keys: on the SAME mesh and bunch
(jobs/<workload>.job) when K + W steps of it fit a few minutes of host time, else on the bounded z-slice.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: meta fixture (tools/make_bench_meta.py), sample job + z fraction for the CPU baseline, description, and the
    # synthetic bunch (count per GPU, transverse sigma / truncation in length-scale units, momentum spread) of the job
    "fel-seeded": dict(meta="bench/fel-seeded.meta.npz", job="jobs/fel-seeded.job", sample="jobs/fel-seeded-sample.job", frac=8, particles=4194304,
                       sigma_t=95.3, trunc_t=400.0, sigma_gb=0.0105,
                       desc="FEL-SEEDED 85x85x8252 nodes, 4194304 macro-particles, 3 sub-pushes/step, NSFD, seed TF/SF, 1 power plane, 7 screens"),
    # BASELINE.json configs[3]: the large-z X-ray FEL mesh on ONE GPU (46 GB resident); --gpus N stacks N of them
    "fel-lcls": dict(meta="bench/fel-lcls.meta.npz", job="jobs/fel-lcls.job", sample="jobs/fel-lcls-sample.job", frac=32, particles=8388608,
                     sigma_t=30.0, trunc_t=180.0, sigma_gb=0.007,
                     desc="FEL-LCLS 102x102x33335 nodes, 8388608 macro-particles, 1 sub-push/step, NSFD, 1 power plane"),
    # BASELINE.json configs[4]: FdTdSC (A + phi) weak-scaling unit, 102 x 102 x 4096 cells and 1 Mi particles per GPU
    "sc-weak": dict(meta="bench/sc-weak.meta.npz", job="jobs/sc-weak.job", sample="jobs/sc-weak-sample.job", frac=4, particles=1048576,
                    sigma_t=30.0, trunc_t=180.0, sigma_gb=0.007,
                    desc="fdtdSC weak-scaling unit 102x102x4098 nodes (A + phi), 1048576 macro-particles, 1 sub-push/step, NSFD, 1 power plane"),
}

BYTES_PER_CELL = {0: 96, 1: 128}      # read an, anm1, J + write anp1, 24 B each (+ 4 x 8 B with phi)
BYTES_PER_PUSH = 112


def peaks():
    fn = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(fn):
        return json.load(open(fn)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# ---------------------------------------------------------------------------------------------------------------
# synthetic state

def halton(n, base, skip=1):
    """Radical-inverse sequence (the reference's generator, stdinclude.cpp:45-73), vectorised."""
    i = np.arange(skip, skip + n, dtype=np.int64)
    f = np.ones(n)
    r = np.zeros(n)
    while i.max() > 0:
        f /= base
        r += f * (i % base)
        i //= base
    return r


def synthetic_bunch(p, n, zfrac=0.8, seed_offset=1, sigma_t=95.3, trunc_t=400.0, sigma_gb=0.0105, zlo=None, zhi=None):
    """Halton bunch inside the undulator: transverse Gaussian (Box-Muller on Halton bases 2,3 / 5,7), uniform in z over
    `zfrac` of the slab, Gaussian momentum spread, entrance flag e = 1."""
    u = [halton(n, b, seed_offset) for b in (2, 3, 5, 7, 11, 13, 17, 19)]
    rad = sigma_t * np.sqrt(-2.0 * np.log(np.maximum(u[0], 1e-300)))
    rad = np.minimum(rad, trunc_t)
    x, y = rad * np.cos(2 * np.pi * u[1]), rad * np.sin(2 * np.pi * u[1])
    zlo = p.zmin if zlo is None else zlo
    zhi = p.zmax if zhi is None else zhi
    zc, zl = 0.5 * (zlo + zhi), (zhi - zlo) * zfrac
    z = zc + zl * (u[2] - 0.5)
    g = sigma_gb * np.sqrt(-2.0 * np.log(np.maximum(u[3], 1e-300)))
    gx, gy = g * np.cos(2 * np.pi * u[4]), g * np.sin(2 * np.pi * u[4])
    gz = 1.0e-3 * (u[5] - 0.5)
    q = np.full(n, 81.85)
    a = np.empty((n, 11))
    a[:, 0], a[:, 1], a[:, 2], a[:, 3] = q, x, y, z
    a[:, 4:7] = a[:, 1:4]
    a[:, 7], a[:, 8], a[:, 9], a[:, 10] = gx, gy, gz, 1.0
    return a


def synthetic_potential(p, amp=1.0e-15, out=None):
    """A_y = amp exp(-(r/sigma)^2) cos(k z) on the local slab, reference layout double[np*N0*N1][3]."""
    N0, N1, npl = p.N0, p.N1, p.np
    x = p.xmin + p.dx * np.arange(N0)
    y = p.ymin + p.dy * np.arange(N1)
    z = p.zmin + p.dz * (p.k0 + np.arange(npl))
    sig = 0.25 * (p.xmax - p.xmin)
    env = np.exp(-(x[:, None] ** 2 + y[None, :] ** 2) / sig ** 2)
    if out is None:
        out = np.zeros((npl, N0, N1, 3))
    else:
        out = out.reshape(npl, N0, N1, 3)
        out[...] = 0.0
    k = 2 * np.pi / (12.0 * p.dz)
    out[..., 1] = amp * np.cos(k * z)[:, None, None] * env[None, :, :]
    out[..., 0] = 0.3 * amp * np.sin(k * z)[:, None, None] * env[None, :, :]
    return out.reshape(-1)


def undulator_time(p, periods=20.0):
    """Bunch time at which the bunch centre sits `periods` undulator periods inside the first module."""
    U = p.undulator[0]
    return periods * U.lu / (p.gamma * p.beta * p.c0) - p.dt_shift


# ---------------------------------------------------------------------------------------------------------------
# clocks

class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML every 2 ms (nvidia-smi every 100 ms if NVML is absent)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.sm, self.mx, self.reasons, self._stop, self._t = index, [], [], set(), threading.Event(), None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _run_nvml(self):
        n = self.nvml
        bits = (("hw_slowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                ("hw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                ("sw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                ("sw_power_cap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))
        try:
            self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
        except Exception:
            pass
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, bit in bits:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def _run_smi(self):
        while not self._stop.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                               "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                r = [c.strip() for c in out.split(",")]
                if r and r[0].replace(".", "").isdigit():
                    self.sm.append(float(r[0]))
                if len(r) > 1 and r[1].replace(".", "").isdigit():
                    self.mx.append(float(r[1]))
                for i, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
                    if len(r) > i and r[i].lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._t = threading.Thread(target=self._run_nvml if self.nvml else self._run_smi, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "how": "nvml" if self.nvml else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------------------
# the reference's own CPU path on a bounded sample (cpu_baseline, --impl reference)

def run_reference_sample(sample_job, steps, warmup, ranks=None):
    from oracle import binding
    if not binding.have_reference():
        return None
    ranks = ranks or max(1, (os.cpu_count() or 1))
    ranks = min(ranks, 64)
    work = os.path.join("/tmp", "mithra-bench-ref-%d" % os.getpid())
    os.makedirs(work, exist_ok=True)
    env = dict(os.environ, MINIMPI_NP=str(ranks))
    out = subprocess.check_output([binding.REF_DUMP, os.path.join(ROOT, sample_job), os.path.join(work, "b"), str(steps),
                                   "--quiet", "--bench", str(warmup)], cwd=work, env=env).decode()
    line = [l for l in out.splitlines() if l.startswith("BENCH ")][-1]
    r = json.loads(line[6:])
    nodes = r["N0"] * r["N1"] * r["N2"]
    r["cells_per_s"] = nodes * r["steps"] / r["seconds"]
    r["pushes_per_s"] = r["pushes"] / r["seconds"]
    r["nodes"] = nodes
    return r


def run_host_binary(job, steps, warmup, nodes):
    """The same job through the C++ host (mithra_b200/host/mithra_b200: the reference's main() / Solver surface over the C
    ABI, INTEGRATION.md option A): the job file as shipped under jobs/, the host's own initialize() (bunch generated on the
    device), `warmup` + `steps` field steps of Solver::solve(), step time from the host's own "Time march" line.  Never
    fatal for the bench line: any failure is reported under "error"."""
    import re
    import shutil
    import tempfile
    exe = os.path.join(ROOT, "mithra_b200", "host", "mithra_b200")
    if not os.path.exists(exe):
        return {"error": "mithra_b200/host/mithra_b200 is not built"}
    work = tempfile.mkdtemp(prefix="mithra-bench-host-")
    try:
        t0 = time.perf_counter()
        r = subprocess.run([exe, os.path.join(ROOT, job), "--steps", str(steps + warmup)], cwd=work, timeout=180,
                           env=dict(os.environ, MITHRA_HOST_TIMING_SKIP=str(warmup)), stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, errors="replace")
        wall = time.perf_counter() - t0
        m = re.findall(r"Time march: (\d+) field steps in ([0-9.eE+-]+) s", r.stdout or "")
        if r.returncode != 0 or not m:
            tail = (r.stdout or "").strip().splitlines()[-1:] or [""]
            return {"error": "exit %d: %s" % (r.returncode, tail[0][-200:])}
        n, sec = int(m[-1][0]), float(m[-1][1])
        return {"what": "%s through the C++ host binary (Solver::solve over the C ABI, its own initialize()): %d field steps after %d warm-up steps" % (job, n, warmup),
                "steps": n, "ms_per_step": 1e3 * sec / n, "value": nodes * n / sec, "unit": "cell-updates/s",
                "wall_seconds_with_initialize": wall}
    except Exception as e:                                              # noqa: BLE001 -- an extra leg must not cost the line
        return {"error": "%s: %s" % (type(e).__name__, e)}
    finally:
        shutil.rmtree(work, ignore_errors=True)


def run_oracle_sample(p_full, steps):
    """Fallback CPU baseline when oracle/_ref is absent: the C port on a thin slab (1 thread)."""
    import copy
    from oracle import binding
    p = copy.copy(p_full)
    p.np, p.N2 = 66, 66
    p.zmax = p.zmin + (p.N2 - 1) * p.dz
    p.zp[0], p.zp[1], p.Lz = p.zmin, p.zmax, p.zmax - p.zmin
    p.power.enabled, p.screens.enabled, p.seed_enabled = 0, 0, 0
    o = binding.Oracle(p)
    n = 32768
    o.upload_particles(synthetic_bunch(p, n))
    o.upload_fields(an=synthetic_potential(p), anm1=synthetic_potential(p))
    tb = undulator_time(p)
    o.set_time(tb, tb, 0)
    o.step(1)
    t0 = time.perf_counter()
    o.step(steps)
    sec = time.perf_counter() - t0
    nodes = p.N0 * p.N1 * p.np
    return {"seconds": sec, "steps": steps, "ranks": 1, "nodes": nodes, "cells_per_s": nodes * steps / sec,
            "pushes_per_s": n * p.n_update_bunch * steps / sec, "particles": n}


# ---------------------------------------------------------------------------------------------------------------
# the N-slab run of a reduced mesh against the 1-slab run of the same problem (carried by every multi-GPU line)

def slab_cross_check(p, wl, dist, rank, world, local_rank, nsteps=20):
    """A mesh of 64 planes per GPU with 65536 macro-particles: `world` slabs (one per GPU, peer-memory exchange) against
    ONE slab holding the whole mesh on rank 0's GPU.  Returns the relative differences of sum |A|^2, sum |J|^2, of the
    centre of charge and the particle count -- the k-slab run must reproduce the single-slab result."""
    import copy
    import torch
    from mithra_b200 import abi, slabs
    pg = copy.copy(p)
    pg.N2 = 64 * world + 2
    pg.np = pg.N2
    pg.zmax = pg.zmin + (pg.N2 - 1) * pg.dz
    pg.Lz = pg.zmax - pg.zmin
    pg.zp[0], pg.zp[1] = pg.zmin, pg.zmax
    pg.power.enabled, pg.screens.enabled = 0, 0
    n = 65536
    bunch = synthetic_bunch(pg, n, sigma_t=wl["sigma_t"], trunc_t=wl["trunc_t"], sigma_gb=wl["sigma_gb"])
    a_n = synthetic_potential(pg)
    tb = undulator_time(pg)

    def run(q, an_loc, part):
        q.device = local_rank
        q.max_particles = n + 1024
        s = abi.GpuSolver(q)
        if q.size > 1:
            s.connect_neighbours(dist, rank, world)
        s.set_time(tb, tb, 0)
        s.upload_fields(an=an_loc, anm1=an_loc * 0.999)
        s.upload_particles(part)
        s.step(nsteps)
        s.synchronize()
        f = s.download_fields(("an", "anp1"))
        pt = s.download_particles()
        s.close()
        return f, pt

    q = slabs.slab_params(pg, rank, world)
    own = (bunch[:, 3] >= q.zp[0]) & (bunch[:, 3] < q.zp[1])
    f, pt = run(q, slabs.scatter_field(pg, a_n, 3, rank, world), bunch[own])
    mine = slabs.owned_planes(pg.N2, rank, world) - q.k0
    loc = [float((f[k].reshape(q.np, -1)[mine] ** 2).sum()) for k in ("an", "anp1")]
    loc += [float(pt.shape[0]), float((pt[:, 0] * pt[:, 3]).sum()), float(pt[:, 0].sum())]
    t = torch.tensor(loc, device="cuda", dtype=torch.float64)
    dist.all_reduce(t)
    got = t.cpu().numpy()
    out = None
    if rank == 0:
        f1, p1 = run(copy.copy(pg), a_n, bunch)
        want = np.array([(f1["an"] ** 2).sum(), (f1["anp1"] ** 2).sum(), p1.shape[0], (p1[:, 0] * p1[:, 3]).sum(), p1[:, 0].sum()])
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
        out = {"what": "%d slabs on %d GPUs against 1 slab, %d x %d x %d nodes, %d macro-particles, %d field steps" % (
                   world, world, pg.N0, pg.N1, pg.N2, n, nsteps),
               "sum_A2": float(want[0]), "sum_J2": float(want[1]),
               "rel_diff_sum_A2": float(rel[0]), "rel_diff_sum_J2": float(rel[1]), "particles": [int(got[2]), int(want[2])],
               "rel_diff_charge_centre_z": float(abs(got[3] / got[4] - want[3] / want[4]) / abs(pg.zmax - pg.zmin)),
               "ok": bool(rel[0] < 1e-9 and rel[1] < 1e-6 and int(got[2]) == int(want[2]))}
    dist.barrier()
    return out


# ---------------------------------------------------------------------------------------------------------------

def mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 1048576.0
    except OSError:
        pass
    return 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--workload", default="fel-lcls", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=("weak", "strong"),
                    help="N > 1: weak = the workload's mesh and bunch per GPU; strong = the workload's mesh split over the GPUs")
    ap.add_argument("--particles", type=int, default=0, help="override the macro-particle count (0 = the workload's)")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="repeat the timed block of K steps until this much device time is measured")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-host-binary", action="store_true", help="skip the extra leg that runs the job through the C++ host binary")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    meta_fn, sample_job, sample_frac, desc = wl["meta"], wl["sample"], wl["frac"], wl["desc"]
    pk, pk_kind = peaks()
    K, W = max(args.steps, 1), max(args.warmup, 3)

    from mithra_b200 import meta as mmeta
    meta = dict(np.load(os.path.join(ROOT, meta_fn)))
    p = mmeta.params_from_meta(meta)
    npart_total = args.particles or wl["particles"]
    sc = int(p.space_charge)

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = min(max(1, (os.cpu_count() or 1)), 64)
        nodes_full = p.N0 * p.N1 * p.N2
        # the unmodified reference advances 3.7e6 (FEL-LCLS, measured) to 7e6 nodes per second and core; it keeps about 110 bytes per node
        est = (K + W) * nodes_full / (3.7e6 * cores) + 30.0
        same = est <= 300.0 and mem_available_gb() >= 1.6 * 110.0 * nodes_full / 1e9 and not args.particles
        job = wl["job"] if same else sample_job
        r = run_reference_sample(job, K, W)
        kind = "reference"
        if r is None:
            r, kind, same = run_oracle_sample(p, K), "port", False
        what = ("the whole job %s (same mesh and macro-particle count as the GPU arm)" % job) if same else (
            "1/%d z-slice (%s), %d nodes, %d macro-particles" % (sample_frac, sample_job, r["nodes"], int(r["particles"])))
        line = {
            "impl": "reference", "metric": "cell-updates/s", "value": r["cells_per_s"], "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * r["seconds"] / r["steps"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "pushes": {"value": r["pushes_per_s"], "unit": "particle-pushes/s"},
            "config": {"workload": desc, "same_mesh_as_gpu_arm": bool(same), "sample": what},
            "cpu_baseline": {"value": r["cells_per_s"], "unit": "cell-updates/s", "cores": r["ranks"], "kind": kind,
                             "sample": "%d field steps of %s" % (K, what), "pushes_per_s": r["pushes_per_s"]},
            "e2e": {"value": r["cells_per_s"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- our arm
    import copy
    from mithra_b200 import abi, slabs
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if abi.load().mithra_gpu_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the time-march has no CPU path")

    check = None
    if world > 1 and not args.no_check:
        check = slab_cross_check(p, wl, dist, rank, world, local_rank)
        if rank == 0 and not check["ok"]:
            raise SystemExit("bench.py: the %d-slab run does not reproduce the single-slab run: %s" % (world, json.dumps(check)))

    strong = world > 1 and args.scaling == "strong"
    if world > 1:
        pg = copy.copy(p)
        if not strong:
            # weak scaling over z-slabs: the mesh grows with the GPU count so that every GPU keeps the workload's z-extent
            # (mesh and bunch scaled per GPU, BASELINE.json configs[4]); the partition is the reference's (solver.cpp:619-641)
            pg.N2 = (p.N2 - 2) * world + 2
            pg.np = pg.N2
            pg.zmax = pg.zmin + (pg.N2 - 1) * pg.dz
            pg.Lz = pg.zmax - pg.zmin
            pg.zp[0], pg.zp[1] = pg.zmin, pg.zmax
        pl = slabs.slab_params(pg, rank, world)
    else:
        pg = pl = p
    pl.device = local_rank
    if strong:
        # the workload's own bunch, every slab keeps what it owns (the bunch fills the middle 80 % of z: uneven slabs)
        gb = synthetic_bunch(pg, npart_total, sigma_t=wl["sigma_t"], trunc_t=wl["trunc_t"], sigma_gb=wl["sigma_gb"])
        bunch = np.ascontiguousarray(gb[(gb[:, 3] >= pl.zp[0]) & (gb[:, 3] < pl.zp[1])])
        del gb
    else:
        bunch = synthetic_bunch(pl, npart_total, seed_offset=1 + rank * npart_total, zlo=pl.zp[0], zhi=pl.zp[1],
                                sigma_t=wl["sigma_t"], trunc_t=wl["trunc_t"], sigma_gb=wl["sigma_gb"])
    npart_local = int(bunch.shape[0])
    pl.max_particles = int(max(npart_local, npart_total // world) * 1.25) + 4096
    pl.max_screen_records = 1 << 16

    # the state lives in pinned host memory from the start: the e2e leg uploads from there
    import torch
    nodes_local = pl.N0 * pl.N1 * pl.np

    # pinned when the box has the memory for it (17 GB per rank at FEL-LCLS scale; 8 ranks share one host), pageable else
    per_rank_gb = mem_available_gb() / max(1, world)
    pin_ok = per_rank_gb > 3.0 * 24.0 * nodes_local / 1e9 + 8.0

    def pinned(n):
        if not pin_ok:
            return np.empty(n, dtype=np.float64)
        return torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()

    a_n = synthetic_potential(pl, out=pinned(nodes_local * 3))
    # A^{n-1} = 0.999 A^n in its own buffer; on a host too small for two levels per rank the two share one buffer
    tight = per_rank_gb < 2.0 * 24.0 * nodes_local / 1e9 + 12.0
    if tight:
        a_nm1 = a_n
    else:
        a_nm1 = pinned(nodes_local * 3)
        np.multiply(a_n, 0.999, out=a_nm1)
    pb = pinned(max(1, bunch.size)); pb[:bunch.size] = bunch.reshape(-1); bunch = pb[:bunch.size].reshape(-1, 11)
    tb = undulator_time(pl)

    clocks = ClockSampler(local_rank)
    solver = abi.GpuSolver(pl)
    if world > 1:
        solver.connect_neighbours(dist, rank, world)

    def load_state(s):
        s.set_time(tb, tb, 0)
        s.upload_fields(an=a_n, anm1=a_nm1)
        s.upload_particles(bunch)

    load_state(solver)

    def barrier():
        solver.synchronize()
        if dist is not None:
            dist.barrier()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(xs):
        if dist is None:
            return [float(x) for x in xs]
        t = torch.tensor(xs, device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        return [float(x) for x in t.cpu().numpy()]

    # device-resident timing: W warm-up steps, then blocks of exactly K steps, each between two barriers, repeated until
    # --min-seconds of device time have been measured; the reported step time is the MEDIAN block (max over ranks per block)
    solver.step(W)
    barrier()
    clocks.start()
    blocks, cells_blk, pushes_blk, launches = [], 0.0, 0.0, 0
    while True:
        c0 = solver.counters()
        barrier()
        ms = solver.step_timed(K)
        barrier()
        c1 = solver.counters()
        blocks.append(allmax(ms))
        cells_blk, pushes_blk = allsum([c1.cell_updates - c0.cell_updates, c1.particle_pushes - c0.particle_pushes])
        launches = int(c1.kernel_launches - c0.kernel_launches)
        # every rank takes the same decision: the block times are already the max over the ranks
        if sum(blocks) >= 1e3 * args.min_seconds or len(blocks) >= 50:
            break
    clk = clocks.stop()
    ms = float(np.median(blocks))
    sec = ms * 1e-3
    cells, pushes = cells_blk, pushes_blk

    # per-phase device times (CUDA events around each kernel group on the library's stream), same state
    nprof = min(K, 10)
    phases = solver.step_profiled(nprof)
    per = {k: v / nprof for k, v in phases.items()}
    step_ms = ms / K
    cell_b, npush = BYTES_PER_CELL[sc], npart_local * pl.n_update_bunch
    algo_step = cell_b * nodes_local + BYTES_PER_PUSH * npush + 56 * npart_local          # bytes one field step must move
    achieved_step = algo_step / (step_ms * 1e-3) / 1e9
    # the dominant kernel, stencil_stream.  Seeded jobs: one launch advances the nodes that are not on the rim (the two outer
    # interior layers in x and y belong to rim_update, timed under "boundary"), (N0-6)(N1-6) nodes of each of the np-2 updated
    # planes.  Jobs without a seed: every interior node and the y faces, (N0-2) N1 nodes per plane (the x faces are the pass of
    # boundary_faces timed under "boundary").
    stencil_ms = per["stencil"]
    rim = pl.N0 >= 8 and pl.N1 >= 8 and pl.np >= 8
    planes = pl.np - 2
    faces_fused = rim and not pl.seed_enabled
    stencil_nodes = ((pl.N0 - 2) * pl.N1 if faces_fused else (pl.N0 - 6) * (pl.N1 - 6) if rim else (pl.N0 - 2) * (pl.N1 - 2)) * planes
    stencil_bytes = cell_b * stencil_nodes
    field_ms = per["stencil"] + per["boundary"]
    traffic, traffic_src = None, None
    tfn = os.path.join(ROOT, "profiles", "traffic.json")            # dram bytes per step from the committed ncu --set full capture
    if os.path.exists(tfn) and world == 1:
        tj = json.load(open(tfn)).get(args.workload, {})
        traffic, traffic_src = tj.get("dram_bytes_per_step"), tj.get("source")

    def gbs(nbytes, t_ms):
        return nbytes / (t_ms * 1e-3) / 1e9 if t_ms > 0 else None

    roofline = {
        "bound": "hbm", "what": "the WHOLE field step: 96/128 B x nodes + 112 B x pushes + 56 B x particles (SURVEY 8d) over ms_per_step",
        "achieved": achieved_step, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved_step / pk["hbm_gbs"], "peak_kind": pk_kind,
        "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_step": algo_step,
        "phases_ms_per_step": per,
        "phases_note": "phases are timed one after the other with a synchronisation in between (10 steps); inside a step the side-stream "
                       "work overlaps, so their sum exceeds ms_per_step",
        "kernels": {
            "stencil_stream": {"ms": stencil_ms, "algorithmic_bytes": stencil_bytes, "units": stencil_nodes,
                               "achieved": gbs(stencil_bytes, stencil_ms), "frac": gbs(stencil_bytes, stencil_ms) / pk["hbm_gbs"],
                               "share_of_step": stencil_ms / max(1e-9, sum(per.values()))},
            "field_update": {"what": "whole fieldUpdate (seed table, stencil_stream, rim_update or the x faces, z shell / faces, edges, corners) over all nodes",
                             "ms": field_ms, "achieved": gbs(cell_b * nodes_local, field_ms), "frac": gbs(cell_b * nodes_local, field_ms) / pk["hbm_gbs"]},
            "push_particles": {"ms": per["push"], "achieved": gbs(BYTES_PER_PUSH * npush, per["push"]), "bytes_per_push": BYTES_PER_PUSH},
            "deposit_current": {"ms": per["deposit"], "achieved": gbs(56 * npart_local, per["deposit"]), "bytes_per_particle": 56},
        },
    }

    # end to end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        out_p = pinned(int(pl.max_particles) * 11)
        solver.close()
        solver = abi.GpuSolver(pl)
        if world > 1:
            solver.connect_neighbours(dist, rank, world)
        solver.set_time(tb, tb, 0)
        solver.upload_fields(an=a_n, anm1=a_nm1)      # warm the allocator / page tables
        solver.step(1)
        barrier()
        t0 = time.perf_counter()
        solver.set_time(tb, tb, 0)
        solver.upload_fields(an=a_n, anm1=a_nm1)
        solver.upload_particles(bunch)
        d2h = 0
        for _ in range(K):
            solver.step(1)
            row = solver.fetch_power()
            d2h += row.nbytes
        got_p = solver.download_particles(out=out_p)
        got_a = solver.download_fields(("an",), out={"an": a_nm1})["an"]      # into the buffer A^n-1 was uploaded from
        barrier()
        e2e_sec = allmax(time.perf_counter() - t0)
        h2d = a_n.nbytes + a_nm1.nbytes + bunch.nbytes
        d2h += got_p.nbytes + got_a.nbytes
        nodes_all, push_all = allsum([nodes_local, npush])
        e2e = {"value": nodes_all * K / e2e_sec, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d / K,
               "d2h_bytes_per_step": d2h / K, "seconds": e2e_sec,
               "what": "%s upload of A^n, A^n-1 and the bunch + K x (step + power row read-back) + download of the bunch and A^n" % ("pinned" if pin_ok else "pageable (host memory too small to pin)"),
               "pushes_per_s": push_all * K / e2e_sec}
    solver.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = run_reference_sample(sample_job, 6, 2)
        kind = "reference"
        if r is None:
            r, kind = run_oracle_sample(p, 3), "port"
        cpu = {"value": r["cells_per_s"], "unit": "cell-updates/s", "cores": r["ranks"], "kind": kind,
               "sample": "%s: 6 field steps of the 1/%d z-slice (%d nodes, %d macro-particles)" % (
                   sample_job, sample_frac, r["nodes"], int(r["particles"])),
               "pushes_per_s": r["pushes_per_s"]}

    host = None
    if rank == 0 and world == 1 and not args.no_host_binary:
        host = run_host_binary(wl["job"], K, W, p.N0 * p.N1 * p.N2)

    nodes_max, parts_all = allmax(float(nodes_local)), allsum([npart_local])[0]
    if rank == 0:
        line = {
            "metric": "cell-updates/s", "value": cells / sec, "unit": "cell-updates/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms, "ms_per_step_min": min(blocks) / K, "blocks": len(blocks), "timed_seconds": sum(blocks) * 1e-3,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "pushes": {"value": pushes / sec, "unit": "particle-pushes/s"},
            "config": {"workload": desc, "parallelism": "z-slabs x%d" % world, "l2": "inputs larger than L2 (%.1f GB of potentials per GPU)" % (
                4 * (32 if sc else 24) * nodes_max / 1e9), "nodes_per_gpu": int(nodes_max), "particles_total": int(parts_all),
                "mesh": "%d x %d x %d" % (pg.N0, pg.N1, pg.N2), "host_buffers": "pinned" if pin_ok else "pageable",
                "initial_levels": "A^n-1 = A^n (host memory)" if tight else "A^n-1 = 0.999 A^n"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk, "check": check,
            "host_binary": host,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
