#!/usr/bin/env python
"""bench.py -- throughput of the B200-native MITHRA time-march on BASELINE.json's workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload fel-seeded|...]

One "step" = one field time step of Solver::solve (solver.cpp:1300-1399): fieldUpdate (stencil + absorbing boundaries
+ TF/SF seed + E/B evaluation), nUpdateBunch Boris sub-pushes of every macro-particle, screens, radiated power,
fieldShift, currentReset, ZigZag deposit.  Workload at N=1: BASELINE.json configs[1], the FEL-SEEDED amplifier
(jobs/fel-seeded.job: 85 x 85 x 8252 nodes, 4,194,304 macro-particles, 3 sub-pushes per step); the mesh constants and
update coefficients are the reference's own Solver::initialize() results for that job (bench/fel-seeded.meta.npz,
written by oracle/_ref/ref_dump), the bunch and the potentials are synthetic (Halton bunch inside the undulator,
smooth wave packet) as SURVEY.md section 8(d) specifies.

The JSON line (rank 0): metric = cell-updates/s of the whole job, with the particle pushes/s of the same timed region
under "pushes"; "roofline" for the dominant kernel (interior stencil) from CUDA-event phase timing inside this run;
"cpu_baseline" = the unmodified reference (oracle/_ref/ref_dump --bench on the mini-MPI ranks of all host cores) on
a bounded 1/8 z-slice of the same job; "e2e" = the same job through the C ABI with host buffers (pinned upload of
potentials + bunch, per-step power read-back, final download inside the timed region).

--impl reference times the reference's own CPU implementation alone, on the bounded sample, with the same keys.
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
    "fel-seeded": dict(meta="bench/fel-seeded.meta.npz", sample="jobs/fel-seeded-sample.job", frac=8, particles=4194304,
                       sigma_t=95.3, trunc_t=400.0, sigma_gb=0.0105,
                       desc="FEL-SEEDED 85x85x8252 nodes, 4194304 macro-particles, 3 sub-pushes/step, NSFD, seed TF/SF, 1 power plane, 7 screens"),
    # BASELINE.json configs[3]: the large-z X-ray FEL mesh on ONE GPU (46 GB resident); --gpus N stacks N of them
    "fel-lcls": dict(meta="bench/fel-lcls.meta.npz", sample="jobs/fel-lcls-sample.job", frac=32, particles=8388608,
                     sigma_t=30.0, trunc_t=180.0, sigma_gb=0.007,
                     desc="FEL-LCLS 102x102x33335 nodes, 8388608 macro-particles, 1 sub-push/step, NSFD, 1 power plane"),
    # BASELINE.json configs[4]: FdTdSC (A + phi) weak-scaling unit, 102 x 102 x 4096 cells and 1 Mi particles per GPU
    "sc-weak": dict(meta="bench/sc-weak.meta.npz", sample="jobs/sc-weak-sample.job", frac=4, particles=1048576,
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

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)          # 20 warm-up + 200 timed field steps
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--workload", default="fel-seeded", choices=sorted(WORKLOADS))
    ap.add_argument("--particles", type=int, default=0, help="override the macro-particle count (0 = the workload's)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    meta_fn, sample_job, sample_frac, desc = wl["meta"], wl["sample"], wl["frac"], wl["desc"]
    pk, pk_kind = peaks()
    K, W = args.steps, max(args.warmup, 3)

    from mithra_b200 import meta as mmeta
    meta = dict(np.load(os.path.join(ROOT, meta_fn)))
    p = mmeta.params_from_meta(meta)
    npart_total = args.particles or wl["particles"]
    sc = int(p.space_charge)

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        r = run_reference_sample(sample_job, K, W)
        kind = "reference"
        if r is None:
            r, kind = run_oracle_sample(p, K), "port"
        line = {
            "impl": "reference", "metric": "cell-updates/s", "value": r["cells_per_s"], "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * r["seconds"] / r["steps"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "pushes": {"value": r["pushes_per_s"], "unit": "particle-pushes/s"},
            "config": {"workload": desc, "sample": "1/%d z-slice (%s), %d nodes, %d macro-particles" % (
                sample_frac, sample_job, r["nodes"], int(r["particles"]))},
            "cpu_baseline": {"value": r["cells_per_s"], "unit": "cell-updates/s", "cores": r["ranks"], "kind": kind,
                             "sample": "%s: %d field steps of the 1/%d z-slice" % (sample_job, K, sample_frac),
                             "pushes_per_s": r["pushes_per_s"]},
            "e2e": {"value": r["cells_per_s"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- our arm
    from mithra_b200 import abi
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if abi.load().mithra_gpu_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the time-march has no CPU path")

    if world > 1:
        # weak scaling over z-slabs: the mesh grows with the GPU count so that every GPU keeps the workload's z-extent
        # (mesh and bunch scaled per GPU, BASELINE.json configs[4]); the partition is the reference's (solver.cpp:619-641)
        import copy
        from mithra_b200 import slabs
        pg = copy.copy(p)
        pg.N2 = (p.N2 - 2) * world + 2
        pg.np = pg.N2
        pg.zmax = pg.zmin + (pg.N2 - 1) * pg.dz
        pg.Lz = pg.zmax - pg.zmin
        pg.zp[0], pg.zp[1] = pg.zmin, pg.zmax
        pl = slabs.slab_params(pg, rank, world)
    else:
        pl = p
    pl.device = local_rank
    npart_local = npart_total
    pl.max_particles = int(npart_local * 1.25) + 1024
    pl.max_screen_records = 1 << 16

    clocks = ClockSampler(local_rank)
    solver = abi.GpuSolver(pl)
    if world > 1:
        solver.connect_neighbours(dist, rank, world)
    bunch = synthetic_bunch(pl, npart_local, seed_offset=1 + rank * npart_local, zlo=pl.zp[0], zhi=pl.zp[1],
                            sigma_t=wl["sigma_t"], trunc_t=wl["trunc_t"], sigma_gb=wl["sigma_gb"])
    a_n = synthetic_potential(pl)
    a_nm1 = a_n * 0.999
    tb = undulator_time(pl)

    def load_state(s):
        s.set_time(tb, tb, 0)
        s.upload_fields(an=a_n, anm1=a_nm1)
        s.upload_particles(bunch)

    load_state(solver)
    nodes_local = pl.N0 * pl.N1 * pl.np

    def barrier():
        solver.synchronize()
        if dist is not None:
            dist.barrier()

    # device-resident timing: W warm-up steps, then exactly K steps between two barriers
    solver.step(W)
    barrier()
    c0 = solver.counters()
    clocks.start()
    ms = solver.step_timed(K)
    clk = clocks.stop()
    barrier()
    c1 = solver.counters()
    if dist is not None:
        import torch
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    cells = (c1.cell_updates - c0.cell_updates)
    pushes = (c1.particle_pushes - c0.particle_pushes)
    launches = int(c1.kernel_launches - c0.kernel_launches)
    if dist is not None:
        import torch
        t = torch.tensor([cells, pushes], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        cells, pushes = float(t[0].item()), float(t[1].item())
    sec = ms * 1e-3

    # per-phase device times (CUDA events around each kernel group on the library's stream), same state
    nprof = min(K, 10)
    phases = solver.step_profiled(nprof)
    # The dominant kernel is stencil_stream: one launch advances the nodes that are not on the rim (the two outer interior
    # layers in x and y belong to rim_update, timed under "boundary"), (N0-6)(N1-6) nodes of each of the np-2 updated planes.
    stencil_ms = phases["stencil"] / nprof
    rim = pl.N0 >= 8 and pl.N1 >= 8 and pl.np >= 8
    planes = pl.np - 2                                      # rank 0 updates its planes 1 .. np-2
    stencil_nodes = ((pl.N0 - 6) * (pl.N1 - 6) if rim else (pl.N0 - 2) * (pl.N1 - 2)) * planes
    algo_bytes = BYTES_PER_CELL[sc] * stencil_nodes
    achieved = algo_bytes / (stencil_ms * 1e-3) / 1e9
    push_ms = phases["push"] / nprof
    field_ms = (phases["stencil"] + phases["boundary"]) / nprof
    traffic = None
    tfn = os.path.join(ROOT, "profiles", "traffic.json")            # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tfn) and world == 1 and args.workload == "fel-seeded":
        traffic = json.load(open(tfn)).get("stencil_stream", {}).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "stencil_stream", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "peak_kind": pk_kind, "traffic": traffic,
                "algorithmic_bytes_per_launch": algo_bytes, "units_per_launch": stencil_nodes, "launch_ms": stencil_ms,
                "note": "algorithmic bytes = 96 B (128 B with phi) per node the launch advances (SURVEY 8d: read A^n, A^n-1, J, write A^n+1); "
                        "the kernel reads J only inside the deposit box, so its DRAM traffic is below the algorithmic figure",
                "share_of_step": stencil_ms / max(1e-9, sum(phases.values()) / nprof),
                "phases_ms_per_step": {k: v / nprof for k, v in phases.items()},
                "field_update": {"what": "whole fieldUpdate (seed table, stencil_stream, rim_update, z shell / faces, edges, corners) over all nodes",
                                 "ms": field_ms, "achieved": BYTES_PER_CELL[sc] * nodes_local / (field_ms * 1e-3) / 1e9, "unit": "GB/s",
                                 "frac": BYTES_PER_CELL[sc] * nodes_local / (field_ms * 1e-3) / 1e9 / pk["hbm_gbs"]},
                # the whole field step against the roofline: every algorithmic byte of SURVEY 8(d) (cell-updates, pushes,
                # deposit reads) over the device time of one step -- the figure north_star's ">= 60 % of HBM roofline" is about
                "time_march": {"what": "96/128 B x nodes + 112 B x pushes + 56 B x particles, per field step, over ms_per_step",
                               "achieved": (BYTES_PER_CELL[sc] * nodes_local + BYTES_PER_PUSH * npart_local * pl.n_update_bunch + 56 * npart_local)
                                           / (ms / K * 1e-3) / 1e9, "unit": "GB/s",
                               "frac": (BYTES_PER_CELL[sc] * nodes_local + BYTES_PER_PUSH * npart_local * pl.n_update_bunch + 56 * npart_local)
                                       / (ms / K * 1e-3) / 1e9 / pk["hbm_gbs"]},
                "push": {"achieved": BYTES_PER_PUSH * npart_local * pl.n_update_bunch / (push_ms * 1e-3) / 1e9 if push_ms > 0 else None,
                         "unit": "GB/s", "bytes_per_push": BYTES_PER_PUSH}}

    # end to end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        import torch
        pin = {}
        for name, arr in (("an", a_n), ("anm1", a_nm1), ("bunch", bunch.reshape(-1))):
            t = torch.empty(arr.size, dtype=torch.float64, pin_memory=True)
            t.numpy()[:] = arr
            pin[name] = t.numpy()
        # results land in pinned host memory too
        pin["out_a"] = torch.empty(a_n.size, dtype=torch.float64, pin_memory=True).numpy()
        pin["out_p"] = torch.empty(int(pl.max_particles) * 11, dtype=torch.float64, pin_memory=True).numpy()
        solver.close()
        solver = abi.GpuSolver(pl)
        if world > 1:
            solver.connect_neighbours(dist, rank, world)
        solver.set_time(tb, tb, 0)
        solver.upload_fields(an=pin["an"], anm1=pin["anm1"])      # warm the allocator / page tables
        solver.step(1)
        barrier()
        t0 = time.perf_counter()
        solver.set_time(tb, tb, 0)
        solver.upload_fields(an=pin["an"], anm1=pin["anm1"])
        solver.upload_particles(pin["bunch"].reshape(-1, 11))
        d2h = 0
        for _ in range(K):
            solver.step(1)
            row = solver.fetch_power()
            d2h += row.nbytes
        out_p = solver.download_particles(out=pin["out_p"])
        out_a = solver.download_fields(("an",), out={"an": pin["out_a"]})["an"]
        barrier()
        e2e_sec = time.perf_counter() - t0
        h2d = pin["an"].nbytes + pin["anm1"].nbytes + pin["bunch"].nbytes
        d2h += out_p.nbytes + out_a.nbytes
        if dist is not None:
            t = torch.tensor([e2e_sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_sec = float(t.item())
        e2e = {"value": nodes_local * world * K / e2e_sec, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d / K,
               "d2h_bytes_per_step": d2h / K, "seconds": e2e_sec,
               "what": "pinned upload of A^n, A^n-1 and the bunch + K x (step + power row read-back) + download of the bunch and A^n",
               "pushes_per_s": npart_local * world * pl.n_update_bunch * K / e2e_sec}
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

    if rank == 0:
        line = {
            "metric": "cell-updates/s", "value": cells / sec, "unit": "cell-updates/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "pushes": {"value": pushes / sec, "unit": "particle-pushes/s"},
            "config": {"workload": desc, "parallelism": "z-slabs x%d" % world, "l2": "inputs larger than L2 (%.1f GB of potentials per GPU)" % (
                4 * (32 if sc else 24) * nodes_local / 1e9), "nodes_per_gpu": nodes_local, "particles_per_gpu": npart_local},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
