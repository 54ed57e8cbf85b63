#!/usr/bin/env python
"""Cross-device slab parity, one PROCESS per GPU (the way bench.py and a production run drive a box): every rank opens
its neighbours' arrays with cudaIpcOpenMemHandle, ghost planes / boundary currents / migrating particles travel as
stores into peer memory over NVLink.  The world's slabs must reproduce the single-slab CPU oracle after 100 field steps
of a fixture job -- the test of tests/test_gpu_slabs.py with the slabs on different devices.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/check_slabs_mp.py <job> [--fused]

--fused drives the loop through mithra_gpu_step (look-ahead of the next field update, side streams) instead of the
separate entry points.  Rank 0 prints one line "SLABS-MP {json}" and exits non-zero on a mismatch."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from scipy.spatial import cKDTree
    from mithra_b200 import abi, slabs
    from oracle import binding
    from tests import helpers
    from tests.test_gpu_slabs import PHASES

    job = sys.argv[1]
    fused = "--fused" in sys.argv
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p, meta, g = helpers.params_for(job)
    nsteps = 100

    q = slabs.slab_params(p, rank, world)
    q.device = local
    q.max_particles = g["p0"].shape[0] + 1024
    s = abi.GpuSolver(q)
    s.connect_neighbours(dist, rank, world)
    own = slabs.owner_of(p, g["p0"][:, 3], world)
    t = g["t0"]
    s.set_time(float(t[0]), float(t[1]), int(t[2]))
    s.upload_particles(g["p0"][own == rank])
    if p.seed_enabled:
        s.seedInitial()
    n_before = s.num_particles()
    dist.barrier()
    if fused:
        s.step(nsteps)
    else:
        for _ in range(nsteps):
            for ph in PHASES:
                getattr(s, ph)()
    s.synchronize()
    dist.barrier()

    names = ("an", "anm1", "anp1") + (("fn", "fnm1", "fnp1") if p.space_charge else ())
    mine = {"fields": s.download_fields(names), "particles": s.download_particles(), "power": s.fetch_power(),
            "n_before": n_before, "device": torch.cuda.current_device(), "pid": os.getpid()}
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    s.close()
    ok, out = True, None
    if rank == 0:
        cpu = binding.Oracle(p)
        helpers.start_from_golden(cpu, g)
        for _ in range(nsteps):
            helpers.solve_step(cpu)
        ref = cpu.download_fields(names)
        err = {}
        for k in names:
            nc = 3 if k.startswith("a") else 1
            glob = slabs.gather_field(p, [x["fields"][k] for x in parts], nc, world)
            err[k] = float(helpers.rel_l2(glob, ref[k]))
        allp = np.concatenate([x["particles"] for x in parts])
        pc = cpu.download_particles()
        scale = np.abs(pc[:, [1, 2, 3, 7, 8, 9]]).max(axis=0) + 1e-300
        d, idx = cKDTree(pc[:, [1, 2, 3, 7, 8, 9]] / scale).query(allp[:, [1, 2, 3, 7, 8, 9]] / scale)
        pw = sum(x["power"] for x in parts)
        ref_pw = cpu.fetch_power()
        perr = float(np.abs(pw - ref_pw).max() / max(np.abs(ref_pw).max(), 1e-300))
        moved = [int(x["particles"].shape[0]) - int(x["n_before"]) for x in parts]
        out = {"job": job, "slabs": world, "devices": [x["device"] for x in parts], "pids": len(set(x["pid"] for x in parts)),
               "fused_step": fused, "steps": nsteps, "rel_l2": err, "particles": [int(allp.shape[0]), int(pc.shape[0])],
               "bijection": bool(np.unique(idx).size == idx.size), "particle_max_dist": float(d.max()),
               "power_max_rel": perr, "net_migration_per_slab": moved}
        ok = (all(v < 1e-9 for v in err.values()) and allp.shape[0] == pc.shape[0] and out["bijection"] and d.max() < 1e-8
              and perr < 1e-8 and len(set(out["devices"])) == world and out["pids"] == world)
        out["ok"] = bool(ok)
        print("SLABS-MP " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
