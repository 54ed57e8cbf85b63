"""Two (or more) slabs, one process per GPU: K steps through mithra_gpu_step, state written to a file.
Run twice (with and without MITHRA_NO_LOOKAHEAD=1) and compare with --compare: the early enqueue of the next field
update's first half must not change the result.

    torchrun --nproc-per-node 2 tools/check_lookahead.py out_a ;  MITHRA_NO_LOOKAHEAD=1 torchrun ... out_b
    python tools/check_lookahead.py --compare out_a out_b 2
"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(prefix, K=37):
    import torch
    import torch.distributed as dist
    import bench
    from mithra_b200 import abi, meta as mmeta, slabs
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p = mmeta.params_from_meta(dict(np.load(os.path.join(ROOT, "bench", "fel-seeded.meta.npz"))))
    pg = copy.copy(p)
    pg.N2 = 600 * world + 2; pg.np = pg.N2
    pg.zmax = pg.zmin + (pg.N2 - 1) * pg.dz; pg.Lz = pg.zmax - pg.zmin
    pg.zp[0], pg.zp[1] = pg.zmin, pg.zmax
    pg.power.enabled, pg.screens.enabled = 0, 0
    pl = slabs.slab_params(pg, rank, world)
    pl.device = local
    n = 200000
    pl.max_particles = 2 * n
    s = abi.GpuSolver(pl)
    s.connect_neighbours(dist, rank, world)
    # the whole slab in z so that particles do cross between slabs
    b = bench.synthetic_bunch(pl, n, zfrac=0.999, seed_offset=1 + rank * n, zlo=pl.zp[0], zhi=pl.zp[1])
    b[:, 9] = 0.4 * (np.arange(n) % 7 - 3) / 3.0                      # gb_z up to +-0.4: a fraction of a cell per step
    a = bench.synthetic_potential(pl)
    tb = bench.undulator_time(pl)
    s.set_time(tb, tb, 0)
    s.upload_fields(an=a, anm1=a * 0.999)
    s.upload_particles(b)
    s.step(K)
    s.synchronize()
    q = s.download_particles()
    f = s.download_fields(("an",))["an"]
    np.savez(prefix + ".r%d.npz" % rank, particles=q, an=f)
    dist.barrier()
    s.close()
    dist.destroy_process_group()


def compare(a, b, world):
    for r in range(world):
        A, B = np.load(a + ".r%d.npz" % r), np.load(b + ".r%d.npz" % r)
        fa, fb = A["an"], B["an"]
        rel = np.linalg.norm(fa - fb) / np.linalg.norm(fb)
        pa, pb = A["particles"], B["particles"]
        assert pa.shape == pb.shape, (pa.shape, pb.shape)
        ka, kb = np.lexsort((pa[:, 1], pa[:, 2], pa[:, 3])), np.lexsort((pb[:, 1], pb[:, 2], pb[:, 3]))
        dp = np.abs(pa[ka] - pb[kb]).max()
        print("rank", r, "fields rel", rel, "particles", pa.shape[0], "max |diff|", dp)
        assert rel < 1e-11 and dp < 1e-8
    print("lookahead == plain order")


if __name__ == "__main__":
    if sys.argv[1] == "--compare":
        compare(sys.argv[2], sys.argv[3], int(sys.argv[4]))
    else:
        run(sys.argv[1])
