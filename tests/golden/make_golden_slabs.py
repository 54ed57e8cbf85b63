#!/usr/bin/env python
"""Golden fixture of the z-slab partition and the initial particle distribution: the UNMODIFIED reference with 2, 3 and 4
mini-MPI ranks (oracle/_ref/ref_dump <job> <prefix> 0 --init-only under MINIMPI_NP=N) -- for every rank its slab of
solver.cpp:619-641 (np, k0, zp) and the particles distributeParticles (solver.cpp:429-487) left it with.
tests/test_host.py::test_slab_partition_and_particle_distribution_match_the_reference_ranks compares the host's
`--gpus N` partition (one slab per GPU) and ownership split with them.

    python tests/golden/make_golden_slabs.py        # needs /root/reference (oracle/_ref/ref_dump)
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import binding  # noqa: E402

CASES = (("micro-nsfd", 2), ("micro-nsfd", 3), ("micro-nsfd", 4), ("micro-seeded", 2), ("micro-lcls", 3))


def main():
    if not binding.have_reference():
        sys.exit("oracle/_ref/ref_dump is missing: run `make -C oracle ref` where /root/reference exists")
    out = {}
    for job, n in CASES:
        work = tempfile.mkdtemp(prefix="golden-slabs-")
        try:
            subprocess.check_call([binding.REF_DUMP, os.path.join(ROOT, "tests", "jobs", job + ".job"), os.path.join(work, "r"), "0",
                                   "--init-only", "--quiet"], cwd=work, env=dict(os.environ, MINIMPI_NP=str(n)), stdout=subprocess.DEVNULL)
            for r in range(n):
                d = binding.read_records(os.path.join(work, "r.rank%d.bin" % r))
                assert int(d["rank"][0]) == r and int(d["size"][0]) == n
                key = "%s/%d/%d/" % (job, n, r)
                out[key + "slab"] = np.array([int(d["np"][0]), int(d["k0"][0])])
                out[key + "zp"] = np.asarray(d["zp"], dtype=np.float64)
                out[key + "particles"] = d["particles"].reshape(-1, 11)
            print(job, n, [(int(out["%s/%d/%d/slab" % (job, n, r)][0]), out["%s/%d/%d/particles" % (job, n, r)].shape[0]) for r in range(n)])
        finally:
            shutil.rmtree(work, ignore_errors=True)
    np.savez_compressed(os.path.join(HERE, "init-slabs.npz"), **out)


if __name__ == "__main__":
    main()
