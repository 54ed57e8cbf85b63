#!/usr/bin/env python
"""Golden fixtures for the bunch generators no time-march fixture uses: `manual`, `3D-crystal`, `file`, the gaussian
ellipsoid, bunching factor with a phase, shot noise (both profiles), several positions / several bunches.

    python tests/golden/make_golden_init.py       # needs /root/reference (oracle/_ref/ref_dump)

Each tests/golden/<job>.npz holds what the UNMODIFIED reference's Solver::initialize() produced for tests/jobs/<job>.job:
    meta/<name>     every scalar / coefficient table (ref_dump meta record)
    p0              the initial particle list (n x 11: q, rnp, rnm, gb, e) in the reference's order
tests/test_host.py compares the host's initialize() with them bit for bit (classes.cpp:60-420, solver.cpp:1126-1180,
263-423, 508-540 of the reference)."""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import binding  # noqa: E402

INIT_JOBS = ("init-manual", "init-crystal", "init-file", "init-gauss", "init-shot", "init-shotg")
SIDE_FILES = ("init-file-bunch.txt",)           # looked up in the working directory by the `file` bunch


def make(job):
    work = tempfile.mkdtemp(prefix="golden-init-")
    try:
        for s in SIDE_FILES:
            shutil.copy(os.path.join(ROOT, "tests", "jobs", s), work)
        prefix = os.path.join(work, "g")
        binding.run_ref_dump(os.path.join(ROOT, "tests", "jobs", job + ".job"), prefix, 0, full_at=(0,), cwd=work)
        out = {"meta/" + k: v for k, v in binding.read_records(prefix + ".meta.bin").items()}
        out["p0"] = binding.read_records(prefix + ".full0.bin")["particles"].reshape(-1, 11)
        np.savez_compressed(os.path.join(HERE, job + ".npz"), **out)
        print("%-14s particles %5d -> %.0f KB" % (job, out["p0"].shape[0], os.path.getsize(os.path.join(HERE, job + ".npz")) / 1024.0))
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    if not binding.have_reference():
        sys.exit("oracle/_ref/ref_dump is missing: run `make -C oracle ref` where /root/reference exists")
    for j in (sys.argv[1:] or INIT_JOBS):
        make(j)
