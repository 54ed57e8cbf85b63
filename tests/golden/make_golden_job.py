#!/usr/bin/env python
"""Golden fixture of a full-size shipped job: the UNMODIFIED reference (oracle/_ref/ref_dump, all host cores through
the forked mini-MPI) runs jobs/<name>.job for NSTEPS field steps from its own initial state; kept are the power rows,
the field-sampling text file and a few scalars of initialize() -- a few KB (the particle lists and field dumps of
make_golden.py would be tens of MB at this size).

    python tests/golden/make_golden_job.py fel-ir 3000       # needs /root/reference; hours on one core: the radiation of the
                                                             # shipped FEL-IR job reaches its power plane after ~2200 field steps,
                                                             # which is why no such fixture is committed
"""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import binding  # noqa: E402


def main(job, nsteps):
    work = tempfile.mkdtemp(prefix="golden-job-")
    try:
        prefix = os.path.join(work, "g")
        os.environ.setdefault("MINIMPI_NP", "1")      # one rank: the layout the host binary reproduces (-0.txt files)
        binding.run_ref_dump(os.path.join(ROOT, "jobs", job + ".job"), prefix, nsteps, cwd=work)
        meta = binding.read_records(prefix + ".meta.bin")
        out = {"meta/" + k: v for k, v in meta.items() if v.size <= 32}
        out["nsteps"] = np.array([nsteps])
        out["power"] = binding.read_records(prefix + ".power.bin")["pG"].reshape(nsteps, -1)
        for d in ("field-sampling",):
            dd = os.path.join(work, d)
            if os.path.isdir(dd):
                for fn in sorted(os.listdir(dd)):
                    out["txt/%s/%s" % (d, fn)] = np.frombuffer(open(os.path.join(dd, fn), "rb").read(), dtype=np.uint8)
        dst = os.path.join(HERE, "job-%s.npz" % job)
        np.savez_compressed(dst, **out)
        print(job, "steps", nsteps, "power[-1]", out["power"][-1], "->", os.path.getsize(dst) // 1024, "KB")
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    if not binding.have_reference():
        sys.exit("oracle/_ref/ref_dump is missing: run `make -C oracle ref` where /root/reference exists")
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 300)
