#!/usr/bin/env python
"""tests/golden/job-ir-mid.npz: the UNMODIFIED reference's own main (oracle/_ref/mithra_ref) on tests/jobs/ir-mid.job,
once with one rank and once with four forked MPI ranks (oracle/mpi_shim); kept are the two radiated-power files and the
single-rank field-sampling file.  ~14 + 4 minutes.

    python tests/golden/make_golden_mid.py          # needs /root/reference (oracle/_ref built by `make -C oracle ref`)
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "mithra_ref")
JOB = os.path.join(ROOT, "tests", "jobs", "ir-mid.job")

if __name__ == "__main__":
    if not os.path.exists(REF):
        sys.exit("oracle/_ref/mithra_ref is missing: run `make -C oracle ref` where /root/reference exists")
    out = {}
    for ranks, key in ((1, "power_1rank"), (4, "power_4ranks")):
        work = tempfile.mkdtemp(prefix="golden-mid-")
        try:
            subprocess.check_call([REF, JOB], cwd=work, env=dict(os.environ, MINIMPI_NP=str(ranks)), stdout=subprocess.DEVNULL)
            out[key] = np.loadtxt(os.path.join(work, "power-sampling", "power-ir-0.txt"))
            if ranks == 1:
                out["field_sampling"] = np.frombuffer(open(os.path.join(work, "field-sampling", "field-0.txt"), "rb").read(), dtype=np.uint8)
        finally:
            shutil.rmtree(work, ignore_errors=True)
    np.savez_compressed(os.path.join(HERE, "job-ir-mid.npz"), **out)
    print({k: v.shape for k, v in out.items()})
