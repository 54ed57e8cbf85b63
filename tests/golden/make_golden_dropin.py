#!/usr/bin/env python
"""Golden fixture of the drop-in proof: the UNMODIFIED reference executable (oracle/_ref/mithra_ref = src/mithra.cpp's own
main(), built by oracle/Makefile with the single-rank MPI shim) runs tests/jobs/micro-dropin.job (and micro-dropin-bunch.job, the same
job with the three bunch outputs) to its end; every file it writes (radiation power, screens, bunch sampling / profile / .vtu) is kept as written.  tests/test_dropin.py runs the same job through
oracle/_ref/mithra_ref_gpu -- the same main(), parser, Solver::initialize() and Solver::solve(), with
integration/mithra_gpu_dropin.cpp in place of fdtd.cpp / fdtdSC.cpp -- and compares the files.

    python tests/golden/make_golden_dropin.py          # needs /root/reference (oracle/_ref/mithra_ref)
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
JOBS = ("micro-dropin", "micro-dropin-bunch", "micro-dropin-field")


def main(JOB):
    exe = os.path.join(ROOT, "oracle", "_ref", "mithra_ref")
    if not os.path.exists(exe):
        sys.exit("oracle/_ref/mithra_ref is missing: run `make -C oracle ref` where /root/reference exists")
    work = tempfile.mkdtemp(prefix="golden-dropin-")
    try:
        log = subprocess.check_output([exe, os.path.join(ROOT, "tests", "jobs", JOB + ".job")], cwd=work, env=dict(os.environ, MINIMPI_NP="1"))
        out = {}
        for d in sorted(os.listdir(work)):
            dd = os.path.join(work, d)
            if os.path.isdir(dd):
                for fn in sorted(os.listdir(dd)):
                    out["txt/%s/%s" % (d, fn)] = np.frombuffer(open(os.path.join(dd, fn), "rb").read(), dtype=np.uint8)
        assert out, "the reference wrote no file"
        dst = os.path.join(HERE, JOB + ".npz")
        np.savez_compressed(dst, **out)
        print(JOB, {k: v.size for k, v in out.items()}, "->", os.path.getsize(dst) // 1024, "KB")
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    for j in (sys.argv[1:] or JOBS):
        main(j)
