#!/usr/bin/env python
"""Error behaviour of the job-file front end: malformed variants of tests/jobs/micro-nsfd.job through the UNMODIFIED reference
(oracle/_ref/ref_dump <job> <prefix> 0 --init-only: its parser and Solver::initialize()) -- exit code and the last message it
prints (the reference's convention: message to stdout, then exit(1)).  Written to tests/golden/parser-errors.json
with the job text of every case; tests/test_host.py::test_malformed_jobs_stop_like_the_reference runs the host on the same texts.

    python tests/golden/make_golden_errors.py        # needs /root/reference (oracle/_ref/ref_dump)
"""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import binding  # noqa: E402


def last_message(out):
    """The text of the last non-empty line without the `date ::: file:line :::` prefix of printmessage (stdinclude.h:79-90)."""
    lines = [ln for ln in out.strip().splitlines() if ln.strip()]
    last = lines[-1] if lines else ""
    return re.sub(r"^.*?::: [A-Za-z_.]+:\d+ ::: \s*", "", last).strip()


def cases():
    base = open(os.path.join(ROOT, "tests", "jobs", "micro-nsfd.job")).read()

    def sub(old, new):
        assert old in base, old
        return base.replace(old, new, 1)
    ell = "type                        = ellipsoid"
    return {
        "negative-back-shift": sub("total-time                    = 30000", "total-time                    = 30000\n  initial-time-back-shift       = -5.0"),
        "unknown-generator": sub("distribution                = uniform", "distribution                = uniform\n    generator                   = sobol"),
        "bunching-factor-above-two": sub("bunching-factor             = 0.01", "bunching-factor             = 2.5"),
        "unknown-distribution": sub("distribution                = uniform", "distribution                = triangular"),
        "unknown-solver": sub("solver                        = NSFD", "solver                        = LEAPFROG"),
        "bad-boolean": sub("space-charge                  = false", "space-charge                  = maybe"),
        "unknown-mesh-key": sub("mesh-truncation-order         = 2", "mesh-truncation-order         = 2\n  mesh-colour                   = blue"),
        "truncation-order-three": sub("mesh-truncation-order         = 2", "mesh-truncation-order         = 3"),
        "crystal-numbers-mismatch": sub(ell, "type                        = 3D-crystal\n    numbers                     = ( 3, 3, 3 )\n    lattice-constants           = ( 1.0, 1.0, 1.0 )"),
        "file-row-count-mismatch": sub(ell, "type                        = file\n    file-name                   = init-file-bunch.txt"),
        "unknown-bunch-type": sub(ell, "type                        = sphere"),
        "unknown-undulator-key": sub("polarization-angle          = 0.0", "polarization-angle          = 0.0\n    taper                       = 1.0"),
        "unknown-length-scale": sub("length-scale                  = MICROMETER", "length-scale                  = FURLONG"),
        "unknown-power-sampling-type": sub("type                        = at-point\n    directory", "type                        = everywhere\n    directory"),
        "empty-bunch-group": re.sub(r"BUNCH\n\{.*?\n\}\n\nUNDULATOR", "BUNCH\n{\n}\n\nUNDULATOR", base, flags=re.S),
        "empty-bunch-initialization-block": re.sub(r"bunch-initialization\n  \{.*?\n  \}", "bunch-initialization\n  {\n  }", base, flags=re.S),
        "zero-bunch-time-step": sub("bunch-time-step               = 1.6", "bunch-time-step               = 0.0"),
        "unknown-top-level-group": base + "\nPLOTTING\n{\n  colour = blue\n}\n",
        "unknown-bunch-key": sub("bunching-factor             = 0.01", "bunching-factor             = 0.01\n    emittance                   = 1.0"),
        "unknown-fel-output-key": sub("normalized-frequency        = 1.00", "normalized-frequency        = 1.00\n    colour                      = blue"),
    }


def main():
    if not binding.have_reference():
        sys.exit("oracle/_ref/ref_dump is missing: run `make -C oracle ref` where /root/reference exists")
    out = {}
    work = tempfile.mkdtemp(prefix="golden-errors-")
    try:
        shutil.copy(os.path.join(ROOT, "tests", "jobs", "init-file-bunch.txt"), work)
        for name, text in cases().items():
            job = os.path.join(work, name + ".job")
            with open(job, "w") as f:
                f.write(text)
            r = subprocess.run([binding.REF_DUMP, job, os.path.join(work, "r"), "0", "--init-only"], cwd=work, stdout=subprocess.PIPE,
                               stderr=subprocess.STDOUT, text=True, errors="replace", timeout=300)
            out[name] = {"job": text, "exit_code": r.returncode, "message": last_message(r.stdout) if r.returncode else ""}
            print("%-36s exit %d  %s" % (name, r.returncode, out[name]["message"]))
    finally:
        shutil.rmtree(work, ignore_errors=True)
    json.dump(out, open(os.path.join(HERE, "parser-errors.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
