#!/usr/bin/env python
"""Every job file the reference ships (prj/*/job-files/*.job) through the UNMODIFIED reference's parser + Solver::initialize()
(oracle/_ref/ref_dump <job> <prefix> 0 --init-only) and through the host's (mithra_b200/host/mithra_b200 <job>
--dump-params): every scalar and coefficient table and the whole boosted particle list must agree bit for bit.

    python tools/check_shipped_jobs.py [--write] [job.job ...]      # needs /root/reference; run where it exists

--write records the reference side in tests/golden/shipped-jobs.json (per job: number of particles, SHA-256 of the meta
records and of the particle list, a few scalars for the reader), so that tests/test_host.py::test_shipped_job_files can
check the host against it without running the reference again.  TEST INFRASTRUCTURE (drives oracle/_ref).

The reference allocates the whole mesh in initialize(); jobs whose mesh does not fit this container's memory are listed as
"skipped" with the reason (address space capped with RLIMIT_AS so that a large job fails with bad_alloc, not the OOM killer)."""
import glob
import json
import os
import resource
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import binding              # noqa: E402
from mithra_b200 import meta as mmeta   # noqa: E402
from tests import helpers               # noqa: E402

PRJ = "/root/reference/prj"
EXE = os.path.join(ROOT, "mithra_b200", "host", "mithra_b200")
OUT = os.path.join(ROOT, "tests", "golden", "shipped-jobs.json")
digest_meta, digest_particles, skipped = helpers.digest_meta, helpers.digest_particles, helpers.skipped_record


def mem_cap():
    avail = 0
    for line in open("/proc/meminfo"):
        if line.startswith("MemAvailable"):
            avail = int(line.split()[1]) * 1024
    return int(avail * float(os.environ.get("SHIPPED_JOBS_MEM_FRACTION", "0.9")))


def run_reference(job, work):
    cap = mem_cap()

    def limit():
        resource.setrlimit(resource.RLIMIT_AS, (cap, cap))
    prefix = os.path.join(work, "r")
    r = subprocess.run([binding.REF_DUMP, job, prefix, "0", "--quiet", "--init-only"], cwd=work,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, preexec_fn=limit, text=True, errors="replace")
    if r.returncode != 0:
        return None, "reference exit %d: %s" % (r.returncode, (r.stdout or "").strip().splitlines()[-1:] or "")
    meta = binding.read_records(prefix + ".meta.bin")
    p = binding.read_records(prefix + ".full0.bin")["particles"].reshape(-1, 11)
    return (meta, p), ""


def run_host(job, work):
    prefix = os.path.join(work, "h")
    r = subprocess.run([EXE, job, "--dump-params", prefix], cwd=work, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, errors="replace")
    if r.returncode != 0:
        return None, "host exit %d: %s" % (r.returncode, (r.stdout or "").strip().splitlines()[-1:] or "")
    rec = mmeta.read_records(prefix + ".meta.bin")
    return (rec, rec["particles"].reshape(-1, 11)), ""


def why_random(job):
    """`generator = random` seeds rand() with the wall clock (classes.cpp:146-152): no two runs produce the same bunch."""
    txt = open(job, errors="replace").read()
    for line in txt.splitlines():
        line = line.split("#")[0]
        if "generator" in line and "random" in line:
            return "generator = random: the reference seeds rand() with time(NULL), classes.cpp:146-152"
    return ""


def key_of(job):
    return os.path.relpath(job, PRJ)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    write = "--write" in sys.argv
    jobs = [os.path.abspath(a) for a in args] or sorted(glob.glob(os.path.join(PRJ, "*", "job-files", "*.job")))
    if not args:
        # the meshes that take the reference minutes to allocate last
        jobs.sort(key=lambda j: ("FEL-ICS" in j or "FEL-LCLS" in j, j))
    table = json.load(open(OUT)) if (os.path.exists(OUT) and args) else {}
    bad = 0
    for job in jobs:
        work = tempfile.mkdtemp(prefix="shipped-")
        t0 = time.time()
        try:
            src, job = job, helpers.localised_job(job, work)
            host, why = run_host(job, work)
            if host is None:
                bad += 1
                print("%-50s HOST FAILED (%s)" % (key_of(src), why), flush=True)
                continue
            hm, hp = host
            t1 = time.time()
            # what the reference's initialize() allocates: A at three levels, E and B as floats, the pic flag (+ phi at three
            # levels and rho with space charge) -- solver.cpp:646-658, fdtdSC.cpp
            nodes = int(hm["N0"][0]) * int(hm["N1"][0]) * int(hm["np"][0])
            need = nodes * (72 + 20 + (32 if int(hm["spaceCharge"][0]) else 0)) + hp.size * 8 * 3          # measured: 92 bytes per node without space charge
            if need > mem_cap():
                why = "the reference's initialize() allocates %.0f GB for this mesh, more than this container has" % (need / 1e9)
                table[key_of(src)] = {"skipped": why}
                print("%-50s SKIPPED (%s)" % (key_of(src), why), flush=True)
                continue
            ref, why = run_reference(job, work)
            if ref is None:
                table[key_of(src)] = {"skipped": why}
                print("%-50s SKIPPED (%s)" % (key_of(src), why), flush=True)
                continue
            rm, rp = ref
            wrong = [k for k, v in rm.items() if not skipped(k, rm) and (k not in hm or not np.array_equal(np.asarray(hm[k]), np.asarray(v)))]
            same_p = hp.shape == rp.shape and np.array_equal(hp, rp)
            dm, nm = digest_meta(rm)
            entry = {"particles": int(rp.shape[0]), "meta_records": nm, "meta_sha256": dm, "particles_sha256": digest_particles(rp),
                     "N0": int(rm["N0"][0]), "N1": int(rm["N1"][0]), "N2": int(rm["N2"][0]), "space_charge": int(rm["spaceCharge"][0])}
            ok = not wrong and same_p
            table[key_of(src)] = entry
            if not ok and why_random(src):
                wrong = [k for k in wrong if k != "dtShift"]            # the time origin follows the head of the (random) bunch
                table[key_of(src)] = {"skipped": why_random(src)}
                print("%-50s SKIPPED (%s; scalars and tables %s)" % (key_of(src), why_random(src), "identical" if not wrong else "DIFFERENT: %s" % wrong[:6]), flush=True)
                bad += 1 if wrong else 0
                continue
            bad += 0 if ok else 1
            print("%-50s %s  %4d x %4d x %6d  particles %8d  (host %.1f s, reference %.1f s)%s" % (
                key_of(src), "identical" if ok else "DIFFERENT", entry["N0"], entry["N1"], entry["N2"], entry["particles"], t1 - t0, time.time() - t1,
                "" if ok else "  meta: %s  particles equal: %s" % (wrong[:6], same_p)), flush=True)
        finally:
            shutil.rmtree(work, ignore_errors=True)
            if write:
                json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)
    if write:
        json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)
        print("wrote", OUT)
    return 1 if bad else 0


if __name__ == "__main__":
    if not binding.have_reference() or not os.path.isdir(PRJ):
        sys.exit("needs oracle/_ref/ref_dump and /root/reference/prj")
    sys.exit(main())
