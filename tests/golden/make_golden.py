#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/ by running the UNMODIFIED reference (oracle/_ref/ref_dump, built
from /root/reference/src by oracle/Makefile) on the small job files of tests/jobs/.

    python tests/golden/make_golden.py            # needs /root/reference (not available on the GPU box)

Every fixture <job>.npz holds, for one job:
    meta/<name>          every scalar / coefficient table Solver::initialize() produced (ref_dump meta record)
    p0, p50, p100        the particle list (n x 11: q, rnp, rnm, gb, e) at the start of field steps 0, 50, 100
    idx                  1024 node indices drawn with a fixed seed
    <arr>_s<step>        the values of array <arr> (an, anm1, jn [, fn, fnm1, rho]) at those nodes, steps 0 and 100
    <arr>_n<step>        (sum, sum of squares, max |.|) of the whole array
    ph99/...             intermediates of field step 99: A after fieldUpdate, particles after the push, E/B at the
                         nodes the reference evaluated (pic), J after the deposit -- same sampling
    power                pG per step (radiation.cpp:209-218), 100 rows
    screen<i>            the reference's screen text files parsed back to doubles (solver.cpp:2229-2252)
    pmap, vts/<file>     power-visualization jobs: the per-pixel map after the last step and the .vts files as written
    vtu/<file>           bunch-visualization jobs: the reference's .vtu / .pvtu files as written
    txt/<dir>/<file>     bunch-sampling / bunch-profile / field-sampling jobs: the reference's text files as written
The fixtures pin oracle/mithra_oracle.c (tests/test_oracle_golden.py) and, through it, the CUDA path.
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

JOBS = ("micro-nsfd", "micro-fd", "micro-o1", "micro-sc", "micro-seeded", "micro-optical", "micro-pviz",
        "micro-ics", "micro-lcls", "micro-trap", "micro-beams")
EXTRA_JOBS = ("micro-bsample", "micro-fsample", "micro-fline", "micro-bvtk", "micro-backshift", "micro-fviz", "micro-fall")        # host-writer fixtures only (same physics as micro-nsfd), not in tests/helpers.JOBS
NSTEPS = 100
NSAMPLE = 1024


def stats(a):
    a = np.asarray(a, dtype=np.float64)
    return np.array([a.sum(), (a * a).sum(), np.abs(a).max() if a.size else 0.0])


def sample(out, key, step, arr, idx, ncomp):
    a = arr.reshape(-1, ncomp)
    out["%s_s%s" % (key, step)] = a[idx].copy()
    out["%s_n%s" % (key, step)] = stats(a)


def make(job):
    work = tempfile.mkdtemp(prefix="golden-")
    try:
        prefix = os.path.join(work, "g")
        binding.run_ref_dump(os.path.join(ROOT, "tests", "jobs", job + ".job"), prefix, NSTEPS,
                             full_at=(0, 50, NSTEPS), phases_at=NSTEPS - 1, cwd=work)
        meta = binding.read_records(prefix + ".meta.bin")
        out = {"meta/" + k: v for k, v in meta.items()}
        nodes = int(meta["N0"][0]) * int(meta["N1"][0]) * int(meta["np"][0])
        rng = np.random.RandomState(20261017)
        idx = np.sort(rng.choice(nodes, size=min(NSAMPLE, nodes), replace=False)).astype(np.int64)
        out["idx"] = idx
        sc = int(meta["spaceCharge"][0]) == 1
        for step in (0, 50, NSTEPS):
            f = binding.read_records("%s.full%d.bin" % (prefix, step))
            out["p%d" % step] = f["particles"].reshape(-1, 11)
            out["t%d" % step] = np.array([f["time"][0], f["timeBunch"][0], float(f["nTime"][0])])
            if step == 50:
                continue
            for k in ("an", "anm1", "jn"):
                sample(out, k, step, f[k], idx, 3)
            if sc:
                for k in ("fn", "fnm1", "rho"):
                    sample(out, k, step, f[k], idx, 1)
        ph = binding.read_records("%s.phase%d.bin" % (prefix, NSTEPS - 1))
        sample(out, "ph99/anp1", "", ph["anp1_after_fieldUpdate"], idx, 3)
        sample(out, "ph99/jn", "", ph["jn_after_deposit"], idx, 3)
        if sc:
            sample(out, "ph99/fnp1", "", ph["fnp1_after_fieldUpdate"], idx, 1)
            sample(out, "ph99/rho", "", ph["rho_after_deposit"], idx, 1)
        out["ph99/particles"] = ph["particles_after_push"].reshape(-1, 11)
        pic = np.flatnonzero(ph["pic"])
        if pic.size > NSAMPLE:
            pic = np.sort(rng.choice(pic, size=NSAMPLE, replace=False))
        out["ph99/pic_idx"] = pic.astype(np.int64)
        out["ph99/pic_count"] = np.array([int(ph["pic"].sum())])
        out["ph99/en"] = ph["en"].reshape(-1, 3)[pic]
        out["ph99/bn"] = ph["bn"].reshape(-1, 3)[pic]
        prec = binding.read_records(prefix + ".power.bin")
        pw = prec["pG"]
        out["power"] = pw.reshape(NSTEPS, -1) if pw.size else np.zeros((0, 1))
        # power-visualization: the per-pixel map after the last step (rp_.pL) and the reference's own .vts files
        for k in prec:
            if k.startswith("pmap"):
                out["pmap"] = prec[k]
        for d in sorted(os.listdir(work)):
            dd = os.path.join(work, d)
            if os.path.isdir(dd):
                for fn in sorted(os.listdir(dd)):
                    if fn.endswith(".vtu") or fn.endswith(".pvtu"):
                        out["vtu/" + fn] = np.frombuffer(open(os.path.join(dd, fn), "rb").read(), dtype=np.uint8)
                    if fn.endswith(".pvts"):
                        out["vts/" + fn] = np.frombuffer(open(os.path.join(dd, fn), "rb").read(), dtype=np.uint8)
                    if fn.endswith(".vts"):
                        out["vts/" + fn] = np.frombuffer(open(os.path.join(dd, fn), "rb").read(), dtype=np.uint8)
                    # bunch-sampling / bunch-profile text files exactly as the reference wrote them
                    if d in ("bunch-sampling", "bunch-profile", "field-sampling", "field-profile") and fn.endswith(".txt"):
                        out["txt/%s/%s" % (d, fn)] = np.frombuffer(open(os.path.join(dd, fn), "rb").read(), dtype=np.uint8)
        scr_dir = os.path.join(work, "screens")
        if os.path.isdir(scr_dir):
            for fn in sorted(os.listdir(scr_dir)):
                i = int(fn.split("screen")[-1].split(".")[0])
                txt = open(os.path.join(scr_dir, fn)).read().split()
                out["screen%d" % i] = np.array([float(t) for t in txt]).reshape(-1, 6)
        np.savez_compressed(os.path.join(HERE, job + ".npz"), **out)
        print("%-14s nodes %7d particles %5d  |an|max %.3e  power[-1] %s  -> %.0f KB" % (
            job, nodes, out["p0"].shape[0], out["an_n%d" % NSTEPS][2], out["power"][-1] if len(out["power"]) else None,
            os.path.getsize(os.path.join(HERE, job + ".npz")) / 1024.0))
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    if not binding.have_reference():
        sys.exit("oracle/_ref/ref_dump is missing: run `make -C oracle ref` where /root/reference exists")
    for j in (sys.argv[1:] or JOBS + EXTRA_JOBS):
        make(j)
