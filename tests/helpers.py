"""Shared helpers of the parity tests: golden fixtures, parameter blocks and the step loop of Solver::solve."""
import hashlib
import os

import numpy as np

from oracle import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
JOBS = ("micro-nsfd", "micro-fd", "micro-o1", "micro-sc", "micro-seeded", "micro-optical", "micro-pviz",
        "micro-ics", "micro-lcls", "micro-trap", "micro-beams")


def load_golden(job):
    g = np.load(os.path.join(GOLDEN, job + ".npz"))
    meta = {k[5:]: g[k] for k in g.files if k.startswith("meta/")}
    return meta, g


def params_for(job, max_particles=0):
    meta, g = load_golden(job)
    return binding.params_from_meta(meta, max_particles=max_particles), meta, g


def start_from_golden(solver, g, step=0):
    """Put a solver (oracle or GPU) into the reference's state at the start of field step 0."""
    assert step == 0
    t = g["t0"]
    solver.set_time(float(t[0]), float(t[1]), int(t[2]))
    solver.upload_particles(g["p0"])
    if solver.params.seed_enabled:
        solver.seedInitial()


def solve_step(s):
    """Body of the second while loop of Solver::solve, solver.cpp:1300-1399."""
    s.fieldUpdate()
    s.bunchUpdate()
    s.screenProfile()
    s.powerSample()
    s.powerVisualize()
    s.fieldShift()
    s.currentReset()
    s.currentUpdate()
    s.currentCommunicate()
    s.advanceTime()


def stats(a):
    a = np.asarray(a, dtype=np.float64)
    return np.array([a.sum(), (a * a).sum(), np.abs(a).max() if a.size else 0.0])


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    n = np.linalg.norm(b)
    return np.linalg.norm(a - b) / n if n > 0 else np.linalg.norm(a)


# --- digests of an initialize() dump (tools/check_shipped_jobs.py writes them, tests/test_host.py checks the host) ---
SKIP_KEYS = (".optical", ".signal")     # redundant views of und<i>.beam / .sig in ref_dump


def skipped_record(k, rec):
    """Records outside the comparison.  dtShift of a job WITHOUT an undulator (prj/FEL-ICS-TRAP/job-files/PAR-TRAP-*.job): the
    reference only sets Solver::dt_ when there is one (solver.cpp:315-323, 372) and its constructor leaves the member
    uninitialised (solver.h:283) -- ref_dump reads heap garbage there (9e-315 here); the host starts from 0."""
    return k in ("particles", "params0") or k.endswith(SKIP_KEYS) or (k == "dtShift" and int(rec["nUndulators"][0]) == 0)


def digest_meta(rec):
    """SHA-256 over the meta records in key order: name, dtype, shape, bytes.  `particles` and `params0` are not meta."""
    h = hashlib.sha256()
    n = 0
    for k in sorted(rec):
        if skipped_record(k, rec):
            continue
        v = np.ascontiguousarray(rec[k])
        h.update(k.encode()); h.update(str(v.dtype).encode()); h.update(str(v.shape).encode()); h.update(v.tobytes())
        n += 1
    return h.hexdigest(), n


def digest_particles(p):
    return hashlib.sha256(np.ascontiguousarray(p, dtype=np.float64).tobytes()).hexdigest()


def localised_job(job, workdir):
    """A copy of a shipped job file in `workdir` whose output directories under the author's cluster scratch
    (/cluster/scratch/afallahi/..., which neither program can create here: both print "Could not create the directory" and
    exit) point below the working directory instead.  Nothing else of the file changes."""
    txt = open(job, errors="replace").read().replace("/cluster/scratch/afallahi/", "./")
    out = os.path.join(workdir, os.path.basename(job))
    with open(out, "w") as f:
        f.write(txt)
    return out
