"""Shared helpers of the parity tests: golden fixtures, parameter blocks and the step loop of Solver::solve."""
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
