#!/usr/bin/env python
"""Debug aid: step the GPU and the oracle side by side on a fixture job, print the first step at which they part."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mithra_b200 import abi
from oracle import binding
from tests import helpers

job = sys.argv[1] if len(sys.argv) > 1 else "micro-trap"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
fused = "--fused" in sys.argv
p, meta, g = helpers.params_for(job)
gpu, cpu = abi.GpuSolver(p), binding.Oracle(p)
for s in (gpu, cpu):
    helpers.start_from_golden(s, g)
for step in range(nsteps):
    if fused:
        gpu.step(1)
    else:
        helpers.solve_step(gpu)
    helpers.solve_step(cpu)
    a, b = gpu.download_fields(("anp1", "an")), cpu.download_fields(("anp1", "an"))
    pg, pc = gpu.download_particles(), cpu.download_particles()
    ej, ea = helpers.rel_l2(a["anp1"], b["anp1"]), helpers.rel_l2(a["an"], b["an"])
    er, eg = helpers.rel_l2(pg[:, 1:4], pc[:, 1:4]), helpers.rel_l2(pg[:, 7:10], pc[:, 7:10])
    de = int((pg[:, 10] != pc[:, 10]).sum())
    w = int(np.abs(pg[:, 7:10] - pc[:, 7:10]).max(axis=1).argmax())
    print("step %3d  J %.2e  A %.2e  r %.2e  gb %.2e  e-flag diffs %d  worst particle %d gb gpu %s cpu %s" % (
        step, ej, ea, er, eg, de, w, pg[w, 7:10], pc[w, 7:10]), flush=True)
    if max(ej, ea, er, eg) > 1e-6:
        tb = cpu.lib.oracle_time_bunch(cpu.o)
        lz = p.gamma * (pc[w, 3] + p.beta * p.c0 * (tb + p.dt_shift))
        print("worst particle: lab z %.6e  x %.4e y %.4e" % (lz, pc[w, 1], pc[w, 2]))
        break
