import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = np.load(os.path.join(ROOT, "tests", "golden", "job-ir-mid.npz"))
r1 = g["power_1rank"]
big = r1[:, 1] > 1e-6 * r1[:, 1].max()
for gpus in (1, 2, 4):
    for env in ({}, {"MITHRA_NO_EB_SPLIT": "1"}, {"MITHRA_NO_EBMASK": "1"}, {"MITHRA_NO_OVERLAP": "1"}, {"MITHRA_STENCIL_PLAIN": "1"}):
        if gpus == 1 and env:
            continue
        d = tempfile.mkdtemp()
        r = subprocess.run([os.path.join(ROOT, "mithra_b200/host/mithra_b200"), os.path.join(ROOT, "tests/jobs/ir-mid.job"), "--gpus", str(gpus)],
                           cwd=d, env=dict(os.environ, **env), stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode:
            print(gpus, env, "FAILED", r.stdout.decode()[-300:]); continue
        got = np.loadtxt(os.path.join(d, "power-sampling", "power-ir-0.txt"))
        rel = np.abs(got[big, 1] - r1[big, 1]) / r1[big, 1]
        print(gpus, env, "max rel %.3e median %.3e argmax row %d" % (rel.max(), np.median(rel), np.flatnonzero(big)[rel.argmax()]), flush=True)
