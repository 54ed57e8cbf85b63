import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import helpers
meta, g = helpers.load_golden("micro-fall")
N0, N1, N2 = int(meta["N0"][0]), int(meta["N1"][0]), int(meta["N2"][0])
d = tempfile.mkdtemp()
subprocess.check_output([os.path.join(ROOT, "mithra_b200/host/mithra_b200"), os.path.join(ROOT, "tests/jobs/micro-fall.job"), "--steps", "100"], cwd=d)
for fn in ("all-p0-49.vts", "all-p0-98.vts"):
    ref = bytes(g["vts/" + fn]).decode().splitlines()
    got = open(os.path.join(d, "field-visualization", fn)).read().splitlines()
    def blocks(lines):
        out, cur = [], []
        for l in lines:
            if l.startswith("<"):
                if cur: out.append(np.array(cur)); cur = []
            else: cur.append([float(x) for x in l.split()])
        return out
    G, R = blocks(got)[1], blocks(ref)[1]
    scale = np.abs(R).max(axis=0)
    bad = np.abs(G - R) > 2e-4 * np.abs(R) + 2e-4 * scale
    plane = N0 * N1
    print(fn, "bad per column", bad.sum(axis=0), "scale", scale)
    for col in range(4):
        idx = np.flatnonzero(bad[:, col])
        if idx.size:
            k, r = idx // plane, idx % plane
            j, i = r // N0, r % N0
            print(" col", col, "planes", np.unique(k)[:10], "...", np.unique(k)[-5:], "i", np.unique(i), "j", np.unique(j))
            for t in idx[:4]: print("   row", t, "got", G[t], "ref", R[t])
