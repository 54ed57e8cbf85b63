"""Write bench/<workload>.meta.npz: the scalars and coefficient tables Solver::initialize() derives for jobs/<workload>.job,
taken from the host binary's --dump-params record (mithra_b200/host/main.cpp; no GPU needed; bit-identical to the
unmodified reference's initialize() on every fixture job, tests/test_host.py).  The particle list and the raw parameter
block are dropped: bench.py builds its own synthetic bunch.

usage: python tools/make_bench_meta.py fel-lcls sc-weak
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mithra_b200 import meta  # noqa: E402

for name in sys.argv[1:]:
    with tempfile.TemporaryDirectory() as work:
        subprocess.check_call([os.path.join(ROOT, "mithra_b200", "host", "mithra_b200"), os.path.join(ROOT, "jobs", name + ".job"),
                               "--dump-params", os.path.join(work, "m")], cwd=work, stdout=subprocess.DEVNULL)
        rec = meta.read_records(os.path.join(work, "m.meta.bin"))
    rec = {k: v for k, v in rec.items() if k not in ("particles", "params0")}
    np.savez(os.path.join(ROOT, "bench", name + ".meta.npz"), **rec)
    print(name, int(rec["N0"][0]), int(rec["N1"][0]), int(rec["N2"][0]), "sc" if rec["spaceCharge"][0] else "")
