import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from tests import helpers
from tests.test_gpu_slabs import make_slabs, PHASES
order = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "01")]
p, meta, g = helpers.params_for("micro-nsfd")
parts = make_slabs(p, g, 2)
for step in range(2):
    for ph in PHASES:
        t0 = time.time()
        try:
            for r in order:
                getattr(parts[r], ph)()
            for s in parts:
                s.synchronize()
        except Exception as e:
            print("step", step, ph, "FAILED:", e); sys.exit(0)
        print("step", step, ph, "%.3f s" % (time.time() - t0), flush=True)
print("ok")
