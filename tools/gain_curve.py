"""Sample a radiated-power file (rows: lab position of the plane in um, P) every 0.5 m and print log10 P at the given
positions in metres:  python tools/gain_curve.py power-0.txt gain_curve.txt 2 5 8 11 14"""
import sys

import numpy as np

r = np.loadtxt(sys.argv[1])
rows = []
for z in np.arange(0.5e6, r[-1, 0], 0.5e6):
    i = int(np.argmin(np.abs(r[:, 0] - z)))
    rows.append((r[i, 0], r[i, 1]))
np.savetxt(sys.argv[2], np.array(rows), header="lab position of the power plane (um), P")
print(r.shape, "first", r[0], "last", r[-1])
for z in sys.argv[3:]:
    i = int(np.argmin(np.abs(r[:, 0] - float(z) * 1e6)))
    print("z = %s m: P = %.6e  log10 = %.3f" % (z, r[i, 1], np.log10(max(r[i, 1], 1e-300))))
