#!/usr/bin/env python
"""Top SASS instructions of an `ncu --page source --csv` export by stall samples, with the dominant stall reason.

    python tools/ncu_src_top.py gpurun_out/r1b_src_push_particles.csv [N]
"""
import csv
import sys


def main(fn, n=40):
    rows = list(csv.reader(open(fn, errors="replace")))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    body = [r for r in rows[h + 1:] if len(r) == len(hdr)]
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stalls = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[si]) for r in body) or 1
    tinst = sum(int(r[ii]) for r in body)
    print("instructions %d, executed warp-instr %d, samples %d" % (len(body), tinst, tot))
    agg = {}
    for r in body:
        for i in stalls:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
    print("stall totals:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    ops = {}
    for r in body:
        op = r[1].split()[0] if r[1].split() else "?"
        if op.startswith("@"):
            op = r[1].split()[1]
        ops[op] = ops.get(op, 0) + int(r[ii])
    print("executed by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, tinst)) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
    order = sorted(range(len(body)), key=lambda k: -int(body[k][si]))[:n]
    for k in sorted(order):
        r = body[k]
        top = max(stalls, key=lambda i: int(r[i]))
        print("%5d %5.1f%% %9s  %-60s %s" % (k, 100.0 * int(r[si]) / tot, r[ii], r[1].strip()[:60], hdr[top]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
