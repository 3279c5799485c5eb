#!/usr/bin/env python
"""Total device time per kernel of two ncu time-only launch lists of the same command (A/B of an environment switch that does not change
the kernel names).  usage: launch_totals.py a.csv b.csv"""
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    i = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[i]
    kn, v, u = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    out = {}
    for r in rows[i + 1:]:
        if len(r) > v:
            t = float(r[v].replace(",", ""))
            t = t / 1000 if r[u].startswith("n") else (t * 1000 if r[u].startswith("m") else t)
            m = re.search(r"(\w+_kernel<[^>]*>|\w+_kernel)", r[kn])
            k = m.group(1) if m else r[kn][:40]
            a = out.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += t
    return out


a, b = load(sys.argv[1]), load(sys.argv[2])
for k in sorted(set(a) | set(b)):
    na, ta = a.get(k, [0, 0.0])
    nb, tb = b.get(k, [0, 0.0])
    print(f"{k}: {na} launches {ta:.0f} us -> {nb} launches {tb:.0f} us" + (f" (x{tb / ta:.3f})" if ta and na == nb else ""))
