#!/usr/bin/env python
"""Per-launch A/B of the cta_group::2 convolution variant: two `ncu --metrics gpu__time_duration.sum -k regex:conv_tc_kernel --csv` launch
lists of the same bench command (BUSCA_CG2=0 / 1), compared launch by launch.  usage: cg2_compare.py l_0.csv l_1.csv"""
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    i = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[i]
    kn, v, u = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    out = []
    for r in rows[i + 1:]:
        if len(r) > v:
            t = float(r[v].replace(",", ""))
            t = t / 1000 if r[u].startswith("n") else (t * 1000 if r[u].startswith("m") else t)
            m = re.search(r"conv_tc_kernel<([^>]*)>", r[kn])
            out.append((m.group(1) if m else r[kn][:30], t))
    return out


a, b = load(sys.argv[1]), load(sys.argv[2])
groups = {}
for (ka, ta), (kb, tb) in zip(a, b):
    if ka != kb:
        g = groups.setdefault((ka, kb, "faster" if tb < ta else "slower"), [0, 0.0, 0.0])
        g[0] += 1
        g[1] += ta
        g[2] += tb
for (ka, kb, w), (n, ta, tb) in sorted(groups.items()):
    print(f"<{ka}> -> <{kb}>: {n} launches {w}: {ta:.0f} -> {tb:.0f} us (x{tb / ta:.3f})")
print("all switched launches:", round(sum(g[1] for g in groups.values())), "->", round(sum(g[2] for g in groups.values())), "us")
