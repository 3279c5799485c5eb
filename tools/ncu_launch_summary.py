#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (count, total ms, share).

    python tools/ncu_launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches_summary.md
"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    acc = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        a = acc.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in acc.values())
    print(f"# ncu launch list summary: {path}\n")
    print(f"{n} launches, {tot:.3f} ms of device time (per-launch times are cold-cache and serialised: compare SHARES)\n")
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.3f} | {a[1] / tot * 100:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
