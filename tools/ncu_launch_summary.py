#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed] --csv` launch list by kernel:
count, total ms, share of device time and - when captured - DRAM traffic per launch, the DRAM bandwidth that
traffic means and the mean tensor-pipe activity.

    python tools/ncu_launch_summary.py gpurun_out/launches.csv [--json profiles/ncu_traffic.json] > profiles/rNN_launches_summary.md

--one-step keeps the first complete step of the list (period detected from the kernel-name sequence).
--json writes {kernel: {dram_bytes_per_launch, launches, ...}}: bench.py reads it for `roofline.traffic`.
"""
import collections
import csv
import json
import re
import sys

UNIT = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "ms": 1.0, "s": 1e3, "second": 1e3}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "bytes": 1.0}


def one_step(per_id):
    """Keep the launches of the first complete step: the list repeats with the step's period (a truncated capture would
    otherwise weight the kernel classes by wherever the profiler was cut off)."""
    ids = list(per_id)
    names = [per_id[i]["name"] for i in ids]
    for start in range(0, 8):                      # a few set-up launches may precede the first step
        for period in range(50, len(names) - start - 20):
            if names[start + period:start + period + 20] == names[start:start + 20]:
                return collections.OrderedDict((i, per_id[i]) for i in ids[start:start + period]), period
    return per_id, len(names)


def main(path, json_out=None, first_step=False):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per_id = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
        e = per_id.setdefault(row["ID"], {"name": name})
        m, u = row.get("Metric Name"), row.get("Metric Unit")
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        if m == "gpu__time_duration.sum":
            e["ms"] = v * UNIT.get(u, 1e-6)
        elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            e["dram"] = e.get("dram", 0.0) + v * BYTES.get(u, 1.0)
        elif m and m.startswith("sm__pipe_tensor_cycles_active"):
            e["tensor"] = v
    captured = len(per_id)
    if first_step:
        per_id, period = one_step(per_id)
    acc = collections.OrderedDict()
    for e in per_id.values():
        if "ms" not in e:
            continue
        a = acc.setdefault(e["name"], {"n": 0, "ms": 0.0, "dram": 0.0, "tensor_ms": 0.0, "has_dram": False, "has_tensor": False})
        a["n"] += 1
        a["ms"] += e["ms"]
        if "dram" in e:
            a["dram"] += e["dram"]
            a["has_dram"] = True
        if "tensor" in e:
            a["tensor_ms"] += e["tensor"] * e["ms"]
            a["has_tensor"] = True
    tot = sum(a["ms"] for a in acc.values())
    n = sum(a["n"] for a in acc.values())
    print(f"# ncu launch list summary: {path}\n")
    if first_step:
        print(f"(first complete step: {len(per_id)} of the {captured} captured launches)\n")
    print(f"{n} launches, {tot:.3f} ms of device time (per-launch times are cold-cache and serialised: compare SHARES)\n")
    print("| kernel | launches | total ms | share | DRAM MB / launch | DRAM GB/s | tensor pipe % (time-weighted) |\n|---|---:|---:|---:|---:|---:|---:|")
    for k, a in sorted(acc.items(), key=lambda kv: -kv[1]["ms"]):
        d = f"{a['dram'] / a['n'] / 1e6:.2f}" if a["has_dram"] else "-"
        bw = f"{a['dram'] / (a['ms'] * 1e-3) / 1e9:.0f}" if a["has_dram"] and a["ms"] > 0 else "-"
        tp = f"{a['tensor_ms'] / a['ms']:.1f}" if a["has_tensor"] and a["ms"] > 0 else "-"
        print(f"| `{k}` | {a['n']} | {a['ms']:.3f} | {a['ms'] / tot * 100:.1f}% | {d} | {bw} | {tp} |")


    if json_out:
        out = {"source": path, "note": "ncu per-launch means; dram = dram__bytes_read.sum + dram__bytes_write.sum",
               "kernels": {k: {"launches": a["n"], "dram_bytes_per_launch": round(a["dram"] / a["n"], 1) if a["has_dram"] else None,
                               "ms_per_launch_under_ncu": round(a["ms"] / a["n"], 5),
                               "tensor_pipe_pct": round(a["tensor_ms"] / a["ms"], 2) if a["has_tensor"] and a["ms"] > 0 else None}
                           for k, a in acc.items()}}
        with open(json_out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None, "--one-step" in sys.argv)
