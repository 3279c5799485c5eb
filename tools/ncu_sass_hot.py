#!/usr/bin/env python
"""Hot SASS instructions of one kernel launch from `ncu -i X.ncu-rep --page source --csv --launch-skip N --launch-count 1`:
top instructions by warp-stall samples with their dominant stall reasons, in program order (so the warp role they belong to can
be read off the neighbouring UTCHMMA / UTMALDG / LDTM / HMNMX2 instructions).   usage: ncu_sass_hot.py src.csv [top]"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1]))]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
ix = {h: i for i, h in enumerate(hdr)}
S, SRC, EX = ix["# Samples"], ix["Source"], ix["Instructions Executed"]
tot = sum(int(r[S] or 0) for r in data)
print("kernel:", rows[0][1][:90] if rows[0] else "?", "| total samples", tot, "| SASS instructions", len(data))
marks = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "HMNMX2", "UTCBAR", "ATOMS", "RED")
top = set(sorted(range(len(data)), key=lambda i: -int(data[i][S] or 0))[:top_n])
for i, r in enumerate(data):
    txt = r[SRC].strip()
    if i in top or any(m in txt for m in marks) and int(r[S] or 0) * 200 > tot:
        st = {h[6:]: int(r[ix[h]]) for h in hdr if h.startswith("stall_") and "Not Issued" not in h and r[ix[h]] not in ("", "0")}
        best = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{i:5d} {100.0 * int(r[S] or 0) / max(tot, 1):5.1f}%  exec {r[EX]:>9s}  {txt[:64]:64s} {best}")
