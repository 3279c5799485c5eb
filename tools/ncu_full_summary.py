#!/usr/bin/env python
"""Per-kernel digest of an `ncu --set full ... --page raw --csv` export: for each kernel instantiation the launch with the longest
duration, with the figures the roofline discussion uses: duration, DRAM bytes and GB/s (vs. the measured 6549 GB/s copy peak), L2 -> SM
traffic, tensor-pipe activity, issue-slot utilisation, shared-memory wavefronts, achieved occupancy, registers / shared memory.
usage: python tools/ncu_full_summary.py full_raw.csv > profiles/rNN_ncu_full_summary.md"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def col(name):
    return ix.get(name)


def val(r, name, scale=1.0):
    i = col(name)
    if i is None or r[i] in ("", "n/a"):
        return None
    try:
        v = float(r[i].replace(",", ""))
    except ValueError:
        return None
    u = units[i]
    mult = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    return v * mult * scale


best = {}
for r in data:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("<unnamed>::", "").replace("void ", "")
    d = val(r, "gpu__time_duration.sum")
    if d is None:
        continue
    if name not in best or d > best[name][0]:
        best[name] = (d, r)
print("# ncu --set full digest (one launch per kernel instantiation: the longest captured)\n")
print("| kernel | us | DRAM MB | DRAM GB/s | of 6549 | L2->L1 MB | tensor pipe % | issue slots % | smem wavefronts M | occupancy % | regs | dyn smem KB |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for name, (d, r) in sorted(best.items(), key=lambda kv: -kv[1][0]):
    dram = (val(r, "dram__bytes_read.sum") or 0.0) + (val(r, "dram__bytes_write.sum") or 0.0)
    gbs = dram / (d * 1e-6) / 1e9 if d else 0.0
    l2 = val(r, "lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_ld.sum") or val(r, "l1tex__m_xbar2l1tex_read_bytes.sum")
    tp = val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") or val(r, "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active")
    iss = val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") or val(r, "sm__inst_issued.avg.pct_of_peak_sustained_active")
    wf = val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
    occ = val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    regs = val(r, "launch__registers_per_thread")
    sm = val(r, "launch__shared_mem_per_block_dynamic")
    f = lambda v, p=1: "" if v is None else f"{v:.{p}f}"
    print(f"| `{name}` | {d:.1f} | {dram / 1e6:.1f} | {gbs:.0f} | {gbs / 6549:.2f} | {f(l2 / 1e6 if l2 else None)} | {f(tp)} | {f(iss)} | {f(wf / 1e6 if wf else None, 2)} | {f(occ)} | {f(regs, 0)} | {f(sm / 1e3 if sm else None)} |")
