#!/usr/bin/env python
"""One-screen summary of a bench.py JSON line (ms/frame, roofline fractions, per-class and per-conv times)."""
import json
import sys

try:
    d = json.load(open(sys.argv[1]))
except Exception as e:  # noqa: BLE001
    print("bad json", e)
    print(open(sys.argv[1]).read()[:500])
    sys.exit(0)
r = d.get("roofline") or {}
print("ms/frame", d.get("ms_per_frame"), "value", d.get("value"), "e2e", (d.get("e2e") or {}).get("value"), "p50/p99", d.get("p50_frame_latency_ms"),
      d.get("p99_frame_latency_ms"))
print("roofline", r.get("kernel"), r.get("frac"), "step_frac", r.get("step_frac"), "hbm", r.get("step_hbm_frac"), "kept", d.get("kept_tracks"),
      "parity", (d.get("parity_vs_reference_golden") or {}).get("decisions_equal"), (d.get("parity_vs_reference_golden") or {}).get("max_abs_dprob"))
ks = d.get("kernels") or {}
print({k: round(v["ms_per_frame"], 3) for k, v in ks.items() if v["ms_per_frame"] > 0.1})
print("sum of kernel ms/frame", round(sum(v["ms_per_frame"] for v in ks.values()), 3), "launches/frame", round(sum(v["launches_per_frame"] for v in ks.values()), 1))
if "-v" in sys.argv:
    for k, v in (d.get("conv_detail_ms_per_frame") or {}).items():
        print(f"   {k:45s} {v:7.3f}")
