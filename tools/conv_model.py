#!/usr/bin/env python
"""Per-conv roofline table from a bench.py JSON line: measured ms/step vs the tensor-pipe and HBM lower bounds."""
import json, re, sys
d = json.load(open(sys.argv[1]))
N = d["config"]["patches_per_step"]
PEAK_TF, PEAK_BW = 1369.1e12, 6549.4e9
rows = []
tot = [0, 0, 0]
nblk = {(64, 96): 3, (128, 48): 4, (256, 24): 6, (512, 12): 3}
for k, ms in d["conv_detail_ms_per_step"].items():
    m = re.match(r"conv(\d)x\d_tc\[(\d+)>(\d+) s(\d) (\d+)x(\d+)( [a-z+]+)?\]", k)
    ks, cin, cout, s, H, W = (int(m.group(i)) for i in range(1, 7))
    mode = (m.group(7) or "").strip()
    Ho, Wo = H // s, W // s
    launches = d["conv_detail_launches_per_step"][k] if "conv_detail_launches_per_step" in d else None
    rows.append((k, ms, ks, cin, cout, s, H, W, mode))
# count launches per step per key from ResNet structure
def count(ks, cin, cout, s, H, mode):
    c = 0
    Hc, inpl = 96, 64
    for planes, blocks, stride in ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)):
        for b in range(blocks):
            st = stride if b == 0 else 1
            Ho = Hc // st
            items = [(1, inpl, planes, 1, Hc, ""), (3, planes, planes, st, Hc, ""), (1, planes, planes * 4, 1, Ho, "stats"),
                     (1, planes, planes * 4, 1, Ho, "final+ds" if b == 0 else "final")]
            if b == 0:
                items.append((1, inpl, planes * 4, st, Hc, "stats"))
            c += sum(1 for it in items if it == (ks, cin, cout, s, H, mode))
            inpl, Hc = planes * 4, Ho
    return c
print(f"{'conv':46s} {'n':>2s} {'ms':>7s} {'TF/s':>6s} {'GB/s':>6s} {'t_tc':>6s} {'t_hbm':>6s} {'eff':>5s}")
for k, ms, ks, cin, cout, s, H, W, mode in rows:
    n = count(ks, cin, cout, s, H, mode)
    Ho, Wo = H // s, W // s
    px = Ho * Wo * N
    fl = 2.0 * px * cin * cout * ks * ks * n
    rd = H * W * N * cin * 2 if s == 1 else (H * W * N * cin * 2 if ks == 3 else px * cin * 2)
    wr = px * cout * 2 if mode in ("", "final", "final+ds") else 0
    if mode == "final": rd += px * cout * 2
    if mode == "final+ds":
        cin_ds = {256: 64, 512: 256, 1024: 512, 2048: 1024}[cout]
        fl += 2.0 * px * cin_ds * cout * n
        rd += px * cin_ds * 2
    by = (rd + wr) * n
    t_tc, t_hbm = fl / PEAK_TF * 1e3, by / PEAK_BW * 1e3
    tot[0] += ms; tot[1] += t_tc; tot[2] += t_hbm
    print(f"{k:46s} {n:2d} {ms:7.3f} {fl / ms / 1e9:6.0f} {by / ms / 1e6:6.0f} {t_tc:6.3f} {t_hbm:6.3f} {max(t_tc, t_hbm) / ms:5.2f}")
print("total", [round(x, 2) for x in tot])
