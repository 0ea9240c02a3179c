#!/usr/bin/env python3
"""Per-CUDA-source-line summary of an ncu report: ncu_lines.py <rep> <kernel-id-filter e.g. :::3> [min_pct]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", kid],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[h]
ix = {}
for i, k in enumerate(hdr):
    ix.setdefault(k, i)
lines = []
for r in rows[h + 1:]:
    if len(r) < len(hdr) or not r[0].isdigit():
        continue
    f = lambda k: float(r[ix[k]] or 0)
    try:
        f("Instructions Executed"); f("# Samples"); f("stall_not_selected")
    except ValueError:
        continue
    lines.append((int(r[0]), r[1], f("Instructions Executed"), f("# Samples"), f("L1 Wavefronts Shared"), f("L1 Wavefronts Shared Excessive"),
                  f("stall_long_sb"), f("stall_short_sb"), f("stall_mio"), f("stall_barrier"), f("stall_math"), f("stall_wait"), f("stall_not_selected")))
ti = sum(l[2] for l in lines); ts = sum(l[3] for l in lines); tw = sum(l[4] for l in lines)
print(f"total inst {ti:.3e} samples {ts:.0f} smem wavefronts {tw:.3e}")
print(" line  inst%  samp%  wf%   exc_wf%  | long short mio barrier math wait notsel | source")
for l in lines:
    if l[2] / ti * 100 >= minp or l[3] / ts * 100 >= minp:
        print(f"{l[0]:5d} {l[2]/ti*100:5.1f} {l[3]/ts*100:6.1f} {l[4]/max(tw,1)*100:5.1f} {l[5]/max(tw,1)*100:6.1f}   | "
              f"{l[6]/ts*100:4.1f} {l[7]/ts*100:4.1f} {l[8]/ts*100:4.1f} {l[9]/ts*100:4.1f} {l[10]/ts*100:4.1f} {l[11]/ts*100:4.1f} {l[12]/ts*100:4.1f} | {l[1][:90]}")
