#!/usr/bin/env python
"""Dev helper: per CUDA source line share of warp instructions / stall samples / active lanes of one kernel in an .ncu-rep.
   python tools/ncu_lines.py <rep> <kernel regex> [instance]"""
import csv, subprocess, sys
from collections import defaultdict
rep, pat = sys.argv[1], sys.argv[2]
inst = int(sys.argv[3]) if len(sys.argv) > 3 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
# a kernel instance = a run of consecutive per-file blocks; take blocks until the file of the first block repeats
files = [rows[i - 2][1] if rows[i - 2] and rows[i - 2][0] in ("File Path", "File Name") else "?" for i in hdr]
starts = [k for k, f in enumerate(files) if f == files[0]]
lo = starts[inst]; hi = starts[inst + 1] if inst + 1 < len(starts) else len(hdr)
h = rows[hdr[lo]]
iI, iT, iS = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
agg = defaultdict(lambda: [0, 0, 0, 0]); src = {}
for k in range(lo, hi):
    end = hdr[k + 1] - 3 if k + 1 < len(hdr) else len(rows)
    cur = None
    for r in rows[hdr[k] + 1:end]:
        if len(r) < 10: continue
        if r[0].strip():
            cur = (files[k].split("/")[-1], int(r[0])); src[cur] = r[1]
        try:
            a = agg[cur]; a[0] += int(r[iI]); a[1] += int(r[iT]); a[2] += int(r[iS]); a[3] += 1
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values()); tt = sum(a[1] for a in agg.values())
print(f"warp-inst {tot}  avg lanes {tt / max(tot, 1):.2f}  samples {ts}")
for ln, a in sorted(agg.items()):
    if a[0] > 0.004 * tot or a[2] > 0.01 * ts:
        print(f"{ln[0][:14]:14s}{ln[1]:4d} sass {a[3]:4d} inst {100 * a[0] / tot:5.1f}% samp {100 * a[2] / max(ts, 1):5.1f}% lanes {a[1] / max(a[0], 1):5.1f} | {src[ln].strip()[:90]}")
