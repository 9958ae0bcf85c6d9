#!/usr/bin/env python
"""Dev helper: summarise the source page of one kernel of an .ncu-rep (per SASS chunk: share of warp instructions, stall samples, active lanes)."""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = next(r for r in rows if r and r[0] == "Address")
data = [r for r in rows if r and r[0].startswith("0x") and len(r) >= 10]
isrc, isamp, iinst, ithr = [h.index(x) for x in ["Source", "# Samples", "Instructions Executed", "Thread Instructions Executed"]]
ti = sum(int(r[iinst]) for r in data); tt = sum(int(r[ithr]) for r in data); ts = sum(int(r[isamp]) for r in data)
print(f"SASS {len(data)}  warp-inst {ti}  avg lanes {tt / max(ti,1):.2f}  samples {ts}")
for c in range(0, len(data), chunk):
    ch = data[c:c + chunk]
    ci = sum(int(r[iinst]) for r in ch); ct = sum(int(r[ithr]) for r in ch); cs = sum(int(r[isamp]) for r in ch)
    if ci > 0.01 * ti or cs > 0.01 * ts:
        print(f"{c:5d} inst {100 * ci / ti:5.1f}%  samples {100 * cs / ts:5.1f}%  lanes {ct / max(ci,1):5.1f}  {ch[0][isrc].strip()[:60]}")
