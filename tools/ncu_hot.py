#!/usr/bin/env python
"""Hottest SASS instructions (warp stall samples) of one captured launch.  usage: ncu_hot.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
reasons = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
seen = set()
for i, r in enumerate(rows[2:]):
    if len(r) <= max(isamp, iex) or r[0] in seen:
        continue
    seen.add(r[0])
    try:
        s = int(r[isamp] or 0)
    except ValueError:
        continue
    rs = sorted(((int(r[j] or 0), h) for j, h in reasons), reverse=True)[:2]
    data.append((s, r[isrc].strip(), int(r[iex] or 0), i, rs))
tot = sum(d[0] for d in data)
print(f"# {rep}: {tot} samples over {len(data)} instructions; total instructions executed {sum(d[2] for d in data)}")
agg = {}
for j, h in reasons:
    agg[h] = 0
for i, r in enumerate(rows[2:]):
    if len(r) <= max(isamp, iex):
        continue
for s, src, ex, i, rs in sorted(data, reverse=True)[:top]:
    print(f"{s:7d} {100*s/tot:5.1f}%  ex={ex:9d}  #{i:5d}: {src[:70]:70s} {rs[0][1]}={rs[0][0]} {rs[1][1]}={rs[1][0]}")
