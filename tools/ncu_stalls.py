#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: stall-reason totals, samples by opcode, hottest instructions.
usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_stalls.py [kernel-substring]"""
import csv
import sys
from collections import Counter

want = sys.argv[1] if len(sys.argv) > 1 else ""
rows = list(csv.reader(sys.stdin))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
f = lambda x: int(float(x)) if x not in ("", None) else 0  # noqa: E731
for s in sections:
    if want not in s["name"]:
        continue
    hdr, data = s["hdr"], s["data"]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    ci = [hdr.index(c) for c in cols]
    tot = sum(f(r[isamp]) for r in data)
    print("==", s["name"][:110])
    print("total samples", tot)
    agg = Counter()
    for r in data:
        for c, i in zip(cols, ci):
            agg[c] += f(r[i])
    print("stalls:", {k: v for k, v in agg.most_common() if v})
    byop = Counter()
    for r in data:
        t = r[isrc].split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        byop[op.split(".")[0]] += f(r[isamp])
    print("by opcode:", byop.most_common(12))
    for r in sorted(data, key=lambda r: -f(r[isamp]))[:18]:
        print("  ", r[0][-5:], r[isamp], r[iex], r[isrc][:48], {c: f(r[i]) for c, i in zip(cols, ci) if f(r[i]) > 0})
