#!/usr/bin/env python
"""Warp stall breakdown (cycles per issued instruction) of the kernels in an .ncu-rep raw CSV page.
    ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_stalls.py raw.csv [kernel substring]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = sys.argv[2] if len(sys.argv) > 2 else ""
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    if want not in r[ki]:
        continue
    out = []
    for h, v in zip(hdr, r):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                out.append((float(v.replace(",", "")), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print(r[ki][:90])
    for v, h in sorted(out, reverse=True)[:10]:
        print("   %7.3f %s" % (v, h))
