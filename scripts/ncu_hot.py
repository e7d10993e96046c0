#!/usr/bin/env python
"""Top stall sites (SASS) of one kernel in an .ncu-rep:  ncu_hot.py <rep> <kernel-regex> [N]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr)]
si, ai, ni, ei = hdr.index("Source"), hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ni]) for r in body)
print("total samples", tot, "instructions", len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ni]))[:N]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[c]), hdr[c][6:]) for c in stall_cols if int(r[c]) > 0), reverse=True)[:3]
    print(f"{i:5d} {int(r[ni]):6d} {100*int(r[ni])/tot:5.1f}% ex={r[ei]:>8s} {r[si].strip()[:70]:70s} {st}")
