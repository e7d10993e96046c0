#!/usr/bin/env python
"""Opcode mix of one kernel in an .ncu-rep: ncu_mix.py <rep> <kernel-regex> [N]"""
import csv, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr)]
si, ei = hdr.index("Source"), hdr.index("Instructions Executed")
agg = collections.Counter()
for r in body:
    op = r[si].strip().split()
    if not op: continue
    o = op[0] if not op[0].startswith('@') else op[1]
    agg[o.split('.')[0]] += int(r[ei])
tot = sum(agg.values())
print("total warp instructions", tot)
for k, v in agg.most_common(N):
    print(f"  {k:12s} {v:12d} {100*v/tot:5.1f}%")
