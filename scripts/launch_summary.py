#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X): per kernel count, mean, share."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
acc = collections.OrderedDict()
for r in rows:
    acc.setdefault(r[4], []).append(float(r[14]) / 1e3)
tot = sum(sum(v) for v in acc.values())
print(f"# {len(rows)} launches, {tot:.1f} us total (cold-cache, serialised by ncu: shares are meaningful, absolutes are not)")
for k, v in acc.items():
    print(f"{k[-60:]:60s} n={len(v):3d} mean={sum(v)/len(v):9.1f} us share={100*sum(v)/tot:5.1f}%")
