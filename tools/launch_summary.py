#!/usr/bin/env python3
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel. usage: launch_summary.py <csv>"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    k = r[4].split('(')[0].replace('void ', '').replace('gpsiq::', '')[:56]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[-1])
tot = sum(a[1] for a in agg.values())
for k, a in agg.items():
    print(f"{k:58s} n={a[0]:4d} total={a[1]/1e3:9.1f}us avg={a[1]/a[0]/1e3:8.1f}us {100*a[1]/tot:5.1f}%")
print(f"total {tot/1e3:.1f} us")
