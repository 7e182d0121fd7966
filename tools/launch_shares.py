#!/usr/bin/env python
"""Per-kernel launch counts, total time and share of an ncu launch list (--metrics gpu__time_duration.sum --csv).
Usage: launch_shares.py launches.csv > shares.txt"""
import csv, collections, sys, re
tot = collections.Counter(); cnt = collections.Counter()
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    ms = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
    name = re.sub(r"<.*", "", r[ki]).replace("void ", "")
    tot[name] += ms; cnt[name] += 1
s = sum(tot.values())
for n, t in tot.most_common():
    print(f"{n:<60} launches {cnt[n]:5d}  total {t:10.3f} ms  share {100 * t / s:5.1f}%")
