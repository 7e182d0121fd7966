#!/usr/bin/env python
"""Per-stage durations from a `bfc -V 4` log (tools/cli_timeline.py): median over the batches."""
import re, sys, statistics
PAIRS = [("count_cb", "read begins", "count_cb", "block read and split", "count: read+split"),
         ("count_cb", "pack begins", "count_cb", "batch packed", "count: pack"),
         ("count_cb", "count begins", "count_cb", "batch counted", "count: GPU"),
         ("ec_cb", "read begins", "ec_cb", "block read and split", "correct: read+split"),
         ("ec_pack", "pack begins", "ec_pack", "batch packed", "correct: pack"),
         ("ec_cb", "batch taken", "ec_cb", "batch corrected", "correct: GPU"),
         ("ec_cb", "write begins", "ec_cb", "batch written", "correct: write")]
for fn in sys.argv[1:]:
    ev = []
    for l in open(fn):
        m = re.match(r"\[D::(\w+) @([\d.]+)\] (.*)", l)
        if m:
            ev.append((float(m.group(2)), m.group(1), m.group(3).strip()))
    print(fn)
    for fa, a, fb, b, label in PAIRS:
        ta = [t for t, f, w in ev if f == fa and w == a]
        tb = [t for t, f, w in ev if f == fb and w == b]
        d = [y - x for x, y in zip(ta, tb)]
        iv = [y - x for x, y in zip(tb, tb[1:])]
        if d:
            print(f"  {label:22s} n={len(d):3d}  median {1e3 * statistics.median(d):6.1f} ms  max {1e3 * max(d):6.1f} ms  interval {1e3 * statistics.median(iv) if iv else 0:6.1f} ms")
    st = {w: t for t, f, w in ev if f == "bfc_count"}
    print("  start-up:", {k: round(v, 3) for k, v in st.items()}, " last stamp:", ev[-1][0] if ev else None)
