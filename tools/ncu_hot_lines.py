#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an .ncu-rep captured with --import-source on.
Usage: ncu_hot_lines.py file.ncu-rep [top_n]"""
import csv, io, subprocess, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0] != "" and len(r) > 8:
        d = dict(zip(hdr[4:], r[4:]))
        try:
            lines.append((fname, int(r[0]), r[1].strip(), int(d["# Samples"]), int(d["Instructions Executed"]),
                          int(d["Thread Instructions Executed"]), int(d.get("stall_long_sb", 0)), int(d.get("stall_no_inst", 0)),
                          int(d.get("L2 Theoretical Sectors Local", 0) or 0)))
        except (ValueError, KeyError):
            pass
tot_s = sum(l[3] for l in lines) or 1
tot_i = sum(l[4] for l in lines) or 1
print(f"total samples {tot_s}  warp insts {tot_i}  thread insts {sum(l[5] for l in lines)}")
print("-- by samples")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{100*l[3]/tot_s:5.1f}% smp {100*l[4]/tot_i:5.1f}% inst thr/inst {l[5]/max(1,l[4]):4.1f} longsb {l[6]:6d} noinst {l[7]:6d} local {l[8]:9d} {l[0]}:{l[1]}: {l[2][:110]}")
