import csv, io, subprocess, sys, collections
path = sys.argv[1]; kern = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr = None, None
agg = collections.defaultdict(lambda: [0,0,0,""])
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]
    elif r[0] == "Line No": hdr = r
    elif hdr and r[0] != "" and len(r) > 8:
        d = dict(zip(hdr[4:], r[4:]))
        try:
            a = agg[(fname, int(r[0]))]
            a[0] += int(d["# Samples"]); a[1] += int(d["Instructions Executed"]); a[2] += int(d["Thread Instructions Executed"]); a[3] = r[1].strip()
        except (ValueError, KeyError): pass
ts = sum(a[0] for a in agg.values()) or 1; ti = sum(a[1] for a in agg.values()) or 1
print("total samples", ts, "warp insts", ti, "thread insts", sum(a[2] for a in agg.values()))
byfile = collections.defaultdict(lambda:[0,0,0])
for (f,l),a in agg.items():
    byfile[f][0]+=a[0]; byfile[f][1]+=a[1]; byfile[f][2]+=a[2]
for f,a in byfile.items(): print(f"  {f}: smp {100*a[0]/ts:.1f}% inst {100*a[1]/ti:.1f}% thr/inst {a[2]/max(1,a[1]):.1f}")
print("-- by warp insts")
for (f,l),a in sorted(agg.items(), key=lambda kv: -kv[1][int(__import__("os").environ.get("SORTCOL","1"))])[:top]:
    print(f"{100*a[1]/ti:5.1f}% inst {100*a[0]/ts:5.1f}% smp thr/inst {a[2]/max(1,a[1]):4.1f} {f}:{l}: {a[3][:120]}")
