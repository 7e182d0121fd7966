import csv, io, subprocess, sys, collections
path = sys.argv[1]; kern = sys.argv[2]; pats = sys.argv[3:]
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
for (f,l),a in sorted(agg.items()):
    if any(p in a[3] for p in pats):
        print(f"{f}:{l} warp_inst {a[1]/1e6:9.1f}M thr_inst {a[2]/1e6:10.1f}M smp {a[0]:7d} | {a[3][:100]}")
