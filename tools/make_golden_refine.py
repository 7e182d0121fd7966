#!/usr/bin/env python
"""tests/golden/<case>/refined.fq.gz: the UNMODIFIED reference's second round (`bfc -R`, correct.c:438-442, 470,
517-546) over its own first-round output (corrected.fq.gz), counting from the original reads -- stdout of
    bfc -R -t1 <args> in.fq corrected.fq
Run in the build container only (needs oracle/_ref/bfc built from /root/reference)."""
import gzip, json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
exe = os.path.join(ROOT, "oracle", "_ref", "bfc")
for name in sorted(os.listdir(GOLD)):
    d = os.path.join(GOLD, name)
    if not os.path.isdir(d):
        continue
    meta = json.load(open(os.path.join(d, "case.json")))
    tmp = tempfile.mkdtemp()
    fq, c1 = os.path.join(tmp, "in.fq"), os.path.join(tmp, "c1.fq")
    open(fq, "wb").write(gzip.open(os.path.join(d, "in.fq.gz")).read())
    open(c1, "wb").write(gzip.open(os.path.join(d, "corrected.fq.gz")).read())
    args = ["-k", str(meta["k"]), "-b", str(meta["b"])] + meta["extra_args"]
    out = subprocess.run([exe, "-R", "-t", "1"] + args + [fq, c1], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    with open(os.path.join(d, "refined.fq.gz"), "wb") as raw:
        with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0, compresslevel=9) as fp:
            fp.write(out)
    # a second input that makes the reference re-run EVERY read: the tags are rewritten to max_heap = 77 (>= 50,
    # correct.c:544) and n_absent alternately 0 / 99, so both outcomes of the n_absent comparison (correct.c:438) occur
    import re
    cnt = [0]
    def retag(m):
        cnt[0] += 1
        return b"ec:Z:0_" + (b"0" if cnt[0] & 1 else b"99") + b":77_" + m.group(3)
    c2 = os.path.join(tmp, "c2.fq")
    forced = re.sub(rb"ec:Z:0_(\d+):(\d+)_(\S+)", retag, open(c1, "rb").read())
    open(c2, "wb").write(forced)
    out2 = subprocess.run([exe, "-R", "-t", "1"] + args + [fq, c2], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    for fn, data in (("refine_forced_in.fq.gz", forced), ("refined_forced.fq.gz", out2)):
        with open(os.path.join(d, fn), "wb") as raw:
            with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0, compresslevel=9) as fp:
                fp.write(data)
    rf2 = {}
    for y in out2.split(b"\n"):
        if y[:1] in b"@>" and b"ec:Z:0_" in y:
            rf2[y.rsplit(b"_", 1)[1]] = rf2.get(y.rsplit(b"_", 1)[1], 0) + 1
    print(name, "forced: rf codes", rf2, "lines differing from the forced input", sum(1 for x, y in zip(forced.split(b"\n"), out2.split(b"\n")) if x != y))
    a, b = open(c1, "rb").read().split(b"\n"), out.split(b"\n")
    hdr_changed = sum(1 for x, y in zip(a, b) if x != y and x[:1] in b"@>")
    rf = {}
    for y in b:
        if y[:1] in b"@>" and b"ec:Z:0_" in y:
            rf[y.rsplit(b"_", 1)[1]] = rf.get(y.rsplit(b"_", 1)[1], 0) + 1
    print(name, "records", sum(1 for y in b if y[:1] in b"@>"), "lines differing", sum(1 for x, y in zip(a, b) if x != y), "headers changed", hdr_changed, "rf codes", rf, "same length", len(a) == len(b))
