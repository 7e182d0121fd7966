#!/usr/bin/env python
"""Adds to every tests/golden/<case>/case.json the digests of the UNMODIFIED reference's output under the two output
flags of the correction phase: `-D` (drop reads whose ec_code is not 0, correct.c:598) and `-Q` (no quality line:
FASTA out, correct.c:596, 605-609), and under both.  Run in the build container only (needs oracle/_ref/bfc)."""
import gzip, hashlib, json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
exe = os.path.join(ROOT, "oracle", "_ref", "bfc")
for name in sorted(os.listdir(GOLD)):
    d = os.path.join(GOLD, name)
    if not os.path.isdir(d):
        continue
    meta = json.load(open(os.path.join(d, "case.json")))
    tmp = tempfile.mkdtemp()
    fq = os.path.join(tmp, "in.fq")
    open(fq, "wb").write(gzip.open(os.path.join(d, "in.fq.gz")).read())
    args = ["-k", str(meta["k"]), "-b", str(meta["b"])] + meta["extra_args"]
    for key, flags in (("discard", ["-D"]), ("noqual", ["-Q"]), ("discard_noqual", ["-D", "-Q"])):
        out = subprocess.run([exe, "-t", "1"] + flags + args + [fq], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
        meta[key + "_sha256"], meta[key + "_bytes"] = hashlib.sha256(out).hexdigest(), len(out)
        meta[key + "_records"] = sum(1 for ln in out.split(b"\n") if ln[:1] in (b"@", b">") and b"ec:Z:" in ln)
    json.dump(meta, open(os.path.join(d, "case.json"), "w"), indent=1, sort_keys=True)
    print(name, {k: meta[k] for k in meta if k.endswith("_records")})
