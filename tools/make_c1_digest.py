#!/usr/bin/env python
"""tests/golden/c1_digest.json: digests of the UNMODIFIED reference on BASELINE.json configs[0] -- 1 M x 100 bp
synthetic reads from a 4.6 Mb genome, `bfc -s 5m -k31 -t1` (=> k = 31, Bloom 2^30 bits) -- too large to commit as
files: sha256 of the first Bloom filter's bytes, of the table's sorted (sub, key) arrays (`-E -d` dump), of the
corrected FASTQ, and of the `-1` trimmed FASTQ with its bf_high.  The input is regenerated from bfc_b200/synth.py
(numpy, deterministic) wherever the test runs.  Run in the build container only (needs oracle/_ref built from
/root/reference); takes about 6 minutes."""
import hashlib, json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc
from bfc_b200 import synth

GEN = dict(G=4_600_000, N=1_000_000, L=100, seed=1)
ARGS = ["-s", "5m", "-k", "31"]

def sha(a): return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()

if __name__ == "__main__":
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fq = os.path.join(tmp, "c1.fq")
    synth.write_fastq(fq, GEN["G"], GEN["N"], GEN["L"], GEN["seed"])
    meta = dict(generator=GEN, args=ARGS, k=31, b=30, input_sha256=hashlib.sha256(open(fq, "rb").read()).hexdigest())
    t0 = time.time()
    pre = os.path.join(tmp, "bloom")
    out = orc.ref_run(ARGS + ["-t1", fq], binary="bfc_bloomdump", env={"BFC_REF_BLOOM_DUMP": pre})
    meta["corrected_sha256"], meta["corrected_bytes"] = hashlib.sha256(out).hexdigest(), len(out)
    _, _, bb = orc.read_bloom_dump(pre + ".0")
    meta["bloom_sha256"] = sha(bb)
    print("correct done", time.time() - t0, flush=True)
    dump = os.path.join(tmp, "dump")
    orc.ref_run(ARGS + ["-t1", "-E", "-d", dump, fq])
    rk, rl, sub, key = orc.parse_ref_dump(dump)
    meta.update(table_k=rk, table_l_pre=rl, table_n=int(len(key)), table_sub_sha256=sha(sub), table_key_sha256=sha(key))
    print("dump done", time.time() - t0, flush=True)
    pre = os.path.join(tmp, "tbloom")
    tout = orc.ref_run(ARGS + ["-1", "-t1", fq], binary="bfc_bloomdump", env={"BFC_REF_BLOOM_DUMP": pre})
    _, _, t1 = orc.read_bloom_dump(pre + ".1")
    meta["trimmed_sha256"], meta["trimmed_bytes"], meta["bf_high_sha256"] = hashlib.sha256(tout).hexdigest(), len(tout), sha(t1)
    meta["reference_seconds"] = round(time.time() - t0, 1)
    json.dump(meta, open(os.path.join(ROOT, "tests", "golden", "c1_digest.json"), "w"), indent=1, sort_keys=True)
    print(meta)
    import shutil; shutil.rmtree(tmp, ignore_errors=True)
