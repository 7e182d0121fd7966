"""Loading of the committed golden fixtures (tests/golden/, produced by tools/make_golden.py
from the unmodified reference)."""
import gzip
import json
import os

import numpy as np

import orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(d for d in os.listdir(GOLD) if os.path.isdir(os.path.join(GOLD, d)))


class Case:
    def __init__(self, name):
        d = os.path.join(GOLD, name)
        self.name = name
        self.meta = json.load(open(os.path.join(d, "case.json")))
        self.fastq = gzip.open(os.path.join(d, "in.fq.gz")).read()
        self.corrected = gzip.open(os.path.join(d, "corrected.fq.gz")).read()
        self.trimmed = gzip.open(os.path.join(d, "trimmed.fq.gz")).read()
        self.refined = gzip.open(os.path.join(d, "refined.fq.gz")).read()                 # `bfc -R` over corrected
        self.refine_forced_in = gzip.open(os.path.join(d, "refine_forced_in.fq.gz")).read()  # tags rewritten: every read re-run
        self.refined_forced = gzip.open(os.path.join(d, "refined_forced.fq.gz")).read()
        t = np.load(os.path.join(d, "table.npz"))
        self.sub, self.key = t["sub"], t["key"]
        self.recs = orc.parse_fastx(self.fastq)
        self.seq, self.qual, self.off = orc.batch_from_records(self.recs)

    def opt(self, **kw):
        m = self.meta
        o = dict(k=m["k"], bf_shift=m["b"])
        ea = m["extra_args"]
        for i in range(0, len(ea), 2):
            o[{"-H": "n_hashes", "-c": "min_cov", "-q": "q", "-w": "win_multi_ec"}[ea[i]]] = int(ea[i + 1])
        o.update(kw)
        return orc.make_opt(**o)


def kat():
    return json.load(open(os.path.join(GOLD, "kat.json")))
