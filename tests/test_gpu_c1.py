"""GPU (B200): parity at a BASELINE.json configuration size.  configs[0] -- 1 M x 100 bp synthetic reads from a 4.6 Mb
genome, `bfc -s 5m -k31 -t1` -- takes the reference about 4 minutes; tools/make_c1_digest.py ran the UNMODIFIED
reference on it once and committed the digests (tests/golden/c1_digest.json): sha256 of the first Bloom filter, of the
table's sorted entries, of the corrected FASTQ, of the `-1` trimmed FASTQ and its bf_high.  Here the same input is
regenerated (numpy, deterministic; its own sha256 is checked first) and pushed through the drop-in command line and
the C ABI; every digest must match."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu

DIGEST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_digest.json")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_config1_matches_reference_digests(tmp_path):
    import bfc_b200
    from bfc_b200 import synth
    ref = json.load(open(DIGEST))
    g = ref["generator"]
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else str(tmp_path)
    fq = os.path.join(shm, "bfc_b200_c1_%d.fq" % os.getpid())
    try:
        # the input, exactly as synth.write_fastq writes it (chunks of 200 000 reads), also kept as arrays for the C ABI
        genome = synth.make_genome(g["G"], g["seed"], 0.0)
        seqs, quals = [], []
        with open(fq, "wb") as fp:
            for done in range(0, g["N"], 200_000):
                n = min(200_000, g["N"] - done)
                s, q = synth.make_reads(genome, n, g["L"], g["seed"] + 7919 * (done // 200_000))
                fp.write(synth.fastq_bytes(s, q, first_index=done))
                seqs.append(s), quals.append(q)
        h = hashlib.sha256()
        with open(fq, "rb") as fp:
            for blk in iter(lambda: fp.read(1 << 24), b""):
                h.update(blk)
        assert h.hexdigest() == ref["input_sha256"], "the generator no longer reproduces the input the reference was run on"
        exe = os.path.join(os.path.dirname(bfc_b200.lib_path()), "bfc")
        args = ref["args"]
        # corrected FASTQ, byte for byte (ec:Z: tags included)
        out = subprocess.run([exe] + args + ["-t", "8", fq], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
        assert len(out) == ref["corrected_bytes"] and hashlib.sha256(out).hexdigest() == ref["corrected_sha256"]
        del out
        # the table, through the reference's dump format
        dump = os.path.join(shm, "bfc_b200_c1_%d.dump" % os.getpid())
        try:
            subprocess.run([exe] + args + ["-E", "-d", dump, fq], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
            k, l_pre, sub, key = orc.parse_ref_dump(dump)
        finally:
            if os.path.exists(dump):
                os.unlink(dump)
        assert (k, l_pre, len(key)) == (ref["table_k"], ref["table_l_pre"], ref["table_n"])
        assert sha(sub) == ref["table_sub_sha256"] and sha(key) == ref["table_key_sha256"]
        # trim mode
        tout = subprocess.run([exe] + args + ["-1", "-t", "8", fq], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
        assert len(tout) == ref["trimmed_bytes"] and hashlib.sha256(tout).hexdigest() == ref["trimmed_sha256"]
        del tout
        # the filters, through the C ABI (they never leave bfc_count, count.c:155)
        s, q, off = synth.concat_batch(np.concatenate(seqs), np.concatenate(quals))
        for fm, want in ((0, ref["bloom_sha256"]), (1, ref["bf_high_sha256"])):
            e = bfc_b200.Engine(bfc_b200.make_opt(k=ref["k"], bf_shift=ref["b"], filter_mode=fm))
            try:
                e.count(s, q, off)
                assert sha(e.bloom_bytes()) == ref["bloom_sha256"]
                if fm:
                    assert sha(e.bloom_bytes(high=True)) == want
            finally:
                e.close()
    finally:
        if os.path.exists(fq):
            os.unlink(fq)
