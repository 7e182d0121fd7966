"""CPU: the oracle restatement (oracle/*.c) against the golden vectors produced by the
unmodified reference.  This is what pins the oracle (the reference ships no tests)."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import orc
from golden_util import CASES, Case, kat

CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def test_kat_kmer_hash(oracle):
    K = kat()
    g = K["sequence"]
    for v in K["kmers"]:
        k = v["k"]
        x = (C.c_uint64 * 4)(0, 0, 0, 0)
        for ch in g[v["start"]:v["start"] + k]:
            oracle.orc_kmer_append(k, x, CODE[ch])
        assert [int(t) for t in x] == v["x"]
        y = (C.c_uint64 * 2)()
        assert int(oracle.orc_kmer_hash(k, x, y)) == v["hash"]
        assert [int(t) for t in y] == v["y"]
    for v, w in zip(K["kmers"], K["change"]):
        x = (C.c_uint64 * 4)(*v["x"])
        oracle.orc_kmer_change(w["k"], x, w["d"], w["c"])
        assert [int(t) for t in x] == w["x"]
    for v in K["hash64"]:
        assert int(oracle.orc_hash_64(v["key"], (1 << v["k"]) - 1)) == v["hash"]


def test_kat_bloom(oracle):
    for v in kat()["bloom"]:
        bf = oracle.orc_bf_new(v["n_shift"], v["n_hashes"])
        rets = [int(oracle.orc_bf_insert(bf, h)) for h in v["hashes"]]
        assert rets == v["insert_ret"]
        gets = [int(oracle.orc_bf_get(bf, h ^ (i & 1))) for i, h in enumerate(v["hashes"][:100])]
        assert gets == v["get_ret"]
        b = np.ctypeslib.as_array(bf.contents.b, shape=(1 << (v["n_shift"] - 3),))
        assert hashlib.sha256(b.tobytes()).hexdigest() == v["bytes_sha256"]
        oracle.orc_bf_free(bf)
    assert not oracle.orc_bf_new(8, 4) and not oracle.orc_bf_new(56, 4)  # bbf.c:9


def test_kat_table(oracle):
    for v in kat()["table"]:
        ch = oracle.orc_ch_new(v["k"], v["l_pre"])
        for i, op in enumerate(v["ops"]):
            y = (C.c_uint64 * 2)(*op["y"])
            assert oracle.orc_ch_get(ch, y) == op["get_before"]
            for j in range(op["n_insert"]):
                oracle.orc_ch_insert(ch, y, (i + j) & 1 if i != 7 else 1)
            assert oracle.orc_ch_get(ch, y) == op["get_after"]
        cnt = (C.c_uint64 * 256)()
        high = (C.c_uint64 * 64)()
        assert oracle.orc_ch_hist(ch, cnt, high) == v["mode"]
        assert int(oracle.orc_ch_count(ch)) == v["count"]
        assert {str(i): int(c) for i, c in enumerate(cnt) if c} == v["cnt_hist"]
        assert {str(i): int(c) for i, c in enumerate(high) if c} == v["high_hist"]
        oracle.orc_ch_free(ch)


@pytest.mark.parametrize("name", CASES)
def test_count_correct_golden(oracle, name):
    c = Case(name)
    o = orc.OracleRun(c.opt())
    o.count(c.seq, c.qual, c.off)
    assert hashlib.sha256(o.bloom_bytes().tobytes()).hexdigest() == c.meta["bloom_sha256"]
    sub, key = o.table()
    assert np.array_equal(sub, c.sub) and np.array_equal(key, c.key)
    s, q, aux, _ = o.correct(c.seq, c.qual, c.off)
    assert orc.format_corrected(c.recs, s, q, c.off, aux) == c.corrected
    o.close()


@pytest.mark.parametrize("name", CASES)
def test_trim_golden(oracle, name):
    c = Case(name)
    o = orc.OracleRun(c.opt(filter_mode=1))
    o.count(c.seq, c.qual, c.off)
    assert hashlib.sha256(o.bloom_bytes().tobytes()).hexdigest() == c.meta["bloom_sha256"]
    assert hashlib.sha256(o.bloom_bytes(high=True).tobytes()).hexdigest() == c.meta["bf_high_sha256"]
    keep, ts, te = o.trim(c.seq, c.off)
    assert orc.format_trimmed(c.recs, c.seq, c.qual, c.off, keep, ts, te) == c.trimmed
    o.close()


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("which", ["refined", "refined_forced"])
def test_oracle_refine_mode(oracle, name, which):
    """`bfc -R` (second round over tagged reads): the oracle + the host rules of worker_ec's refine branch against
    the reference's stdout -- over its own first-round output, and over the same with every tag rewritten so that
    every read is re-run and both outcomes of the n_absent comparison occur (tools/make_golden_refine.py)."""
    c = Case(name)
    data = c.corrected if which == "refined" else c.refine_forced_in
    recs = orc.parse_fastx(data)
    o = orc.OracleRun(c.opt())
    r = orc.OracleRun(c.opt(refine_ec=1))
    try:
        o.count(c.seq, c.qual, c.off)                 # counted from the original reads
        r.ch, r_ch = o.ch, r.ch                       # the refine run corrects against that table
        comments, skip, ori = orc.refine_plan(recs)
        todo = [x for x, sk in zip(recs, skip) if not sk]
        seq, qual, off = orc.batch_from_records(todo)
        s, q, aux, _ = r.correct(seq, qual, off, ori=[v for v, sk in zip(ori, skip) if not sk]) if todo else (seq, qual, np.zeros(0, np.uint32), None)
        assert orc.format_refined(recs, comments, skip, s, q, off, aux) == (c.refined if which == "refined" else c.refined_forced)
    finally:
        r.ch = r_ch
        o.close(); r.close()
