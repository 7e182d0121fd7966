"""GPU (B200): the CUDA path, called through the C ABI (include/bfc_b200.h), against
 (a) the golden fixtures produced by the unmodified reference, and
 (b) the CPU oracle on seeded random inputs,
bit-exact: identical Bloom bytes, identical table entries (identity, cnt8, high6),
byte-identical corrected / trimmed FASTQ including the ec:Z: tag."""
import hashlib
import os

import numpy as np
import pytest

import orc
from golden_util import CASES, Case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bfc():
    import bfc_b200
    L = bfc_b200.lib()
    assert L.bfcg_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return bfc_b200


def as_bfc_opt(bfc, o):
    return bfc.make_opt(**{n: getattr(o, n) for n, _ in orc.Opt._fields_})


def mark_noqual(recs, qual, off):
    """Reads without quality carry 0xFF bytes in the flat layout."""
    if qual is None:
        return None
    q = qual.copy()
    for i, r in enumerate(recs):
        if r[3] is None:
            q[int(off[i]):int(off[i + 1]) - 1] = 0xFF
    return q


@pytest.mark.parametrize("path", ["part", "probe"])
@pytest.mark.parametrize("name", CASES)
def test_golden_count_correct(bfc, monkeypatch, name, path):
    monkeypatch.setenv("BFC_B200_COUNT", path)  # both count paths (see test_oracle_random)
    c = Case(name)
    qual = mark_noqual(c.recs, c.qual, c.off)
    e = bfc.Engine(as_bfc_opt(bfc, c.opt()))
    try:
        e.count(c.seq, qual, c.off)
        assert hashlib.sha256(e.bloom_bytes().tobytes()).hexdigest() == c.meta["bloom_sha256"]
        sub, key = e.table()
        assert len(key) == c.meta["table_n"] == e.n_distinct()
        assert np.array_equal(sub, c.sub) and np.array_equal(key, c.key)
        s, q, aux = e.correct(c.seq, qual, c.off)
        assert orc.format_corrected(c.recs, s, q, c.off, aux) == c.corrected
    finally:
        e.close()


@pytest.mark.parametrize("name", CASES)
def test_golden_trim(bfc, name):
    c = Case(name)
    qual = mark_noqual(c.recs, c.qual, c.off)
    e = bfc.Engine(as_bfc_opt(bfc, c.opt(filter_mode=1)))
    try:
        e.count(c.seq, qual, c.off)
        assert hashlib.sha256(e.bloom_bytes().tobytes()).hexdigest() == c.meta["bloom_sha256"]
        assert hashlib.sha256(e.bloom_bytes(high=True).tobytes()).hexdigest() == c.meta["bf_high_sha256"]
        keep, ts, te = e.trim(c.seq, c.off)
        assert orc.format_trimmed(c.recs, c.seq, c.qual, c.off, keep, ts, te) == c.trimmed
    finally:
        e.close()


def synth_batch(G, N, L, seed, err=0.01, repeat=0.0):
    from bfc_b200 import synth
    genome = synth.make_genome(G, seed, repeat)
    seq, qual = synth.make_reads(genome, N, L, seed, err)
    return synth.concat_batch(seq, qual)


@pytest.mark.parametrize("k,b,G,N,L,repeat,sub", [
    (21, 20, 30000, 20000, 100, 0.0, None),     # tiny filter: almost every occurrence conflicts
    (31, 26, 200000, 60000, 100, 0.2, None),
    (33, 30, 300000, 60000, 150, 0.3, 1 << 18),  # several sub-batches, few conflicts
    (55, 24, 100000, 30000, 150, 0.3, 1 << 17),
    (63, 22, 50000, 12000, 151, 0.1, 50000),
    (13, 24, 3000, 12000, 60, 0.0, None),        # block index wider than k: not a bit field of y0 => probe path either way
])
@pytest.mark.parametrize("path", ["part", "probe"])
def test_oracle_random(bfc, monkeypatch, k, b, G, N, L, repeat, sub, path):
    # both count paths: "part" = partitioned windows with the filter slices in shared memory (count_part.cu),
    # "probe" = probe / resolve / replay against the filter in HBM (count.cu; also what k < n_shift - 9 uses)
    monkeypatch.setenv("BFC_B200_COUNT", path)
    if sub:
        monkeypatch.setenv("BFC_B200_SUBBATCH", str(sub))
        monkeypatch.setenv("BFC_B200_COUNT_WINDOW", str(sub))
    seq, qual, off = synth_batch(G, N, L, seed=k * 1000 + b, repeat=repeat)
    o = orc.OracleRun(orc.make_opt(k=k, bf_shift=b))
    e = bfc.Engine(bfc.make_opt(k=k, bf_shift=b))
    try:
        # two count batches: state carries over exactly
        h = N // 2
        cut = int(off[h])
        o.count(seq[:cut], qual[:cut], off[:h + 1])
        o.count(seq[cut:], qual[cut:], off[h:] - off[h])
        e.count(seq[:cut], qual[:cut], off[:h + 1])
        e.count(seq[cut:], qual[cut:], off[h:] - off[h])
        assert np.array_equal(e.bloom_bytes(), o.bloom_bytes())
        assert int(e.stats.n_kmers) == int(o.stats[0]) and int(e.stats.n_pass) == int(o.stats[1])
        sub_o, key_o = o.table()
        sub_e, key_e = e.table()
        assert np.array_equal(sub_e, sub_o) and np.array_equal(key_e, key_o)
        mo, co, ho = o.hist()
        me, ce, he = e.hist()
        assert mo == me and np.array_equal(co, ce) and np.array_equal(ho, he)
        so, qo, ao, ctr = o.correct(seq, qual, off)
        se, qe, ae = e.correct(seq, qual, off)
        assert np.array_equal(ae, ao)
        assert np.array_equal(se, so) and np.array_equal(qe, qo)
        assert int(e.stats.n_lookups) > 0 and int(ctr[0]) > 0  # (the GPU batches and memoises lookups: the counts differ)
    finally:
        e.close()
        o.close()


def ragged_records(G, n, seed, max_len, err=0.015):
    """Reads of every length from 0 to max_len (half of them short, a few very long), some in lower case, some with
    runs of N, some without quality: what a uniform (N, L) batch never shows the kernels."""
    from bfc_b200 import synth
    rng = np.random.default_rng(seed)
    genome = synth.make_genome(G, seed, 0.2)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs, quals = [], []
    for i in range(n):
        r = rng.random()
        L = int(rng.integers(0, 80)) if r < 0.2 else int(rng.integers(80, 400)) if r < 0.9 else int(rng.integers(400, max_len + 1))
        L = min(L, G)
        st = int(rng.integers(0, G - L + 1))
        codes = genome[st:st + L].copy()
        if rng.random() < 0.5:
            codes = (3 - codes[::-1]).astype(np.uint8)
        e = rng.random(L) < err
        codes = np.where(e, (codes + rng.integers(1, 4, size=L, dtype=np.uint8)) & 3, codes).astype(np.uint8)
        s = lut[codes].copy()
        q = np.where(e & (rng.random(L) < 0.8), rng.integers(35, 45, size=L), rng.integers(53, 74, size=L)).astype(np.uint8)
        if L and rng.random() < 0.1:                       # a run of N somewhere
            a = int(rng.integers(0, L)); s[a:a + int(rng.integers(1, 6))] = ord("N")
        if rng.random() < 0.1:
            s = np.frombuffer(s.tobytes().lower(), dtype=np.uint8)
        seqs.append(s.tobytes())
        quals.append(None if rng.random() < 0.05 else q.tobytes())
    return seqs, quals


@pytest.mark.parametrize("k,b,max_len,path", [(21, 22, 3000, "part"), (33, 24, 3000, "part"), (55, 23, 5000, "part"), (31, 22, 1500, "probe")])
def test_oracle_ragged_reads(bfc, monkeypatch, k, b, max_len, path):
    """Count, correct and trim on ragged input (empty reads, reads shorter than k, reads of several thousand bases that
    span count segments and search scratch, N runs, lower case, reads without quality) against the oracle; small
    windows so that reads straddle window boundaries."""
    monkeypatch.setenv("BFC_B200_COUNT", path)
    monkeypatch.setenv("BFC_B200_COUNT_WINDOW", str(1 << 17))
    monkeypatch.setenv("BFC_B200_SUBBATCH", str(1 << 17))
    monkeypatch.setenv("BFC_B200_EC_BATCH", str(1 << 18))
    seqs, quals = ragged_records(60000, 4000, 100 * k + b, max_len)
    seq, qual, off = bfc.records_to_batch(seqs, quals)   # the C ABI marks a read without quality by 0xFF bytes,
    oq = qual.copy()                                     # the oracle by 0 bytes (oracle.h)
    oq[oq == 0xFF] = 0
    for fm in (0, 1):
        o = orc.OracleRun(orc.make_opt(k=k, bf_shift=b, filter_mode=fm))
        e = bfc.Engine(bfc.make_opt(k=k, bf_shift=b, filter_mode=fm))
        try:
            o.count(seq, oq, off)
            e.count(seq, qual, off)
            assert np.array_equal(e.bloom_bytes(), o.bloom_bytes())
            if fm:
                assert np.array_equal(e.bloom_bytes(high=True), o.bloom_bytes(high=True))
                ko, so, eo = o.trim(seq, off)
                ke, se, ee = e.trim(seq, off)
                assert np.array_equal(ke, ko) and np.array_equal(se, so) and np.array_equal(ee, eo)
            else:
                sub_o, key_o = o.table()
                sub_e, key_e = e.table()
                assert np.array_equal(sub_e, sub_o) and np.array_equal(key_e, key_o)
                so, qo, ao, _ = o.correct(seq, oq, off)
                se, qe, ae = e.correct(seq, qual, off)
                assert np.array_equal(ae, ao)
                qe = qe.copy()
                qe[qual == 0xFF] = 0                     # (a read without quality keeps its marker on both sides)
                assert np.array_equal(se, so) and np.array_equal(qe, qo)
        finally:
            e.close()
            o.close()


@pytest.mark.parametrize("k,b,H", [(31, 24, 4), (51, 25, 4), (27, 22, 9)])
def test_oracle_random_trim(bfc, k, b, H):
    seq, qual, off = synth_batch(100000, 30000, 120, seed=k + b)
    o = orc.OracleRun(orc.make_opt(k=k, bf_shift=b, n_hashes=H, filter_mode=1))
    e = bfc.Engine(bfc.make_opt(k=k, bf_shift=b, n_hashes=H, filter_mode=1))
    try:
        o.count(seq, qual, off)
        e.count(seq, qual, off)
        assert np.array_equal(e.bloom_bytes(), o.bloom_bytes())
        assert np.array_equal(e.bloom_bytes(high=True), o.bloom_bytes(high=True))
        ko, so, eo = o.trim(seq, off)
        ke, se, ee = e.trim(seq, off)
        assert np.array_equal(ke, ko) and np.array_equal(se, so) and np.array_equal(ee, eo)
    finally:
        e.close()
        o.close()


def test_small_search_scratch_is_redone_on_gpu(bfc, monkeypatch):
    """Searches whose list of edited bases outgrows the per-thread scratch are re-run by the same
    kernel with a larger one; results stay identical to the oracle."""
    monkeypatch.setenv("BFC_B200_EC_EDITS", "4")
    seq, qual, off = synth_batch(20000, 8000, 150, seed=99, err=0.03, repeat=0.5)
    o = orc.OracleRun(orc.make_opt(k=21, bf_shift=22))
    e = bfc.Engine(bfc.make_opt(k=21, bf_shift=22))
    try:
        o.count(seq, qual, off)
        e.count(seq, qual, off)
        so, qo, ao, ctr = o.correct(seq, qual, off)
        se, qe, ae = e.correct(seq, qual, off)
        assert np.array_equal(ae, ao) and np.array_equal(se, so) and np.array_equal(qe, qo)
        assert int(e.stats.n_redo) > 0
    finally:
        e.close()
        o.close()


def test_reference_api_single_item(bfc):
    """bfc_bf_insert/get, bfc_ch_insert/get/kmer_occ through one-element kernels (KATs from the reference)."""
    import ctypes as C
    from golden_util import kat
    L = bfc.lib()
    K = kat()
    for v in K["bloom"][:2]:
        bf = L.bfc_bf_init(v["n_shift"], v["n_hashes"])
        rets = [int(L.bfc_bf_insert(bf, h)) for h in v["hashes"][:120]]
        assert rets == v["insert_ret"][:120]
        L.bfc_bf_destroy(bf)
    assert not L.bfc_bf_init(8, 4)
    for v in K["table"][:4]:
        ch = L.bfc_ch_init(v["k"], v["l_pre"])
        for i, op in enumerate(v["ops"][:12]):
            y = (C.c_uint64 * 2)(*op["y"])
            assert L.bfc_ch_get(ch, y) == op["get_before"]
            for j in range(op["n_insert"]):
                assert L.bfc_ch_insert(ch, y, (i + j) & 1 if i != 7 else 1, 1) == 0
            assert L.bfc_ch_get(ch, y) == op["get_after"]
        L.bfc_ch_destroy(ch)


def test_cli_matches_reference_golden(bfc, tmp_path):
    """The drop-in `bfc` binary end to end (FASTQ file -> stdout) against the reference's stdout."""
    import gzip, subprocess
    exe = os.path.join(os.path.dirname(bfc.lib_path()), "bfc")
    for name in ("k31_edge", "k33_rep"):
        c = Case(name)
        fq = tmp_path / (name + ".fq")
        fq.write_bytes(c.fastq)
        args = ["-k", str(c.meta["k"]), "-b", str(c.meta["b"])] + c.meta["extra_args"]
        out = subprocess.run([exe] + args + ["-t", "4", str(fq)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
        assert out == c.corrected
        out = subprocess.run([exe] + args + ["-1", str(fq)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
        assert out == c.trimmed
        # dump -> restore -> correct gives the same output; the dump holds the reference's keys
        dump = tmp_path / (name + ".dump")
        subprocess.run([exe] + args + ["-E", "-d", str(dump), str(fq)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
        k, l_pre, sub, key = orc.parse_ref_dump(str(dump))
        assert k == c.meta["table_k"] and l_pre == c.meta["table_l_pre"]
        assert np.array_equal(sub, c.sub) and np.array_equal(key, c.key)
        out = subprocess.run([exe] + args + ["-r", str(dump), str(fq)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
        assert out == c.corrected


@pytest.mark.parametrize("keep_max,gz", [(None, False), (None, True), ("0", True), ("150000", False), ("100000000", False)])
def test_cli_many_small_batches_with_and_without_kept_blocks(bfc, tmp_path, keep_max, gz):
    """`-L` small enough for dozens of batches per pass: the four-step pipeline, the ring of pinned buffers and the
    blocks kept from the count pass (by default for gzip'd input only; BFC_B200_KEEP_MAX sets a budget for any file),
    read again (budget 0, or plain input) or given up half-way (a budget that overflows) all produce the reference's
    bytes."""
    import gzip, subprocess
    exe = os.path.join(os.path.dirname(bfc.lib_path()), "bfc")
    env = dict(os.environ)
    env.pop("BFC_B200_KEEP_MAX", None)
    if keep_max is not None:
        env["BFC_B200_KEEP_MAX"] = keep_max
    kept = (keep_max is None and gz) or keep_max == "100000000"
    for name in ("k31_edge", "k33_rep"):
        c = Case(name)
        fq = tmp_path / (name + (".fq.gz" if gz else ".fq"))
        fq.write_bytes(gzip.compress(c.fastq, 1) if gz else c.fastq)
        args = ["-k", str(c.meta["k"]), "-b", str(c.meta["b"])] + c.meta["extra_args"] + ["-L", "8000", "-V", "4"]
        r = subprocess.run([exe] + args + ["-t", "4", str(fq)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True, env=env)
        assert r.stdout == c.corrected
        assert r.stderr.count(b"batch counted") > 10
        assert (b"still in memory from the count pass" in r.stderr) == kept
        r = subprocess.run([exe] + args + ["-1", str(fq)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True, env=env)
        assert r.stdout == c.trimmed


@pytest.mark.parametrize("path", ["part", "probe"])
@pytest.mark.parametrize("b", [12, 20])
def test_table_grows_when_regions_fill(bfc, monkeypatch, b, path):
    """A saturated (too small) first filter lets nearly every occurrence through, so a window adds far more keys than
    the growth estimate reserved and whole table regions fill up: the inserts that found no room are applied again
    after the table has grown, however many there are (the reference only slows down here).  The parking list of the
    probe path is cut to 16 entries to make sure nothing depends on its size."""
    monkeypatch.setenv("BFC_B200_COUNT", path)
    monkeypatch.setenv("BFC_B200_TAB_DEFCAP", "16")
    monkeypatch.setenv("BFC_B200_COUNT_WINDOW", str(1 << 19))
    monkeypatch.setenv("BFC_B200_SUBBATCH", str(1 << 19))
    seq, qual, off = synth_batch(400000, 30000, 100, seed=b, repeat=0.0)
    o = orc.OracleRun(orc.make_opt(k=31, bf_shift=b))
    e = bfc.Engine(bfc.make_opt(k=31, bf_shift=b))
    try:
        o.count(seq, qual, off)
        e.count(seq, qual, off)
        assert np.array_equal(e.bloom_bytes(), o.bloom_bytes())
        assert int(e.stats.n_pass) == int(o.stats[1])
        so, ko = o.table()
        se, ke = e.table()
        assert len(ke) == e.n_distinct() > 100000
        assert np.array_equal(se, so) and np.array_equal(ke, ko)
    finally:
        e.close()
        o.close()


def test_reference_main_relinked(bfc, tmp_path):
    """INTEGRATION.md level 1 as a binary: the reference's own bfc.c (its main, getopt loop and phase orchestration,
    compiled against the reference's own bfc.h) linked with libbfc_b200.so in place of the reference's count / correct /
    bbf / htab / kthread / bseq objects (oracle/Makefile: _ref/bfc_relinked).  Its output is the reference's."""
    import subprocess
    exe = os.path.join(orc.ROOT, "oracle", "_ref", "bfc_relinked")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/bfc_relinked is not built on this box")
    for name in ("k31_edge", "k55_rep"):
        c = Case(name)
        fq = tmp_path / (name + ".fq")
        fq.write_bytes(c.fastq)
        args = ["-k", str(c.meta["k"]), "-b", str(c.meta["b"])] + c.meta["extra_args"]
        out = subprocess.run([exe] + args + ["-t", "4", str(fq)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
        assert out == c.corrected
        out = subprocess.run([exe] + args + ["-1", str(fq)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
        assert out == c.trimmed


@pytest.mark.parametrize("name", CASES)
def test_cli_discard_and_noqual_flags(bfc, tmp_path, name):
    """`-D` (drop reads with ec_code != 0, correct.c:598) and `-Q` (FASTA out, correct.c:596, 605-609) against digests of
    the unmodified reference's output (tools/make_golden_flags.py)."""
    import subprocess
    exe = os.path.join(os.path.dirname(bfc.lib_path()), "bfc")
    c = Case(name)
    fq = tmp_path / (name + ".fq")
    fq.write_bytes(c.fastq)
    args = ["-k", str(c.meta["k"]), "-b", str(c.meta["b"])] + c.meta["extra_args"]
    for key, flags in (("discard", ["-D"]), ("noqual", ["-Q"]), ("discard_noqual", ["-D", "-Q"])):
        out = subprocess.run([exe] + flags + args + ["-t", "3", str(fq)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
        assert len(out) == c.meta[key + "_bytes"] and hashlib.sha256(out).hexdigest() == c.meta[key + "_sha256"], key


@pytest.mark.parametrize("n,vb,begin,end", [(1, 1, 8, 28), (4095, 1, 8, 28), (4097, 1, 8, 28), (1_000_003, 1, 8, 28), (300_000, 8, 3, 36),
                                             (70_000, 2, 8, 19), (50_000, 4, 0, 7), (2_500_000, 1, 10, 31)])
def test_stable_radix_partition(bfc, monkeypatch, n, vb, begin, end):
    """csrc/partition.cuh on its own against torch's stable sort: ordered by the chosen key bits, ties in input order
    (stream order inside a Bloom block is what the exact count rests on), values travelling with their keys."""
    import torch
    import ctypes as C
    L = bfc.lib()
    g = torch.Generator(device="cuda").manual_seed(n + vb)
    key = torch.randint(0, 1 << 62, (n,), dtype=torch.int64, device="cuda", generator=g)
    if n > 1000:
        key[: n // 3] &= ~(((1 << (end - begin)) - 1) << begin) | (5 << begin)   # a long run of one digit
    vdt = {1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64}[vb]
    val = (torch.arange(n, device="cuda", dtype=torch.int64) * 2654435761 % 251).to(vdt)
    ko, vo = torch.empty_like(key), torch.empty_like(val)
    torch.cuda.synchronize()
    for cub in (False, True):
        if cub:
            monkeypatch.setenv("BFC_B200_CUB_SORT", "1")  # (read once per process: only checks the call path stays valid)
        assert L.bfcg_partition_records(C.c_void_p(key.data_ptr()), C.c_void_p(val.data_ptr()), C.c_void_p(ko.data_ptr()),
                                        C.c_void_p(vo.data_ptr()), n, vb, begin, end) == 0
        digit = (key >> begin) & ((1 << (end - begin)) - 1)
        order = torch.sort(digit, stable=True).indices
        assert torch.equal(ko, key[order]) and torch.equal(vo, val[order])


def test_filter_occupancy_telemetry(bfc):
    """bfcg_bf_load (a working version of the reference's unused bfc_bf_load, bbf.c:65-79) against numpy on the downloaded
    filter bytes, and bfcg_bf_suggest_shift against the load formula."""
    import ctypes as C
    seq, qual, off = synth_batch(50000, 12000, 100, seed=9, repeat=0.0)
    e = bfc.Engine(bfc.make_opt(k=31, bf_shift=22))
    L = bfc.lib()
    try:
        e.count(seq, qual, off)
        load, blocks, fp = C.c_double(), C.c_double(), C.c_double()
        assert L.bfcg_bf_load(e.bf, 1, C.byref(load), C.byref(blocks), C.byref(fp)) == 0
        b = e.bloom_bytes()
        bits = int(np.unpackbits(b).sum())
        n_blocks = len(b) // 64
        assert abs(load.value - bits / (n_blocks * 504)) < 1e-12
        assert abs(blocks.value - float((b.reshape(n_blocks, 64).any(axis=1)).mean())) < 1e-12
        assert abs(fp.value - load.value ** 4) < 1e-12 and 0 < fp.value < 1
        # 1e6 distinct k-mers, H = 4, p <= 1 %: load = 0.01^(1/4) = 0.316, m = -4e6 / ln(1 - 0.316) = 1.05e7 bits -> 2^24
        assert L.bfcg_bf_suggest_shift(1_000_000, 4, 0.01) == 24
        assert L.bfcg_bf_suggest_shift(10**12, 4, 0.01) == 37  # capped at BFC_MAX_BF_SHIFT
    finally:
        e.close()


def test_truncated_dump_is_rejected(bfc, tmp_path):
    """bfc_ch_restore on a dump cut short returns NULL (the reference asserts, htab.c:161-170) instead of a partial table."""
    seq, qual, off = synth_batch(20000, 4000, 100, seed=3, repeat=0.0)
    e = bfc.Engine(bfc.make_opt(k=31, bf_shift=22))
    L = bfc.lib()
    try:
        e.count(seq, qual, off)
        full = tmp_path / "full.dump"
        assert L.bfc_ch_dump(e.ch, str(full).encode()) == 0
        raw = full.read_bytes()
        ch = L.bfc_ch_restore(str(full).encode())
        assert ch and int(L.bfc_ch_count(ch)) == e.n_distinct()
        L.bfc_ch_destroy(ch)
        for cut in (4, len(raw) // 2, len(raw) - 8):
            part = tmp_path / f"cut{cut}.dump"
            part.write_bytes(raw[:cut])
            assert not L.bfc_ch_restore(str(part).encode())
    finally:
        e.close()


def test_reference_accepts_gpu_dump(bfc, tmp_path):
    """A table counted on the GPU and dumped by this library is read back by the UNMODIFIED reference: `bfc -r dump`
    corrects to the reference's own golden output, and hash2cnt (hash2cnt.c:37-64) reports the same number of keys."""
    import subprocess
    ref = os.path.join(orc.ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, "bfc")):
        pytest.skip("oracle/_ref is not built on this box")
    exe = os.path.join(os.path.dirname(bfc.lib_path()), "bfc")
    for name in ("k31_edge", "k33_rep"):
        c = Case(name)
        fq = tmp_path / (name + ".fq")
        fq.write_bytes(c.fastq)
        args = ["-k", str(c.meta["k"]), "-b", str(c.meta["b"])] + c.meta["extra_args"]
        dump = tmp_path / (name + ".dump")
        subprocess.run([exe] + args + ["-E", "-d", str(dump), str(fq)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
        out = subprocess.run([os.path.join(ref, "bfc")] + args + ["-t1", "-r", str(dump), str(fq)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
        assert out == c.corrected
        # hash2cnt -h: the count / high-count histograms of the dump = bfc_ch_hist of the reference's own table
        h = subprocess.run([os.path.join(ref, "hash2cnt"), "-h", str(dump)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
        rows = [ln.split() for ln in h.decode().splitlines()]
        cnt = np.array([int(r[1]) for r in rows], dtype=np.uint64)
        high = np.array([int(r[2]) for r in rows[:64]], dtype=np.uint64)
        assert np.array_equal(cnt, np.bincount((c.key & 0xff).astype(np.int64), minlength=256).astype(np.uint64))
        assert np.array_equal(high, np.bincount((c.key >> 8 & 0x3f).astype(np.int64), minlength=64).astype(np.uint64))
        # hash2cnt's default listing inverts the hash back to k-mer strings (k <= 37): one line per key
        lst = subprocess.run([os.path.join(ref, "hash2cnt"), str(dump)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
        assert lst.count(b"\n") == len(c.key)


@pytest.mark.parametrize("path", ["part", "probe"])
@pytest.mark.parametrize("world,k,b,trim", [(4, 31, 22, False), (2, 33, 20, True), (8, 55, 24, False), (2, 33, 30, False), (8, 33, 26, False)])
def test_sharded_count_kernels_emulated_on_one_gpu(bfc, monkeypatch, world, k, b, trim, path):
    """The sharded count path (bfcg_enum_records -> bucket exchange -> bfcg_count_records on 1/N filters) with
    the N ranks emulated one after the other on one GPU: the shards concatenate to the oracle's filter and the
    union of the shard tables is the oracle's table."""
    import ctypes as C
    import torch
    from bfc_b200.dist import CudaBackend, piece_bounds
    monkeypatch.setenv("BFC_B200_COUNT", path)
    monkeypatch.setenv("BFC_B200_SUBBATCH", str(1 << 17))
    monkeypatch.setenv("BFC_B200_COUNT_WINDOW", str(1 << 17))
    monkeypatch.setenv("BFC_B200_TAB_DEFCAP", "64")  # (round 1's N = 8 failure: a shard's keys fall into 1/N of the regions)
    seq, qual, off = synth_batch(60000, 16000, 120, seed=k + world, repeat=0.2)
    N = len(off) - 1
    opt = bfc.make_opt(k=k, bf_shift=b, filter_mode=1 if trim else 0)
    o = orc.OracleRun(orc.make_opt(k=k, bf_shift=b, filter_mode=1 if trim else 0))
    ranks = [CudaBackend(opt, world, 0, rank=r) for r in range(world)]
    try:
        o.count(seq, qual, off)
        chunk = 5000
        for lo in range(0, N, chunk):
            hi = min(N, lo + chunk)
            sent = []
            for r in range(world):  # every rank enumerates and buckets its piece
                p0, p1 = piece_bounds(lo, hi, r, world)
                s, q, f = seq[int(off[p0]):int(off[p1])], qual[int(off[p0]):int(off[p1])], off[p0:p1 + 1] - off[p0]
                y0, y1, counts = ranks[r].enum_records(bfc.api.host_batch(s, q, f), world)
                starts = np.concatenate([[0], np.cumsum(counts)])
                sent.append([(y0[int(starts[d]):int(starts[d + 1])].clone(), y1[int(starts[d]):int(starts[d + 1])].clone()) for d in range(world)])
            for d in range(world):  # the all-to-all: destination d receives the pieces in source-rank order
                r0 = torch.cat([sent[r][d][0] for r in range(world)])
                r1 = torch.cat([sent[r][d][1] for r in range(world)])
                ranks[d].count_record_runs(r0, r1, [int(sent[r][d][0].numel()) for r in range(world)], world)
        bloom = np.concatenate([ranks[r].bf_shard().cpu().numpy() for r in range(world)])
        assert np.array_equal(bloom, o.bloom_bytes())
        assert sum(int(r.stats.n_kmers) for r in ranks) == int(o.stats[0])
        assert sum(int(r.stats.n_pass) for r in ranks) == int(o.stats[1])
        if trim:
            high = np.concatenate([ranks[r].bf_high_shard().cpu().numpy() for r in range(world)])
            assert np.array_equal(high, o.bloom_bytes(high=True))
        else:
            parts = [ranks[r].export_table() for r in range(world)]
            ranks[0].import_table(parts)
            L = bfc.lib()
            n = int(L.bfcg_ch_export(ranks[0].full_ch, None, None))
            sub, key = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint64)
            L.bfcg_ch_export(ranks[0].full_ch, sub.ctypes.data_as(bfc.api.u32p), key.ctypes.data_as(bfc.api.u64p))
            so, ko = o.table()
            assert np.array_equal(sub, so) and np.array_equal(key, ko)
    finally:
        for r in ranks:
            r.close()
        o.close()


@pytest.mark.parametrize("k,b,trim", [(33, 26, False), (31, 22, True), (55, 24, False)])
def test_library_exchange_pipeline_single_rank(bfc, k, b, trim):
    """csrc/dist.cu with a world of one (what a 1-GPU box can run of it): the enumerate -> exchange -> cascade pipeline
    with its two buffer sets, NCCL send / receive to itself, several chunks in flight -- against the oracle."""
    from bfc_b200.dist import CudaBackend, NativeShardedCount
    seq, qual, off = synth_batch(60000, 15000, 120, seed=k + b, repeat=0.2)
    N = len(off) - 1
    opt = bfc.make_opt(k=k, bf_shift=b, filter_mode=1 if trim else 0)
    o = orc.OracleRun(orc.make_opt(k=k, bf_shift=b, filter_mode=1 if trim else 0))
    be = CudaBackend(opt, 1, 0, rank=0)
    try:
        sc = NativeShardedCount(be, 0, 1)
        o.count(seq, qual, off)
        empty = bfc.api.host_batch(seq[:0], qual[:0], off[:1])
        for i, lo in enumerate(range(0, N, 2500)):
            hi = min(N, lo + 2500)
            sc.count_piece(bfc.api.host_batch(seq[int(off[lo]):int(off[hi])], qual[int(off[lo]):int(off[hi])], off[lo:hi + 1] - off[lo]))
            if i == 2:
                sc.count_piece(empty)  # a rank that has run out of reads still takes part
        sc.gather()
        assert np.array_equal(be.bf_shard().cpu().numpy(), o.bloom_bytes())
        assert int(be.stats.n_kmers) == int(o.stats[0]) and int(be.stats.n_pass) == int(o.stats[1])
        if trim:
            assert np.array_equal(be.bf_high_shard().cpu().numpy(), o.bloom_bytes(high=True))
            keep, ts, te = np.zeros(N, dtype=np.uint8), np.zeros(N, dtype=np.int32), np.zeros(N, dtype=np.int32)
            be.trim_batch(bfc.api.host_batch(seq, None, off), keep.ctypes.data, ts.ctypes.data, te.ctypes.data)
            ko, tso, teo = o.trim(seq, off)
            assert np.array_equal(keep, ko) and np.array_equal(ts, tso) and np.array_equal(te, teo)
        else:
            L = bfc.lib()
            n = int(L.bfcg_ch_export(be.full_ch, None, None))
            sub, key = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint64)
            L.bfcg_ch_export(be.full_ch, sub.ctypes.data_as(bfc.api.u32p), key.ctypes.data_as(bfc.api.u64p))
            so, ko = o.table()
            assert np.array_equal(sub, so) and np.array_equal(key, ko)
            s2, q2 = seq.copy(), qual.copy()
            aux = np.zeros(2 * N, dtype=np.uint32)
            be.correct_batch(bfc.api.host_batch(s2, q2, off), aux.ctypes.data)
            s1, q1, a1, _ = o.correct(seq, qual, off)
            assert np.array_equal(aux, a1) and np.array_equal(s2, s1) and np.array_equal(q2, q1)
        st = sc.stats()
        assert st["received_records"] == int(o.stats[0]) and st["sent_records"] == 0
    finally:
        be.close()
        NativeShardedCount.finalize(bfc.lib())
        o.close()


@pytest.mark.parametrize("path", ["part", "probe", "part-wire-noext"])
def test_device_batches_and_many_windows(bfc, monkeypatch, path):
    """What bench.py runs: batches resident in HBM (BFCG_DEVICE), and host batches cut into many count /
    correct windows (the double-buffered copy streams cycle several times); both equal the oracle."""
    import ctypes as C
    if path == "part-wire-noext":  # the A/B switches: 16-byte records inside the count, no end-of-read lookup memo
        path = "part"
        monkeypatch.setenv("BFC_B200_COUNT_WIRE", "1")
        monkeypatch.setenv("BFC_B200_EC_NOEXT", "1")
    monkeypatch.setenv("BFC_B200_COUNT", path)
    monkeypatch.setenv("BFC_B200_EC_BATCH", "150000")
    monkeypatch.setenv("BFC_B200_COUNT_WINDOW", "65536")
    monkeypatch.setenv("BFC_B200_SUBBATCH", "65536")
    api, L = bfc.api, bfc.lib()
    seq, qual, off = synth_batch(40000, 12000, 150, seed=7, repeat=0.2)
    n, nb = len(off) - 1, int(off[-1])
    o = orc.OracleRun(orc.make_opt(k=33, bf_shift=26))
    eh = bfc.Engine(bfc.make_opt(k=33, bf_shift=26))
    ed = bfc.Engine(bfc.make_opt(k=33, bf_shift=26))
    d_seq, d_qual, d_off, d_aux = L.bfcg_dev_alloc(nb), L.bfcg_dev_alloc(nb), L.bfcg_dev_alloc(8 * (n + 1)), L.bfcg_dev_alloc(8 * n)
    try:
        o.count(seq, qual, off)
        so, qo, ao, _ = o.correct(seq, qual, off)
        # host batches, many windows
        eh.count(seq, qual, off)
        assert np.array_equal(eh.bloom_bytes(), o.bloom_bytes())
        sub_o, key_o = o.table()
        sub_e, key_e = eh.table()
        assert np.array_equal(sub_e, sub_o) and np.array_equal(key_e, key_o)
        se, qe, ae = eh.correct(seq, qual, off)
        assert np.array_equal(ae, ao) and np.array_equal(se, so) and np.array_equal(qe, qo)
        # device batches
        assert d_seq and d_qual and d_off and d_aux
        assert L.bfcg_h2d(d_seq, seq.ctypes.data, nb) == 0 and L.bfcg_h2d(d_qual, qual.ctypes.data, nb) == 0
        assert L.bfcg_h2d(d_off, off.ctypes.data, 8 * (n + 1)) == 0
        b = api.Batch()
        b.n_reads, b.n_bytes, b.where = n, nb, api.DEVICE
        b.off, b.seq, b.qual = C.cast(d_off, api.u64p), C.cast(d_seq, api.u8p), C.cast(d_qual, api.u8p)
        ed.count_batch(b)
        assert np.array_equal(ed.bloom_bytes(), o.bloom_bytes())
        sub_e, key_e = ed.table()
        assert np.array_equal(sub_e, sub_o) and np.array_equal(key_e, key_o)
        ed.correct_batch(b, d_aux)
        sd, qd, ad = np.empty_like(seq), np.empty_like(qual), np.empty(2 * n, dtype=np.uint32)
        assert L.bfcg_d2h(sd.ctypes.data, d_seq, nb) == 0 and L.bfcg_d2h(qd.ctypes.data, d_qual, nb) == 0
        assert L.bfcg_d2h(ad.ctypes.data, d_aux, 8 * n) == 0
        assert np.array_equal(ad, ao) and np.array_equal(sd, so) and np.array_equal(qd, qo)
    finally:
        for p in (d_seq, d_qual, d_off, d_aux):
            L.bfcg_dev_free(p)
        eh.close(); ed.close(); o.close()


@pytest.mark.parametrize("k,b,repeat", [(33, 24, 0.3), (21, 20, 0.5), (55, 24, 0.2)])
def test_refine_mode_against_oracle(bfc, k, b, repeat):
    """-R through the C ABI (aux carries the earlier stats in, the new ones out): bases taken back from the quality
    string, the n_absent comparison (rf_code 2 / 3), failures (rf_code 1) -- equal to the oracle, which is pinned to
    the reference's `bfc -R` by tests/test_oracle_golden.py."""
    seq, qual, off = synth_batch(40000, 9000, 120, seed=k + 3, err=0.02, repeat=repeat)
    n = len(off) - 1
    o1 = orc.OracleRun(orc.make_opt(k=k, bf_shift=b))
    o2 = orc.OracleRun(orc.make_opt(k=k, bf_shift=b, refine_ec=1))
    e1 = bfc.Engine(bfc.make_opt(k=k, bf_shift=b))
    e2 = bfc.Engine(bfc.make_opt(k=k, bf_shift=b, refine_ec=1))
    try:
        o1.count(seq, qual, off)
        e1.count(seq, qual, off)
        s1, q1, a1 = e1.correct(seq, qual, off)        # first round
        # earlier stats as a second round would read them from the tags; every other read claims 0 absent k-mers
        ori = a1.copy().reshape(-1, 2)
        ori[:, 1] = (ori[:, 1] & ~np.uint32(3 << 8)) | np.uint32(1 << 8)
        ori[::2, 1] &= np.uint32(0x3ff)
        o2.ch, keep_o = o1.ch, o2.ch
        e2.ch, keep_e = e1.ch, e2.ch
        so, qo, ao, _ = o2.correct(s1, q1, off, ori=ori)
        aux = ori.reshape(-1).copy()
        sg, qg = s1.copy(), q1.copy()
        e2.correct_batch(bfc.api.host_batch(sg, qg, off), aux.ctypes.data)
        assert np.array_equal(aux, ao) and np.array_equal(sg, so) and np.array_equal(qg, qo)
        rf = (aux.reshape(-1, 2)[:, 1] >> 8) & 3
        assert (rf == 2).any() and (rf == 3).any()
        assert n == len(rf)
    finally:
        o2.ch, e2.ch = keep_o, keep_e
        for x in (o1, o2, e1, e2):
            x.close()


def test_cli_refine_matches_reference_golden(bfc, tmp_path):
    """`bfc -R` end to end (tag parsing, reads left alone, sticky e->ori_st, fresh tags) against the reference's stdout."""
    import subprocess
    exe = os.path.join(os.path.dirname(bfc.lib_path()), "bfc")
    for name in ("k31_edge", "k33_rep", "k63_h7", "k27_opts"):
        c = Case(name)
        fq = tmp_path / (name + ".fq")
        fq.write_bytes(c.fastq)
        args = ["-k", str(c.meta["k"]), "-b", str(c.meta["b"])] + c.meta["extra_args"]
        for data, want in ((c.corrected, c.refined), (c.refine_forced_in, c.refined_forced)):
            c1 = tmp_path / (name + ".c1.fq")
            c1.write_bytes(data)
            out = subprocess.run([exe, "-R"] + args + ["-t", "3", str(fq), str(c1)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
            assert out == want
