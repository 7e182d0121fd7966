"""CPU: the block reader / writer of the phase drivers (csrc/fqblock.c) against the tolerant record-at-a-time
reader (csrc/bseq.c, which mirrors the reference's bseq.c over kseq.h): same records on plain four-line FASTQ (the
parallel path) and on everything else (hand-over to the tolerant parser mid-stream), sticky comments included."""
import ctypes as C
import gzip
import os
import random

import pytest

import bfc_b200

u64p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)


class Bseq1(C.Structure):
    _fields_ = [("l_seq", C.c_int), ("aux", C.c_uint32), ("aux2", C.c_uint32), ("name", C.c_char_p),
                ("comment", C.c_char_p), ("seq", C.c_char_p), ("qual", C.c_char_p)]


class Block(C.Structure):
    _fields_ = [("buf", C.POINTER(C.c_char)), ("buf_len", C.c_size_t), ("n", C.c_int64),
                ("name_off", u64p), ("com_off", u64p), ("seq_off", u64p), ("qual_off", u64p),
                ("name_len", u32p), ("com_len", u32p), ("seq_len", u32p), ("n_bases", C.c_uint64), ("any_qual", C.c_int)]


NONE = (1 << 64) - 1


@pytest.fixture(autouse=True, params=[None, 64, 1000])
def read_piece(request, monkeypatch):
    """The grain of the parallel read / split (2 MB by default): tiny grains put part boundaries inside every line
    and every record of the small test inputs."""
    if request.param is None:
        monkeypatch.delenv("BFC_B200_READ_PIECE", raising=False)
    else:
        monkeypatch.setenv("BFC_B200_READ_PIECE", str(request.param))
    return request.param


@pytest.fixture(scope="module")
def L():
    lib = C.CDLL(bfc_b200.lib_path())
    lib.bseq_open.restype = C.c_void_p
    lib.bseq_open.argtypes = [C.c_char_p]
    lib.bseq_read.restype = C.POINTER(Bseq1)
    lib.bseq_read.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.bseq_close.argtypes = [C.c_void_p]
    lib.bseq_at_eof.argtypes = [C.c_void_p]
    lib.fq_open.restype = C.c_void_p
    lib.fq_open.argtypes = [C.c_char_p, C.c_int]
    lib.fq_next.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(Block)]
    lib.fq_block_free.argtypes = [C.POINTER(Block)]
    lib.fq_close.argtypes = [C.c_void_p]
    lib.fq_reader_is_fast.argtypes = [C.c_void_p]
    return lib


def read_slow(L, path, keep_comment):
    f = L.bseq_open(path.encode())
    out = []
    while True:
        n = C.c_int(0)
        seqs = L.bseq_read(f, 1000, keep_comment, C.byref(n))
        if not seqs or n.value == 0:
            if L.bseq_at_eof(f):
                break
            continue  # the batch ended at a record with a wrong quality length (kseq's -2): reading goes on behind it
        for i in range(n.value):
            s = seqs[i]
            out.append((s.name, s.comment, s.seq, s.qual))
    L.bseq_close(f)
    return out


def read_fast(L, path, keep_comment, target, threads=3):
    f = L.fq_open(path.encode(), threads)
    out, fast_blocks = [], 0
    while True:
        b = Block()
        rc = L.fq_next(f, target, keep_comment, C.byref(b))
        assert rc >= 0
        if rc == 0:
            break
        fast_blocks += L.fq_reader_is_fast(f)
        raw = C.string_at(b.buf, b.buf_len)
        for i in range(b.n):
            name = raw[b.name_off[i]:b.name_off[i] + b.name_len[i]]
            com = None if b.com_off[i] == NONE else raw[b.com_off[i]:b.com_off[i] + b.com_len[i]]
            seq = raw[b.seq_off[i]:b.seq_off[i] + b.seq_len[i]]
            qual = None if b.qual_off[i] == NONE else raw[b.qual_off[i]:b.qual_off[i] + b.seq_len[i]]
            out.append((name, com, seq, qual))
        L.fq_block_free(C.byref(b))
    L.fq_close(f)
    return out, fast_blocks


def rand_fastq(n, seed, comments=0.0, first_comment=True):
    rng = random.Random(seed)
    recs = []
    for i in range(n):
        l = rng.randint(1, 180)
        seq = "".join(rng.choice("ACGTN") for _ in range(l))
        qual = "".join(chr(rng.randint(33, 74)) for _ in range(l))
        if qual[0] in "@+>" and rng.random() < 0.5:
            qual = "@" + qual[1:]  # quality lines may start with '@'
        com = ""
        if rng.random() < comments and (i > 0 or first_comment):
            com = rng.choice([" ", "\t"]) + rng.choice(["ec:Z:0_0:1_0_0:0_0", "1:N:0", "", " x  y "])
        recs.append(f"@r{i}{com}\n{seq}\n+\n{qual}\n")
    return "".join(recs).encode()


CASES = {
    "plain": rand_fastq(5000, 1),
    "comments": rand_fastq(3000, 2, comments=0.3, first_comment=False),
    "no_final_newline": rand_fastq(500, 3)[:-1],
    "crlf": rand_fastq(300, 4).replace(b"\n", b"\r\n"),
    "blank_line_midway": rand_fastq(2000, 5) + b"\n" + rand_fastq(2000, 6, comments=0.2),
    "multiline_then_plain": b"@m1 c1\nACGT\nAC\n+\nIIII\nII\n" + rand_fastq(1500, 7, comments=0.1),
    "fasta_mixed": rand_fastq(800, 8, comments=0.2) + b">f1 hello\nACGTNN\nACG\n>f2\nAC\n" + rand_fastq(10, 9),
    "truncated_tail": rand_fastq(1000, 10) + b"@last\nACGT\n+\nII",
    "bad_qual_length_stops": rand_fastq(700, 11) + b"@bad\nACGT\n+\nIIIII\n" + rand_fastq(50, 12),
    "leading_garbage": b"garbage\n\n" + rand_fastq(100, 13),
    "empty": b"",
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("keep_comment", [0, 1])
@pytest.mark.parametrize("target", [4096, 50_000, 10_000_000])
def test_block_reader_equals_record_reader(L, tmp_path, name, keep_comment, target):
    p = str(tmp_path / "in.fq")
    with open(p, "wb") as f:
        f.write(CASES[name])
    want = read_slow(L, p, keep_comment)
    got, fast_blocks = read_fast(L, p, keep_comment, target)
    assert got == want
    if name in ("plain", "comments", "no_final_newline"):
        assert fast_blocks > 0  # the parallel path did the work


def test_block_reader_gzip_and_threads(L, tmp_path):
    p = str(tmp_path / "in.fq.gz")
    with gzip.open(p, "wb") as f:
        f.write(CASES["comments"])
    want = read_slow(L, p, 1)
    for threads in (1, 2, 7):
        got, fast_blocks = read_fast(L, p, 1, 30_000, threads)
        assert got == want and fast_blocks > 0


class Batch(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_bytes", C.c_uint64), ("where", C.c_int), ("off", u64p),
                ("seq", C.POINTER(C.c_uint8)), ("qual", C.POINTER(C.c_uint8))]


class Flat(C.Structure):
    _fields_ = [("b", Batch), ("off", u64p), ("flat_idx", C.POINTER(C.c_int64)), ("seq_buf", C.c_void_p), ("qual_buf", C.c_void_p),
                ("cap_bytes", C.c_size_t), ("cap_reads", C.c_size_t), ("pinned", C.c_int)]


class Out(C.Structure):
    _fields_ = [("filter_mode", C.c_int), ("discard", C.c_int), ("no_qual", C.c_int), ("refine", C.c_int),
                ("aux", u32p), ("keep", C.POINTER(C.c_uint8)), ("tstart", C.POINTER(C.c_int32)), ("tend", C.POINTER(C.c_int32))]


@pytest.mark.parametrize("refine", [0, 1])
def test_writer_with_records_left_out_of_the_batch(L, tmp_path, refine):
    """fq_flat_fill(skip) + fq_write: the -R writer (records left alone keep comment and bytes, the others get a fresh
    tag from aux) and the plain writer, against the printers of tests/orc.py (which mirror correct.c:591-611)."""
    import gzip as gz
    import numpy as np
    import orc
    from golden_util import Case
    c = Case("k33_rep")
    data = c.refine_forced_in if refine else gz.open(os.path.join(os.path.dirname(__file__), "golden", "k33_rep", "in.fq.gz")).read()
    p = str(tmp_path / "in.fq")
    open(p, "wb").write(data)
    recs = orc.parse_fastx(data)
    comments, skip, ori = orc.refine_plan(recs)
    if refine:
        skip = [i % 3 == 0 for i in range(len(recs))]   # leave every third record out, whatever its tag says
    else:
        skip = [False] * len(recs)
    L.fq_flat_fill.argtypes = [C.POINTER(Flat), C.POINTER(Block), C.POINTER(C.c_uint8), C.c_int]
    L.fq_write.argtypes = [C.c_void_p, C.POINTER(Block), C.POINTER(Flat), C.POINTER(Out), C.c_int]
    L.fq_flat_free.argtypes = [C.POINTER(Flat)]
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    f = L.fq_open(p.encode(), 3)
    b, flat = Block(), Flat()
    assert L.fq_next(f, 1 << 30, 1 if refine else 0, C.byref(b)) == 1 and b.n == len(recs)
    sk = (C.c_uint8 * len(recs))(*[1 if s else 0 for s in skip])
    assert L.fq_flat_fill(C.byref(flat), C.byref(b), sk if refine else None, 3) == 0
    m = flat.b.n_reads
    assert m == sum(1 for s in skip if not s)
    rng = np.random.default_rng(5)
    aux = rng.integers(0, 1 << 32, size=2 * m, dtype=np.uint32)
    aux[0::2] &= ~np.uint32(6)  # ec_code 0 or 1
    o = Out(0, 0, 0, refine, aux.ctypes.data_as(u32p), None, None, None)
    outp = str(tmp_path / "out.fq")
    fp = libc.fopen(outp.encode(), b"wb")
    assert L.fq_write(fp, C.byref(b), C.byref(flat), C.byref(o), 3) == 0
    libc.fclose(fp)
    n_bytes = flat.b.n_bytes
    seq = np.ctypeslib.as_array(flat.b.seq, shape=(n_bytes,)).copy()
    qual = np.ctypeslib.as_array(flat.b.qual, shape=(n_bytes,)).copy()
    off = np.ctypeslib.as_array(flat.b.off, shape=(m + 1,)).copy()
    want = orc.format_refined(recs, comments, skip, seq, qual, off, aux)
    assert open(outp, "rb").read() == want
    L.fq_flat_free(C.byref(flat))
    L.fq_block_free(C.byref(b))
    L.fq_close(f)


def mutate(rng, base: bytearray) -> bytes:
    for _ in range(rng.randint(0, 6)):
        if not base:
            break
        k = rng.randint(0, len(base) - 1)
        op = rng.randint(0, 5)
        if op == 0:
            base[k:k + 1] = b""
        elif op == 1:
            base[k:k] = rng.choice([b"\n", b"@", b"+", b">", b"\r", b" ", b"\n\n", b"\t"])
        elif op == 2:
            base[k] = rng.choice(b"@+>\nACGT !\r")
        elif op == 3:
            del base[k:k + rng.randint(1, 300)]
        elif op == 4:
            base[k:k] = base[max(0, k - rng.randint(1, 200)):k]
        else:
            base = base[:k]
    return bytes(base)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fuzz_block_reader_equals_record_reader(L, tmp_path, seed):
    """Randomly damaged FASTQ (deleted / inserted / duplicated stretches, stray '@' '+' '>' CR and blank lines,
    truncation): whatever the tolerant reader makes of it, the block reader delivers the same records."""
    rng = random.Random(seed)
    p = str(tmp_path / "f.fq")
    for it in range(60):
        data = mutate(rng, bytearray(rand_fastq(rng.randint(1, 300), rng.randint(0, 1 << 30), comments=rng.choice([0, 0.3]))))
        with open(p, "wb") as f:
            f.write(data)
        for keep in (0, 1):
            want = read_slow(L, p, keep)
            for target in (4096, 20_000, 1 << 22):
                got, _ = read_fast(L, p, keep, target, threads=rng.choice([1, 3]))
                assert got == want, (seed, it, keep, target)


REF_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libbfcref.so")


@pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref (the unmodified reference, built where /root/reference exists) is absent")
@pytest.mark.parametrize("seed", [11, 12])
def test_fuzz_record_reader_equals_the_reference_reader(L, tmp_path, seed):
    """csrc/bseq.c against the UNMODIFIED reference's bseq_read (kseq.h underneath) on the same damaged inputs, one
    batch each: same records, same stops, same sticky comments."""
    R = C.CDLL(REF_LIB)
    R.bseq_open.restype = C.c_void_p
    R.bseq_open.argtypes = [C.c_char_p]
    R.bseq_read.restype = C.POINTER(Bseq1)
    R.bseq_read.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    R.bseq_close.argtypes = [C.c_void_p]

    def one_batch(lib, path, keep):
        f = lib.bseq_open(path.encode())
        n = C.c_int(0)
        seqs = lib.bseq_read(f, 1 << 30, keep, C.byref(n))
        out = [(seqs[i].name, seqs[i].comment, seqs[i].seq, seqs[i].qual) for i in range(n.value)]
        lib.bseq_close(f)
        return out

    rng = random.Random(seed)
    p = str(tmp_path / "f.fq")
    for it in range(150):
        data = mutate(rng, bytearray(rand_fastq(rng.randint(1, 300), rng.randint(0, 1 << 30), comments=rng.choice([0, 0.3]))))
        with open(p, "wb") as f:
            f.write(data)
        for keep in (0, 1):
            assert one_batch(L, p, keep) == one_batch(R, p, keep), (seed, it, keep)


def block_records(b):
    raw = C.string_at(b.buf, b.buf_len)
    out = []
    for i in range(b.n):
        com = None if b.com_off[i] == NONE else raw[b.com_off[i]:b.com_off[i] + b.com_len[i]]
        qual = None if b.qual_off[i] == NONE else raw[b.qual_off[i]:b.qual_off[i] + b.seq_len[i]]
        out.append((raw[b.name_off[i]:b.name_off[i] + b.name_len[i]], com, raw[b.seq_off[i]:b.seq_off[i] + b.seq_len[i]], qual))
    return out


def test_blocks_kept_for_the_second_pass(L, tmp_path, monkeypatch):
    """fq_keep_*: what the count pass parsed is handed to the correct pass of the same file; a changed file, another
    keep_comment or a budget that is too small means "read it again" (fq_keep_match = -1)."""
    L.fq_keep_begin.argtypes = [C.c_char_p, C.c_int]
    L.fq_keep_add.argtypes = [C.POINTER(Block)]
    L.fq_keep_end.argtypes = [C.c_int]
    L.fq_keep_match.restype = C.c_long
    L.fq_keep_match.argtypes = [C.c_char_p, C.c_int]
    L.fq_keep_take.argtypes = [C.c_long, C.POINTER(Block)]
    path = str(tmp_path / "in.fq")
    with open(path, "wb") as fp:
        fp.write(rand_fastq(3000, 5, comments=0.5))
    want, _ = read_fast(L, path, 1, 20000)

    def first_pass(budget=1 << 30):
        monkeypatch.setenv("BFC_B200_KEEP_MAX", str(budget))
        L.fq_keep_begin(path.encode(), 1)
        f = L.fq_open(path.encode(), 3)
        n_blocks = 0
        while True:
            b = Block()
            if L.fq_next(f, 20000, 1, C.byref(b)) != 1:
                break
            n_blocks += 1
            if not L.fq_keep_add(C.byref(b)):
                L.fq_block_free(C.byref(b))
            else:
                assert not b.buf and b.n == 0       # ownership moved
        L.fq_close(f)
        L.fq_keep_end(1)
        return n_blocks

    n_blocks = first_pass()
    assert n_blocks > 3
    assert L.fq_keep_match(path.encode(), 0) == -1                     # the other pass wants other blocks
    assert L.fq_keep_match((path + "x").encode(), 1) == -1
    assert L.fq_keep_match(path.encode(), 1) == n_blocks
    got = []
    for i in range(n_blocks):
        b = Block()
        assert L.fq_keep_take(i, C.byref(b)) == 1
        got += block_records(b)
        L.fq_block_free(C.byref(b))
    assert got == want
    L.fq_keep_drop()
    assert L.fq_keep_match(path.encode(), 1) == -1

    first_pass(budget=50000)                                           # two blocks fit, the third does not: nothing is kept
    assert L.fq_keep_match(path.encode(), 1) == -1
    first_pass(budget=0)
    assert L.fq_keep_match(path.encode(), 1) == -1

    monkeypatch.delenv("BFC_B200_KEEP_MAX")                              # default: plain files are read again, gzip'd ones are kept
    L.fq_keep_begin(path.encode(), 1)
    b = Block()
    f = L.fq_open(path.encode(), 3)
    assert L.fq_next(f, 20000, 1, C.byref(b)) == 1 and L.fq_keep_add(C.byref(b)) == 0
    L.fq_block_free(C.byref(b))
    L.fq_close(f)
    gz = path + ".gz"
    with open(path, "rb") as src, gzip.open(gz, "wb") as dst:
        dst.write(src.read())
    L.fq_keep_begin(gz.encode(), 1)
    f = L.fq_open(gz.encode(), 3)
    assert L.fq_next(f, 20000, 1, C.byref(b)) == 1 and L.fq_keep_add(C.byref(b)) == 1
    L.fq_close(f)
    L.fq_keep_drop()

    first_pass()
    os.utime(path, ns=(1, 1))                                          # the file changed between the passes
    assert L.fq_keep_match(path.encode(), 1) == -1
    L.fq_keep_drop()
