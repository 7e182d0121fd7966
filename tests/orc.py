"""Test-only helpers: ctypes access to the CPU oracle (oracle/liboracle.so) and, when
it has been built, to the unmodified reference (oracle/_ref/).  Nothing in the product
package imports this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)


class Opt(C.Structure):
    """bfc_opt_t (reference bfc.h:15-33)."""
    _fields_ = [(n, C.c_int) for n in ("chunk_size", "n_threads", "no_mt_io", "q", "k",
                                        "filter_mode", "refine_ec", "no_qual")] + \
               [("min_frac", C.c_float)] + \
               [(n, C.c_int) for n in ("l_pre", "bf_shift", "n_hashes", "discard", "max_end_ext",
                                        "win_multi_ec", "min_cov", "w_ec", "w_ec_high", "w_absent",
                                        "w_absent_high", "max_path_diff", "max_heap")]


def make_opt(**kw) -> Opt:
    """Defaults of bfc_opt_init (reference bfc.c:17-40)."""
    o = Opt(chunk_size=100000000, n_threads=1, no_mt_io=0, q=20, k=33, filter_mode=0, refine_ec=0,
            no_qual=0, min_frac=0.9, l_pre=20, bf_shift=33, n_hashes=4, discard=0, max_end_ext=5,
            win_multi_ec=10, min_cov=3, w_ec=1, w_ec_high=7, w_absent=3, w_absent_high=1,
            max_path_diff=15, max_heap=100)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class Batch(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("off", u64p), ("seq", u8p), ("qual", u8p)]


class BF(C.Structure):
    _fields_ = [("n_shift", C.c_int), ("n_hashes", C.c_int), ("b", u8p)]


def build_oracle(ref: bool = False) -> None:
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
    if ref and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(path):
        build_oracle()
    L = C.CDLL(path)
    L.orc_bf_new.restype = C.POINTER(BF)
    L.orc_bf_new.argtypes = [C.c_int, C.c_int]
    L.orc_bf_free.argtypes = [C.POINTER(BF)]
    L.orc_bf_insert.argtypes = [C.POINTER(BF), C.c_uint64]
    L.orc_bf_get.argtypes = [C.POINTER(BF), C.c_uint64]
    L.orc_ch_new.restype = C.c_void_p
    L.orc_ch_new.argtypes = [C.c_int, C.c_int]
    L.orc_ch_free.argtypes = [C.c_void_p]
    L.orc_ch_k.argtypes = [C.c_void_p]
    L.orc_ch_lpre.argtypes = [C.c_void_p]
    L.orc_ch_subkey.argtypes = [C.c_void_p, u64p, u32p, u64p]
    L.orc_ch_insert.argtypes = [C.c_void_p, u64p, C.c_int]
    L.orc_ch_get.argtypes = [C.c_void_p, u64p]
    L.orc_ch_kmer_occ.argtypes = [C.c_void_p, u64p]
    L.orc_ch_count.restype = C.c_uint64
    L.orc_ch_count.argtypes = [C.c_void_p]
    L.orc_ch_hist.argtypes = [C.c_void_p, u64p, u64p]
    L.orc_ch_export.restype = C.c_uint64
    L.orc_ch_export.argtypes = [C.c_void_p, u32p, u64p]
    L.orc_ch_put_raw.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64]
    L.orc_kmer_append.argtypes = [C.c_int, u64p, C.c_int]
    L.orc_kmer_change.argtypes = [C.c_int, u64p, C.c_int, C.c_int]
    L.orc_hash_64.restype = C.c_uint64
    L.orc_hash_64.argtypes = [C.c_uint64, C.c_uint64]
    L.orc_kmer_hash.restype = C.c_uint64
    L.orc_kmer_hash.argtypes = [C.c_int, u64p, u64p]
    L.orc_count_batch.argtypes = [C.POINTER(Opt), C.POINTER(BF), C.POINTER(BF), C.c_void_p,
                                  C.POINTER(Batch), u64p]
    L.orc_enum_records.restype = C.c_uint64
    L.orc_enum_records.argtypes = [C.POINTER(Opt), C.POINTER(Batch), u64p, u64p]
    L.orc_hash_from_y.restype = C.c_uint64
    L.orc_hash_from_y.argtypes = [C.c_int, C.c_uint64, C.c_uint64]
    L.orc_count_records.argtypes = [C.POINTER(Opt), C.POINTER(BF), C.POINTER(BF), C.c_void_p, C.c_uint64, u64p, u64p, u64p]
    L.orc_correct_batch.argtypes = [C.POINTER(Opt), C.c_void_p, C.c_int, C.c_int64, u64p, u8p, u8p,
                                    u32p, u64p]
    L.orc_max_streak.restype = C.c_uint64
    L.orc_max_streak.argtypes = [C.c_int, C.POINTER(BF), C.c_char_p, C.c_int]
    L.orc_trim_batch.argtypes = [C.POINTER(Opt), C.POINTER(BF), C.c_int64, u64p, u8p, u8p, i32p, i32p]
    _lib = L
    return L


def have_ref() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "bfc")) and os.path.exists(os.path.join(REF_DIR, "libbfcref.so"))


_ref = None


def reflib():
    global _ref
    if _ref is not None:
        return _ref
    R = C.CDLL(os.path.join(REF_DIR, "libbfcref.so"))
    R.ref_kmer_append.argtypes = [C.c_int, u64p, C.c_int]
    R.ref_kmer_change.argtypes = [C.c_int, u64p, C.c_int, C.c_int]
    R.ref_hash_64.restype = C.c_uint64
    R.ref_hash_64.argtypes = [C.c_uint64, C.c_uint64]
    R.ref_hash_64_inv.restype = C.c_uint64
    R.ref_hash_64_inv.argtypes = [C.c_uint64, C.c_uint64]
    R.ref_kmer_hash.restype = C.c_uint64
    R.ref_kmer_hash.argtypes = [C.c_int, u64p, u64p]
    R.bfc_bf_init.restype = C.POINTER(BF)
    R.bfc_bf_init.argtypes = [C.c_int, C.c_int]
    R.bfc_bf_destroy.argtypes = [C.POINTER(BF)]
    R.bfc_bf_insert.argtypes = [C.POINTER(BF), C.c_uint64]
    R.bfc_bf_get.argtypes = [C.POINTER(BF), C.c_uint64]
    R.bfc_ch_init.restype = C.c_void_p
    R.bfc_ch_init.argtypes = [C.c_int, C.c_int]
    R.bfc_ch_destroy.argtypes = [C.c_void_p]
    R.bfc_ch_insert.argtypes = [C.c_void_p, u64p, C.c_int, C.c_int]
    R.bfc_ch_get.argtypes = [C.c_void_p, u64p]
    R.bfc_ch_count.restype = C.c_uint64
    R.bfc_ch_count.argtypes = [C.c_void_p]
    R.bfc_ch_hist.argtypes = [C.c_void_p, u64p, u64p]
    _ref = R
    return R


# ----------------------------------------------------------------------------- data plumbing

def as_u8p(a: np.ndarray):
    return a.ctypes.data_as(u8p)


def as_u64p(a: np.ndarray):
    return a.ctypes.data_as(u64p)


def parse_fastx(data: bytes):
    """Minimal FASTA/FASTQ reader for test inputs (4-line FASTQ, or multi-line FASTA).
    Returns list of (name, comment, seq, qual-or-None), all bytes."""
    recs = []
    lines = data.split(b"\n")
    i = 0
    while i < len(lines):
        ln = lines[i]
        if not ln:
            i += 1
            continue
        if ln[:1] == b"@":
            hdr = ln[1:]
            seq = lines[i + 1]
            qual = lines[i + 3] or None  # kseq: an empty quality string means "no quality" (bseq.c:68)
            i += 4
        elif ln[:1] == b">":
            hdr = ln[1:]
            i += 1
            parts = []
            while i < len(lines) and lines[i][:1] not in (b">", b"@"):
                parts.append(lines[i])
                i += 1
            seq, qual = b"".join(parts), None
        else:
            raise ValueError("bad record at line %d" % i)
        sp = hdr.split(None, 1)
        name = sp[0] if sp else b""
        comment = sp[1] if len(sp) > 1 else None
        recs.append((name, comment, seq, qual))
    return recs


def batch_from_records(recs):
    """-> (seq u8[], qual u8[] or None, off u64[n+1]) in the C-ABI host batch layout."""
    n = len(recs)
    off = np.zeros(n + 1, dtype=np.uint64)
    tot = 0
    for i, r in enumerate(recs):
        off[i] = tot
        tot += len(r[2]) + 1
    off[n] = tot
    seq = np.zeros(tot, dtype=np.uint8)
    any_q = any(r[3] is not None for r in recs)
    qual = np.zeros(tot, dtype=np.uint8) if any_q else None
    for i, r in enumerate(recs):
        o = int(off[i])
        seq[o:o + len(r[2])] = np.frombuffer(r[2], dtype=np.uint8)
        if r[3] is not None:
            qual[o:o + len(r[3])] = np.frombuffer(r[3], dtype=np.uint8)
    return seq, qual, off


def format_corrected(recs, seq: np.ndarray, qual, off: np.ndarray, aux: np.ndarray, no_qual=False) -> bytes:
    """The step-2 printer of the reference (correct.c:591-611), normal mode, no -D."""
    out = []
    for i, r in enumerate(recs):
        o, e = int(off[i]), int(off[i + 1]) - 1
        a, a2 = int(aux[2 * i]), int(aux[2 * i + 1])
        has_q = r[3] is not None
        is_fq = has_q and not no_qual
        tag = b"\tec:Z:%d" % (a & 7)
        if (a & 7) == 0:
            tag += b"_%d:%d_%d_%d:%d_%d" % (a2 >> 10, a2 & 0xff, a >> 3 & 1, a >> 18 & 0x3fff,
                                           a >> 4 & 0x3fff, a2 >> 8 & 3)
        out.append((b"@" if is_fq else b">") + r[0] + tag + b"\n" + seq[o:e].tobytes() + b"\n")
        if is_fq:
            out.append(b"+\n" + qual[o:e].tobytes() + b"\n")
    return b"".join(out)


def parse_ec_tag(comment: bytes):
    """parse_stats (reference correct.c:517-531) -> (aux, aux2) as worker_ec packs them (correct.c:552-553)."""
    import re
    v = [int(x) for x in re.findall(rb"-?\d+", comment[5:])] + [0] * 6
    ec_code = v[0]
    if ec_code != 0:
        return ec_code & 7, 0
    n_absent, max_heap, brute, n_ec, n_ec_high = v[1:6]
    return (n_ec & 0x3fff) << 18 | (n_ec_high & 0x3fff) << 4 | (brute & 1) << 3, (n_absent & 0x3fffff) << 10 | 1 << 8 | (max_heap & 0xff)


def refine_plan(recs):
    """worker_ec's refine branch (correct.c:542-550) under -t1.  Returns (comments, skip, ori): the comment every record
    carries (kseq's sticky comment), which records are left alone (earlier round fine and max_heap < 50), and for the
    others the earlier stats e->ori_st holds when they are corrected -- those of the latest tagged record so far."""
    comments, skip, ori = [], [], []
    sticky, cur = None, (0, 0)
    for r in recs:
        if r[1] is not None:
            sticky = r[1]
        comments.append(sticky)
        sk = False
        if sticky is not None and sticky.startswith(b"ec:Z:"):
            cur = parse_ec_tag(sticky)
            sk = (cur[0] & 7) == 0 and (cur[1] & 0xff) < 50
        skip.append(sk)
        ori.append(cur)
    return comments, skip, ori


def format_refined(recs, comments, skip, seq, qual, off, aux, no_qual=False) -> bytes:
    """The step-2 printer in -R mode: skipped records keep their comment and their bytes; the others (whose corrected
    bytes and stats are the i-th entries of seq / qual / off / aux, in order) get a fresh tag."""
    out, j = [], 0
    for i, r in enumerate(recs):
        is_fq = r[3] is not None and not no_qual
        if skip[i]:
            out.append((b"@" if is_fq else b">") + r[0] + b"\t" + comments[i] + b"\n" + r[2] + b"\n")
            if is_fq:
                out.append(b"+\n" + r[3] + b"\n")
            continue
        o, e = int(off[j]), int(off[j + 1]) - 1
        a, a2 = int(aux[2 * j]), int(aux[2 * j + 1])
        tag = b"\tec:Z:%d" % (a & 7)
        if (a & 7) == 0:
            tag += b"_%d:%d_%d_%d:%d_%d" % (a2 >> 10, a2 & 0xff, a >> 3 & 1, a >> 18 & 0x3fff, a >> 4 & 0x3fff, a2 >> 8 & 3)
        out.append((b"@" if is_fq else b">") + r[0] + tag + b"\n" + seq[o:e].tobytes() + b"\n")
        if is_fq:
            out.append(b"+\n" + qual[o:e].tobytes() + b"\n")
        j += 1
    return b"".join(out)


def format_trimmed(recs, seq, qual, off, keep, tstart, tend, no_qual=False) -> bytes:
    """The step-2 printer of the reference in `-1` mode (correct.c:605-611)."""
    out = []
    sticky = None  # kseq never clears comment.s, and bseq_read strdup()s whatever is there
    for i, r in enumerate(recs):  # (bseq.c:66, kseq.h:190): a comment persists until replaced
        if r[1]:
            sticky = r[1]
        if not keep[i]:
            continue
        o = int(off[i])
        is_fq = r[3] is not None and not no_qual
        hdr = (b"@" if is_fq else b">") + r[0] + (b"\t" + sticky if sticky else b"")
        s, e = o + int(tstart[i]), o + int(tend[i])
        out.append(hdr + b"\n" + seq[s:e].tobytes() + b"\n")
        if is_fq:
            out.append(b"+\n" + qual[s:e].tobytes() + b"\n")
    return b"".join(out)


# ----------------------------------------------------------------------------- reference runs

def ref_run(args, binary="bfc", env=None):
    """Run the unmodified reference binary; returns stdout bytes."""
    e = dict(os.environ)
    if env:
        e.update(env)
    p = subprocess.run([os.path.join(REF_DIR, binary)] + list(args), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, env=e, check=True)
    return p.stdout


def parse_ref_dump(path: str):
    """Reference `-d` file (htab.c:129-149) -> (k, l_pre, sub u32[], key u64[]) sorted by (sub, key)."""
    raw = np.fromfile(path, dtype=np.uint32)
    k, l_pre = int(raw[0]), int(raw[1])
    pos = 2
    subs, keys = [], []
    for s in range(1 << l_pre):
        size = int(raw[pos + 1])
        pos += 2
        if size:
            kk = raw[pos:pos + 2 * size].view(np.uint64)
            keys.append(kk)
            subs.append(np.full(size, s, dtype=np.uint32))
            pos += 2 * size
    if keys:
        sub = np.concatenate(subs)
        key = np.concatenate(keys)
    else:
        sub = np.zeros(0, np.uint32)
        key = np.zeros(0, np.uint64)
    order = np.lexsort((key, sub))
    return k, l_pre, sub[order], key[order]


def read_bloom_dump(path: str):
    """File written by oracle/ref_hooks.c -> (n_shift, n_hashes, bytes u8[])."""
    with open(path, "rb") as fp:
        hdr = np.frombuffer(fp.read(8), dtype=np.int32)
        b = np.frombuffer(fp.read(), dtype=np.uint8)
    return int(hdr[0]), int(hdr[1]), b


# ----------------------------------------------------------------------------- oracle runs

class OracleRun:
    """Count (and optionally correct / trim) a list of records with the CPU oracle."""

    def __init__(self, opt: Opt):
        self.L = lib()
        self.opt = opt
        self.bf = self.L.orc_bf_new(opt.bf_shift, opt.n_hashes)
        self.bf_high = self.L.orc_bf_new(opt.bf_shift, opt.n_hashes) if opt.filter_mode else None
        self.ch = None if opt.filter_mode else self.L.orc_ch_new(opt.k, opt.l_pre)
        self.stats = np.zeros(2, dtype=np.uint64)

    def close(self):
        self.L.orc_bf_free(self.bf)
        if self.bf_high:
            self.L.orc_bf_free(self.bf_high)
        if self.ch:
            self.L.orc_ch_free(self.ch)

    def count(self, seq, qual, off):
        b = Batch(len(off) - 1, as_u64p(off), as_u8p(seq), as_u8p(qual) if qual is not None else None)
        self.L.orc_count_batch(C.byref(self.opt), self.bf, self.bf_high, self.ch, C.byref(b), as_u64p(self.stats))

    def bloom_bytes(self, high=False) -> np.ndarray:
        bf = self.bf_high if high else self.bf
        n = 1 << (self.opt.bf_shift - 3)
        return np.ctypeslib.as_array(bf.contents.b, shape=(n,)).copy()

    def table(self):
        n = int(self.L.orc_ch_export(self.ch, None, None))
        sub = np.zeros(n, dtype=np.uint32)
        key = np.zeros(n, dtype=np.uint64)
        if n:
            self.L.orc_ch_export(self.ch, sub.ctypes.data_as(u32p), as_u64p(key))
        return sub, key

    def hist(self):
        cnt = np.zeros(256, dtype=np.uint64)
        high = np.zeros(64, dtype=np.uint64)
        mode = self.L.orc_ch_hist(self.ch, as_u64p(cnt), as_u64p(high))
        return mode, cnt, high

    def correct(self, seq, qual, off, ori=None):
        """Returns (seq', qual', aux[2n], counters[3]); inputs are not modified.  Refine mode (opt.refine_ec): `ori` =
        the earlier stats, 2 words per read (aux is in/out in the C call)."""
        s = seq.copy()
        q = qual.copy() if qual is not None else None
        n = len(off) - 1
        aux = np.zeros(2 * n, dtype=np.uint32) if ori is None else np.array(ori, dtype=np.uint32).reshape(-1).copy()
        counters = np.zeros(3, dtype=np.uint64)
        mode, _, _ = self.hist()
        self.L.orc_correct_batch(C.byref(self.opt), self.ch, mode, n, as_u64p(off), as_u8p(s),
                                 as_u8p(q) if q is not None else None, aux.ctypes.data_as(u32p), as_u64p(counters))
        return s, q, aux, counters

    def trim(self, seq, off):
        n = len(off) - 1
        keep = np.zeros(n, dtype=np.uint8)
        ts = np.zeros(n, dtype=np.int32)
        te = np.zeros(n, dtype=np.int32)
        self.L.orc_trim_batch(C.byref(self.opt), self.bf_high, n, as_u64p(off), as_u8p(seq), as_u8p(keep),
                              ts.ctypes.data_as(i32p), te.ctypes.data_as(i32p))
        return keep, ts, te
