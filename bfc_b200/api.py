"""ctypes binding of include/bfc_b200.h + the reference-named entry points (bbf.h, htab.h, bfc.h).

`Engine` mirrors what the reference's main() does around its two phases
(bfc.c:131-150): create the Bloom filter(s) / table, count batches, take the
histogram mode, correct or trim batches.  All arrays are numpy; `where` says whether a
Batch's pointers are host or device memory.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)

HOST, DEVICE = 0, 1


class BfcError(RuntimeError):
    pass


class Opt(C.Structure):
    """bfc_opt_t (include/bfc.h; reference bfc.h:15-33)."""
    _fields_ = [(n, C.c_int) for n in ("chunk_size", "n_threads", "no_mt_io", "q", "k",
                                        "filter_mode", "refine_ec", "no_qual")] + \
               [("min_frac", C.c_float)] + \
               [(n, C.c_int) for n in ("l_pre", "bf_shift", "n_hashes", "discard", "max_end_ext",
                                        "win_multi_ec", "min_cov", "w_ec", "w_ec_high", "w_absent",
                                        "w_absent_high", "max_path_diff", "max_heap")]


class Batch(C.Structure):
    """bfcg_batch_t."""
    _fields_ = [("n_reads", C.c_int64), ("n_bytes", C.c_uint64), ("where", C.c_int),
                ("off", u64p), ("seq", u8p), ("qual", u8p)]


class Stats(C.Structure):
    """bfcg_stats_t."""
    _fields_ = [(n, C.c_uint64) for n in ("n_kmers", "n_pass", "n_pending", "n_conflict", "n_lookups",
                                           "n_redo", "n_launches")] + [("kernel_ms", C.c_double), ("n_search_lookups", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class BF(C.Structure):
    """bfc_bf_t (include/bbf.h); `b` is a device pointer."""
    _fields_ = [("n_shift", C.c_int), ("n_hashes", C.c_int), ("b", C.c_void_p)]


def lib_path() -> str:
    return os.path.join(LIBDIR, "libbfc_b200.so")


def build(verbose: bool = False) -> None:
    """Compile csrc/ into lib/ with nvcc for sm_100a (cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC, "-j", "8"], stdout=out)


_lib = None


def lib():
    """The loaded shared library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise BfcError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    L = C.CDLL(p, mode=C.RTLD_GLOBAL)
    L.bfcg_device_count.restype = C.c_int
    L.bfcg_set_device.argtypes = [C.c_int]
    L.bfcg_last_error.restype = C.c_char_p
    L.bfcg_set_timing.argtypes = [C.c_int]
    L.bfc_opt_init.argtypes = [C.POINTER(Opt)]
    L.bfc_opt_by_size.argtypes = [C.POINTER(Opt), C.c_long]
    L.bfc_bf_init.restype = C.POINTER(BF)
    L.bfc_bf_init.argtypes = [C.c_int, C.c_int]
    L.bfc_bf_destroy.argtypes = [C.POINTER(BF)]
    L.bfc_bf_insert.argtypes = [C.POINTER(BF), C.c_uint64]
    L.bfc_bf_get.argtypes = [C.POINTER(BF), C.c_uint64]
    L.bfc_ch_init.restype = C.c_void_p
    L.bfc_ch_init.argtypes = [C.c_int, C.c_int]
    L.bfc_ch_destroy.argtypes = [C.c_void_p]
    L.bfc_ch_insert.argtypes = [C.c_void_p, u64p, C.c_int, C.c_int]
    L.bfc_ch_get.argtypes = [C.c_void_p, u64p]
    L.bfc_ch_kmer_occ.argtypes = [C.c_void_p, u64p]
    L.bfc_ch_count.restype = C.c_uint64
    L.bfc_ch_count.argtypes = [C.c_void_p]
    L.bfc_ch_hist.argtypes = [C.c_void_p, u64p, u64p]
    L.bfc_ch_dump.argtypes = [C.c_void_p, C.c_char_p]
    L.bfc_ch_restore.restype = C.c_void_p
    L.bfc_ch_restore.argtypes = [C.c_char_p]
    L.bfc_ch_get_k.argtypes = [C.c_void_p]
    L.bfc_count.restype = C.c_void_p
    L.bfc_count.argtypes = [C.c_char_p, C.POINTER(Opt)]
    L.bfc_correct.argtypes = [C.c_char_p, C.POINTER(Opt), C.c_void_p]
    L.bfcg_count_batch.argtypes = [C.POINTER(Opt), C.POINTER(BF), C.POINTER(BF), C.c_void_p,
                                   C.POINTER(Batch), C.POINTER(Stats)]
    L.bfcg_correct_batch.argtypes = [C.POINTER(Opt), C.c_void_p, C.c_int, C.POINTER(Batch), C.c_void_p,
                                     C.POINTER(Stats)]
    L.bfcg_trim_batch.argtypes = [C.POINTER(Opt), C.POINTER(BF), C.POINTER(Batch), C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.POINTER(Stats)]
    L.bfcg_enum_records.argtypes = [C.POINTER(Opt), C.POINTER(Batch), C.c_int, C.c_void_p, C.c_void_p, u64p]
    L.bfcg_count_record_runs.argtypes = [C.POINTER(Opt), C.POINTER(BF), C.POINTER(BF), C.c_void_p, C.c_int, u64p,
                                         C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Stats)]
    L.bfcg_count_records.argtypes = [C.POINTER(Opt), C.POINTER(BF), C.POINTER(BF), C.c_void_p, C.c_uint64,
                                     C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Stats)]
    L.bfcg_bf_init_shard.restype = C.POINTER(BF)
    L.bfcg_bf_init_shard.argtypes = [C.c_int, C.c_int, C.c_int]
    L.bfcg_ch_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.bfcg_dist_unique_id.argtypes = [C.c_void_p]
    L.bfcg_dist_init.argtypes = [C.c_int, C.c_int, C.c_void_p]
    L.bfcg_dist_finalize.restype = None
    L.bfcg_dist_count_piece.argtypes = [C.POINTER(Opt), C.POINTER(BF), C.POINTER(BF), C.c_void_p, C.POINTER(Batch), C.POINTER(Stats)]
    L.bfcg_dist_count_finish.argtypes = [C.POINTER(Opt), C.POINTER(BF), C.POINTER(BF), C.c_void_p, C.POINTER(Stats)]
    L.bfcg_dist_gather_table.argtypes = [C.c_void_p, C.c_void_p]
    L.bfcg_dist_gather_filter.argtypes = [C.POINTER(BF), C.POINTER(BF)]
    L.bfcg_dist_allreduce_sum_u64.argtypes = [u64p, C.c_int]
    L.bfcg_dist_allreduce_max_f64.argtypes = [C.POINTER(C.c_double), C.c_int]
    L.bfcg_dist_stats.argtypes = [u64p, u64p, C.POINTER(C.c_double)]
    L.bfcg_ch_export_device.restype = C.c_uint64
    L.bfcg_ch_export_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.bfcg_ch_import_device.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    L.bfcg_dev_alloc.restype = C.c_void_p
    L.bfcg_dev_alloc.argtypes = [C.c_uint64]
    L.bfcg_dev_free.argtypes = [C.c_void_p]
    L.bfcg_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.bfcg_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.bfcg_host_alloc_pinned.restype = C.c_void_p
    L.bfcg_host_alloc_pinned.argtypes = [C.c_uint64]
    L.bfcg_host_free_pinned.argtypes = [C.c_void_p]
    L.bfcg_bf_download.argtypes = [C.POINTER(BF), u8p]
    L.bfcg_bf_upload.argtypes = [C.POINTER(BF), u8p]
    L.bfcg_bf_clear.argtypes = [C.POINTER(BF)]
    L.bfcg_bf_load.argtypes = [C.POINTER(BF), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.bfcg_bf_suggest_shift.argtypes = [C.c_uint64, C.c_int, C.c_double]
    L.bfcg_ch_export.restype = C.c_uint64
    L.bfcg_ch_export.argtypes = [C.c_void_p, u32p, u64p]
    L.bfcg_ch_l_pre.argtypes = [C.c_void_p]
    L.bfcg_ch_capacity_log2.argtypes = [C.c_void_p]
    L.bfcg_ch_clear.argtypes = [C.c_void_p]
    L.bfcg_ch_reserve.argtypes = [C.c_void_p, C.c_uint64]
    L.bfcg_partition_records.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int]
    L.bfcg_ch_get_batch.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
    L.bfcg_kernel_times.argtypes = [C.POINTER(C.c_double), u64p, C.c_int]
    L.bfcg_event_record.argtypes = [C.c_int]
    L.bfcg_event_elapsed_ms.restype = C.c_double
    L.bfcg_event_elapsed_ms.argtypes = [C.c_int, C.c_int]
    L.bfcg_synth_genome.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    L.bfcg_synth_reads.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int64, C.c_int, C.c_double,
                                   C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    _lib = L
    return L


KERNEL_NAMES = ["count_probe", "count_resolve", "conflict_sort", "count_replay", "correct", "correct_redo", "trim",
                "tab_rehash", "tab_hist", "tab_apply", "enum", "ec_lookup", "ec_setup", "ec_merge", "bucket",
                "count_part", "count_bounds", "enum_lin", "ec_ext"]


def kernel_times():
    """{kernel: (ms, launches)} of device time accumulated since the last call (timing must be on)."""
    ms = (C.c_double * 24)()
    n = (C.c_uint64 * 24)()
    k = lib().bfcg_kernel_times(ms, n, 24)
    return {KERNEL_NAMES[i]: (ms[i], int(n[i])) for i in range(min(k, len(KERNEL_NAMES)))}


def make_opt(**kw) -> Opt:
    """bfc_opt_init() defaults (reference bfc.c:17-40) with keyword overrides."""
    o = Opt()
    lib().bfc_opt_init(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def opt_by_size(o: Opt, size: float) -> Opt:
    """The `-s` rule (reference bfc.c:42-53, 112-121)."""
    lib().bfc_opt_by_size(C.byref(o), int(size) + 1)
    return o


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise BfcError(f"{what} failed ({rc}): {lib().bfcg_last_error().decode()}")


def records_to_batch(seqs, quals=None):
    """Lists of bytes -> (seq u8[], qual u8[] or None, off u64[n+1]) in the flat layout
    (a read without quality gets 0xFF bytes)."""
    n = len(seqs)
    off = np.zeros(n + 1, dtype=np.uint64)
    lens = np.fromiter((len(s) + 1 for s in seqs), dtype=np.uint64, count=n)
    np.cumsum(lens, out=off[1:])
    tot = int(off[n])
    seq = np.zeros(tot, dtype=np.uint8)
    any_q = quals is not None and any(q is not None for q in quals)
    qual = np.zeros(tot, dtype=np.uint8) if any_q else None
    for i, s in enumerate(seqs):
        o = int(off[i])
        seq[o:o + len(s)] = np.frombuffer(s, dtype=np.uint8)
        if any_q:
            q = quals[i]
            if q is None:
                qual[o:o + len(s)] = 0xFF
            else:
                qual[o:o + len(q)] = np.frombuffer(q, dtype=np.uint8)
    return seq, qual, off


def host_batch(seq: np.ndarray, qual, off: np.ndarray) -> Batch:
    b = Batch()
    b.n_reads = len(off) - 1
    b.n_bytes = int(off[-1])
    b.where = HOST
    b.off = off.ctypes.data_as(u64p)
    b.seq = seq.ctypes.data_as(u8p)
    b.qual = qual.ctypes.data_as(u8p) if qual is not None else None
    return b


class Engine:
    """One count(+correct / trim) job on one GPU, mirroring the reference's main()."""

    def __init__(self, opt: Opt, timing: bool = False):
        self.L = lib()
        self.opt = opt
        self.stats = Stats()
        self.L.bfcg_set_timing(1 if timing else 0)
        self.bf = self.L.bfc_bf_init(opt.bf_shift, opt.n_hashes)
        if not self.bf:
            raise BfcError("bfc_bf_init failed: " + self.L.bfcg_last_error().decode())
        self.bf_high = None
        self.ch = None
        if opt.filter_mode:
            self.bf_high = self.L.bfc_bf_init(opt.bf_shift, opt.n_hashes)
            if not self.bf_high:
                raise BfcError("bfc_bf_init failed: " + self.L.bfcg_last_error().decode())
        else:
            self.ch = self.L.bfc_ch_init(opt.k, opt.l_pre)
            if not self.ch:
                raise BfcError("bfc_ch_init failed: " + self.L.bfcg_last_error().decode())
        self._mode = None

    def close(self):
        if self.bf:
            self.L.bfc_bf_destroy(self.bf)
            self.bf = None
        if self.bf_high:
            self.L.bfc_bf_destroy(self.bf_high)
            self.bf_high = None
        if self.ch:
            self.L.bfc_ch_destroy(self.ch)
            self.ch = None

    def reset(self):
        """Empty filter(s) and table (between bench steps)."""
        _check(self.L.bfcg_bf_clear(self.bf), "bfcg_bf_clear")
        if self.bf_high:
            _check(self.L.bfcg_bf_clear(self.bf_high), "bfcg_bf_clear")
        if self.ch:
            _check(self.L.bfcg_ch_clear(self.ch), "bfcg_ch_clear")
        self._mode = None

    # ---- count
    def count(self, seq: np.ndarray, qual, off: np.ndarray):
        b = host_batch(seq, qual, off)
        self.count_batch(b)

    def count_batch(self, b: Batch):
        _check(self.L.bfcg_count_batch(C.byref(self.opt), self.bf, self.bf_high, self.ch, C.byref(b),
                                       C.byref(self.stats)), "bfcg_count_batch")
        self._mode = None

    # ---- inspection
    def bloom_bytes(self, high: bool = False) -> np.ndarray:
        bf = self.bf_high if high else self.bf
        out = np.zeros(1 << (self.opt.bf_shift - 3), dtype=np.uint8)
        _check(self.L.bfcg_bf_download(bf, out.ctypes.data_as(u8p)), "bfcg_bf_download")
        return out

    def table(self):
        n = int(self.L.bfcg_ch_export(self.ch, None, None))
        sub = np.zeros(n, dtype=np.uint32)
        key = np.zeros(n, dtype=np.uint64)
        if n:
            got = int(self.L.bfcg_ch_export(self.ch, sub.ctypes.data_as(u32p), key.ctypes.data_as(u64p)))
            if got != n:
                raise BfcError("bfcg_ch_export failed: " + self.L.bfcg_last_error().decode())
        return sub, key

    def hist(self):
        cnt = np.zeros(256, dtype=np.uint64)
        high = np.zeros(64, dtype=np.uint64)
        mode = self.L.bfc_ch_hist(self.ch, cnt.ctypes.data_as(u64p), high.ctypes.data_as(u64p))
        return mode, cnt, high

    def mode(self) -> int:
        if self._mode is None:
            self._mode = self.hist()[0]
        return self._mode

    def n_distinct(self) -> int:
        return int(self.L.bfc_ch_count(self.ch))

    # ---- correct / trim
    def correct(self, seq: np.ndarray, qual, off: np.ndarray):
        """Returns (seq', qual', aux[2n]); the inputs are left untouched."""
        s = seq.copy()
        q = qual.copy() if qual is not None else None
        aux = np.zeros(2 * (len(off) - 1), dtype=np.uint32)
        b = host_batch(s, q, off)
        self.correct_batch(b, aux.ctypes.data)
        return s, q, aux

    def correct_batch(self, b: Batch, aux_ptr: int):
        _check(self.L.bfcg_correct_batch(C.byref(self.opt), self.ch, self.mode(), C.byref(b), aux_ptr,
                                         C.byref(self.stats)), "bfcg_correct_batch")

    def trim(self, seq: np.ndarray, off: np.ndarray):
        n = len(off) - 1
        keep = np.zeros(n, dtype=np.uint8)
        ts = np.zeros(n, dtype=np.int32)
        te = np.zeros(n, dtype=np.int32)
        b = host_batch(seq, None, off)
        self.trim_batch(b, keep.ctypes.data, ts.ctypes.data, te.ctypes.data)
        return keep, ts, te

    def trim_batch(self, b: Batch, keep_ptr: int, ts_ptr: int, te_ptr: int):
        _check(self.L.bfcg_trim_batch(C.byref(self.opt), self.bf_high, C.byref(b), keep_ptr, ts_ptr, te_ptr,
                                      C.byref(self.stats)), "bfcg_trim_batch")
