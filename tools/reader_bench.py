#!/usr/bin/env python
"""Host-only throughput of the block reader / flat packer / writer (csrc/fqblock.c) next to the record-at-a-time
reader (csrc/bseq.c) on a synthetic four-line FASTQ in tmpfs.  No GPU needed.  Usage: reader_bench.py [n_reads] [threads]"""
import ctypes as C, os, sys, tempfile, time, shutil
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, bfc_b200
from test_fqblock import Block, Flat, Out, Bseq1, u32p
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else len(os.sched_getaffinity(0))
L = C.CDLL(bfc_b200.lib_path())
L.fq_open.restype = C.c_void_p; L.fq_open.argtypes = [C.c_char_p, C.c_int]
L.fq_next.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(Block)]
L.fq_block_free.argtypes = [C.POINTER(Block)]; L.fq_close.argtypes = [C.c_void_p]
L.fq_flat_fill.argtypes = [C.POINTER(Flat), C.POINTER(Block), C.POINTER(C.c_uint8), C.c_int]
L.fq_write.argtypes = [C.c_void_p, C.POINTER(Block), C.POINTER(Flat), C.POINTER(Out), C.c_int]
L.fq_flat_free.argtypes = [C.POINTER(Flat)]
L.bseq_open.restype = C.c_void_p; L.bseq_open.argtypes = [C.c_char_p]
L.bseq_read.restype = C.POINTER(Bseq1); L.bseq_read.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
L.bseq_close.argtypes = [C.c_void_p]
libc = C.CDLL(None); libc.fopen.restype = C.c_void_p; libc.fopen.argtypes = [C.c_char_p, C.c_char_p]; libc.fclose.argtypes = [C.c_void_p]
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
try:
    fq = os.path.join(d, "in.fq")
    rng = np.random.default_rng(1)
    with open(fq, "wb") as fp:
        for lo in range(0, n, 500_000):
            m = min(500_000, n - lo)
            s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(m, bench.READ_LEN))]
            q = rng.integers(35, 74, size=(m, bench.READ_LEN), dtype=np.uint8)
            fp.write(bench.fastq_fixed(s, q, lo))
    size = os.path.getsize(fq)
    t = {"read+split": 0.0, "pack": 0.0, "write": 0.0}
    f = L.fq_open(fq.encode(), threads)
    flat, got = Flat(), 0
    out = libc.fopen(b"/dev/null", b"wb")
    t_all = time.perf_counter()
    while True:
        b = Block()
        t0 = time.perf_counter()
        if L.fq_next(f, 1 << 30, 0, C.byref(b)) != 1:
            break
        t1 = time.perf_counter()
        assert L.fq_flat_fill(C.byref(flat), C.byref(b), None, threads) == 0
        t2 = time.perf_counter()
        aux = np.zeros(2 * b.n, dtype=np.uint32)
        o = Out(0, 0, 0, 0, aux.ctypes.data_as(u32p), None, None, None)
        t3 = time.perf_counter()
        assert L.fq_write(out, C.byref(b), C.byref(flat), C.byref(o), threads) == 0
        t4 = time.perf_counter()
        t["read+split"] += t1 - t0; t["pack"] += t2 - t1; t["write"] += t4 - t3
        got += b.n
        L.fq_block_free(C.byref(b))
    t_all = time.perf_counter() - t_all
    libc.fclose(out); L.fq_flat_free(C.byref(flat)); L.fq_close(f)
    assert got == n
    print(f"{n} reads, {size / 1e9:.2f} GB, {threads} threads")
    for k, v in t.items():
        print(f"  block path  {k:10s} {v:6.2f} s  {n / v / 1e6:7.1f} Mreads/s  {size / v / 1e9:5.1f} GB/s")
    t0 = time.perf_counter()
    f = L.bseq_open(fq.encode()); tot = 0
    while True:
        m = C.c_int(0)
        seqs = L.bseq_read(f, 100_000_000, 0, C.byref(m))
        if not seqs or m.value == 0:
            break
        tot += m.value
    dt = time.perf_counter() - t0
    L.bseq_close(f)
    print(f"  record-at-a-time reader (bseq_read, records not freed) {dt:6.2f} s  {tot / dt / 1e6:7.1f} Mreads/s")
finally:
    shutil.rmtree(d, ignore_errors=True)
