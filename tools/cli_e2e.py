#!/usr/bin/env python
"""End-to-end timing of the drop-in CLI (FASTQ file on tmpfs -> corrected FASTQ to /dev/null) next to the reference
binary on the same file.  Usage: cli_e2e.py [n_reads] [threads]"""
import os, subprocess, sys, tempfile, time, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
t = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
try:
    fq = os.path.join(d, "in.fq")
    t0 = time.time()
    # the device generator of bench.py (csrc/synth.cu; same reads as bfc_b200/synth.py, which is its slow CPU twin)
    import ctypes as C, numpy as np
    from bfc_b200 import api
    L = api.lib()
    G, RB = bench.genome_size(n), bench.READ_LEN + 1
    d_gen, d_off = L.bfcg_dev_alloc(G), L.bfcg_dev_alloc(8 * 1_000_001)
    assert L.bfcg_synth_genome(d_gen, G, bench.SEED) == 0
    with open(fq, "wb") as fp:
        for lo in range(0, n, 1_000_000):
            m = min(1_000_000, n - lo)
            d_s, d_q = L.bfcg_dev_alloc(m * RB), L.bfcg_dev_alloc(m * RB)
            assert L.bfcg_synth_reads(d_gen, G, bench.SEED, lo, m, bench.READ_LEN, bench.ERR, bench.N_RATE, d_s, d_q, d_off) == 0
            hs, hq = np.empty(m * RB, dtype=np.uint8), np.empty(m * RB, dtype=np.uint8)
            L.bfcg_d2h(hs.ctypes.data, d_s, m * RB); L.bfcg_d2h(hq.ctypes.data, d_q, m * RB)
            L.bfcg_dev_free(d_s); L.bfcg_dev_free(d_q)
            fp.write(bench.fastq_fixed(hs.reshape(m, RB)[:, :-1], hq.reshape(m, RB)[:, :-1], lo))
    L.bfcg_dev_free(d_gen); L.bfcg_dev_free(d_off)
    print(f"generated {n} reads ({os.path.getsize(fq) / 1e9:.2f} GB) in {time.time() - t0:.1f} s")
    for name, exe, m in (("bfc_b200", os.path.join(ROOT, "bfc_b200", "lib", "bfc"), n), ("reference", os.path.join(ROOT, "oracle", "_ref", "bfc"), min(n, 1_000_000))):
        src = fq
        if m < n:
            src = os.path.join(d, "sub.fq")
            with open(fq, "rb") as f, open(src, "wb") as g:
                g.write(b"".join(f.readline() for _ in range(4 * m)))
        t0 = time.time()
        p = subprocess.run([exe, "-k", "33", "-b", "37", "-t", str(t), src], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        dt = time.time() - t0
        print(f"{name}: {m} reads in {dt:.2f} s = {m / dt / 1e6:.3f} Mreads/s (rc {p.returncode}, -t {t})")
        lines = p.stderr.strip().splitlines()
        print("   " + "\n   ".join(lines if name == "bfc_b200" and len(lines) < 60 else lines[-4:]))
finally:
    shutil.rmtree(d, ignore_errors=True)
