#!/usr/bin/env python
"""Timeline of the drop-in CLI on the GPU box: fixed start-up cost (a 1000-read input) and the stamps of a big run."""
import os, re, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from bfc_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
L = api.lib()
d = tempfile.mkdtemp(dir="/dev/shm")
try:
    G, RB = bench.genome_size(n), bench.READ_LEN + 1
    d_gen, d_off = L.bfcg_dev_alloc(G), L.bfcg_dev_alloc(8 * 1_000_001)
    assert L.bfcg_synth_genome(d_gen, G, bench.SEED) == 0
    fq, tiny = os.path.join(d, "in.fq"), os.path.join(d, "tiny.fq")
    with open(fq, "wb") as fp:
        for lo in range(0, n, 1_000_000):
            m = min(1_000_000, n - lo)
            d_s, d_q = L.bfcg_dev_alloc(m * RB), L.bfcg_dev_alloc(m * RB)
            assert L.bfcg_synth_reads(d_gen, G, bench.SEED, lo, m, bench.READ_LEN, bench.ERR, bench.N_RATE, d_s, d_q, d_off) == 0
            hs, hq = np.empty(m * RB, dtype=np.uint8), np.empty(m * RB, dtype=np.uint8)
            L.bfcg_d2h(hs.ctypes.data, d_s, m * RB); L.bfcg_d2h(hq.ctypes.data, d_q, m * RB)
            L.bfcg_dev_free(d_s); L.bfcg_dev_free(d_q)
            fp.write(bench.fastq_fixed(hs.reshape(m, RB)[:, :-1], hq.reshape(m, RB)[:, :-1], lo))
    with open(fq, "rb") as f, open(tiny, "wb") as g:
        g.write(b"".join(f.readline() for _ in range(4000)))
    exe = os.path.join(ROOT, "bfc_b200", "lib", "bfc")
    envs = [e for e in os.environ.get("TIMELINE_ENVS", "").split(";") if e]       # e.g. "BFC_B200_EC_BATCH=40000000;BFC_B200_EC_BATCH=24000000"
    runs = [("tiny", tiny, ""), ("full", fq, "")] + [("full", fq, e) for e in envs] + [("full", fq, "")]
    for name, src, extra in runs:
        t0 = time.time()
        env = dict(os.environ)
        for kv in extra.split(","):
            if "=" in kv:
                env[kv.split("=")[0]] = kv.split("=")[1]
        name = name + ("_" + extra.replace("=", "-").replace(",", "_") if extra else "")
        p = subprocess.run([exe] + os.environ.get("TIMELINE_FLAGS", "-k 33").split() + ["-b", "37", "-t", "16", "-V", "4", src], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env)
        dt = time.time() - t0
        print(f"== {name}: wall {dt:.2f} s")
        lines = p.stderr.splitlines()
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"cli_timeline_{name}_{time.time():.1f}.log"), "w") as lf:
            lf.write(p.stderr)
        keep = [l for l in lines if "@" in l or "Real time" in l]
        for l in (keep if len(keep) <= 40 else keep[:8] + ["..."] + keep[-3:]):
            print("   ", l[:200])
finally:
    shutil.rmtree(d, ignore_errors=True)
