#!/usr/bin/env python
"""Quick device-resident timing of count / correct / trim at a chosen size (development aid;
bench.py is the contract).  Example:
    python tools/perf_probe.py --reads 4000000 --genome 20000000 --k 33 --b 37
"""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bfc_b200  # noqa: E402
from bfc_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--genome", type=int, default=20_000_000)
    ap.add_argument("--len", type=int, default=150)
    ap.add_argument("--k", type=int, default=33)
    ap.add_argument("--b", type=int, default=37)
    ap.add_argument("--trim", action="store_true")
    ap.add_argument("--no-correct", action="store_true")
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    L = api.lib()
    N, G, RL = a.reads, a.genome, a.len
    nb = N * (RL + 1)
    d_gen = L.bfcg_dev_alloc(G)
    d_seq = L.bfcg_dev_alloc(nb)
    d_qual = L.bfcg_dev_alloc(nb)
    d_off = L.bfcg_dev_alloc(8 * (N + 1))
    d_aux = L.bfcg_dev_alloc(8 * N)
    d_seq2 = L.bfcg_dev_alloc(nb)
    d_qual2 = L.bfcg_dev_alloc(nb)
    assert d_gen and d_seq and d_qual and d_off and d_aux and d_seq2 and d_qual2
    assert L.bfcg_synth_genome(d_gen, G, 2) == 0
    assert L.bfcg_synth_reads(d_gen, G, 2, 0, N, RL, 0.01, 2e-4, d_seq, d_qual, d_off) == 0
    opt = bfc_b200.make_opt(k=a.k, bf_shift=a.b, filter_mode=1 if a.trim else 0)
    e = bfc_b200.Engine(opt, timing=True)
    b = api.Batch()
    b.n_reads, b.n_bytes, b.where = N, nb, api.DEVICE
    b.off = C.cast(d_off, api.u64p)
    for rep in range(a.reps):
        e.reset()
        e.stats = api.Stats()
        api.kernel_times()
        b.seq, b.qual = C.cast(d_seq, api.u8p), C.cast(d_qual, api.u8p)
        t0 = time.time()
        e.count_batch(b)
        t1 = time.time()
        st = e.stats
        print(f"[rep {rep}] count: {t1 - t0:.3f} s  {N / (t1 - t0) / 1e6:.2f} Mreads/s  kmers={st.n_kmers} "
              f"pass={st.n_pass} ({st.n_pass / max(1, st.n_kmers):.3f}) pending={st.n_pending} conflict={st.n_conflict} "
              f"launches={st.n_launches}")
        if not a.trim:
            print(f"        distinct={e.n_distinct()} cap=2^{L.bfcg_ch_capacity_log2(e.ch)}")
        kt = api.kernel_times()
        print("        " + "  ".join(f"{k}={v[0]:.1f}ms/{v[1]}" for k, v in kt.items() if v[1]))
        if a.no_correct:
            continue
        # correct edits in place: work on a copy
        L.bfcg_sync()
        import ctypes
        cudart = ctypes.CDLL("libcudart.so.12")
        cudart.cudaMemcpy(C.c_void_p(d_seq2), C.c_void_p(d_seq), C.c_size_t(nb), 3)
        cudart.cudaMemcpy(C.c_void_p(d_qual2), C.c_void_p(d_qual), C.c_size_t(nb), 3)
        b.seq, b.qual = C.cast(d_seq2, api.u8p), C.cast(d_qual2, api.u8p)
        t0 = time.time()
        if a.trim:
            keep = L.bfcg_dev_alloc(N)
            ts = L.bfcg_dev_alloc(4 * N)
            te = L.bfcg_dev_alloc(4 * N)
            e.trim_batch(b, keep, ts, te)
            for p in (keep, ts, te):
                L.bfcg_dev_free(p)
        else:
            mode = e.mode()
            e.correct_batch(b, d_aux)
        t1 = time.time()
        st = e.stats
        print(f"[rep {rep}] {'trim' if a.trim else 'correct'}: {t1 - t0:.3f} s  {N / (t1 - t0) / 1e6:.2f} Mreads/s  "
              f"lookups={st.n_lookups} ({st.n_lookups / N:.1f}/read) redo={st.n_redo}")
        kt = api.kernel_times()
        print("        " + "  ".join(f"{k}={v[0]:.1f}ms/{v[1]}" for k, v in kt.items() if v[1]))
    e.close()


if __name__ == "__main__":
    main()
