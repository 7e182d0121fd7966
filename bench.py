#!/usr/bin/env python
"""bench.py -- count + correct throughput of the B200 engine (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the unmodified reference on the host cores

Workload (BASELINE.json configs[2]): synthetic 150 bp reads at 30x (1 % substitutions,
Q-correlated, rare N), k = 33, Bloom 2^37 bits (`-s 3g`), count + correct.  Default
100 M reads from a 500 Mb genome; `--reads` scales both (coverage stays 30x).

A *step* is one complete job: empty filter/table -> count every read -> histogram ->
correct every read.  `value` times K steps with the reads already resident in HBM
(CUDA events on the engine's stream, max over ranks); `e2e` times the same job through
the C ABI with HOST buffers (H2D of every batch and D2H of the corrected reads inside
the timed region).  N > 1 (one process per GPU under torchrun): the k-mers are sharded
by Bloom-block prefix -- every rank enumerates its pieces of the reads, one all-to-all
per global chunk delivers the 16-byte records to their owners in global read order,
the table shards are all-gathered and the correction is partitioned by reads
(bfc_b200/dist.py); total work is fixed => "strong".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN = 150
COVERAGE = 30.0
ERR = 0.01
N_RATE = 2e-4
SEED = 2


def genome_size(n_reads: int) -> int:
    return max(100_000, int(n_reads * READ_LEN / COVERAGE))


# ----------------------------------------------------------------------------- clocks

class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm

def fastq_fixed(seq: np.ndarray, qual: np.ndarray, first: int) -> bytes:
    """Vectorised FASTQ writer with fixed-width names (@r%09d)."""
    n, L = seq.shape
    rec = np.empty((n, 1 + 10 + 1 + L + 3 + L + 1), dtype=np.uint8)
    rec[:, 0] = ord("@")
    rec[:, 1] = ord("r")
    idx = np.arange(first, first + n, dtype=np.int64)
    for d in range(9):
        rec[:, 2 + d] = (idx // 10 ** (8 - d)) % 10 + ord("0")
    rec[:, 11] = ord("\n")
    rec[:, 12:12 + L] = seq
    rec[:, 12 + L:15 + L] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, 15 + L:15 + 2 * L] = qual
    rec[:, 15 + 2 * L] = ord("\n")
    return rec.tobytes()


def write_sample_fastq(path: str, n_reads: int) -> int:
    from bfc_b200 import synth
    G = genome_size(n_reads)
    with open(path, "wb") as fp:
        for lo in range(0, n_reads, 250_000):
            n = min(250_000, n_reads - lo)
            s, q = synth.cb_reads(G, SEED, lo, n, READ_LEN, ERR, N_RATE)
            fp.write(fastq_fixed(s, q, lo))
    return G


def run_reference_once(fq: str, k: int, b: int, threads: int) -> float:
    exe = os.path.join(ROOT, "oracle", "_ref", "bfc")
    t0 = time.time()
    with open(os.devnull, "wb") as null:
        subprocess.run([exe, "-k", str(k), "-b", str(b), "-t", str(threads), fq], stdout=null, stderr=null, check=True)
    return time.time() - t0


def reference_sample_reads(steps: int, warmup: int) -> int:
    return int(os.environ.get("BFC_BENCH_REF_READS", max(200_000, 4_000_000 // max(1, steps + warmup))))


def reference_arm(args) -> dict:
    """The unmodified reference (oracle/_ref/bfc, built from /root/reference by oracle/Makefile) on the
    host cores, all threads, on a bounded sample of the same workload."""
    exe = os.path.join(ROOT, "oracle", "_ref", "bfc")
    threads = os.cpu_count() or 1
    n = reference_sample_reads(args.steps, args.warmup)
    tmpdir = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        fq = os.path.join(tmpdir, "sample.fq")
        G = write_sample_fastq(fq, n)
        if not os.path.exists(exe):
            raise RuntimeError("oracle/_ref/bfc is missing (built by `make -C oracle ref` where /root/reference exists)")
        for _ in range(args.warmup):
            run_reference_once(fq, args.k, args.bf_shift, threads)
        times = [run_reference_once(fq, args.k, args.bf_shift, threads) for _ in range(args.steps)]
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)
    tot = sum(times)
    v = n * len(times) / tot / 1e6
    sample = (f"{n} reads x {READ_LEN} bp from a {G} bp genome ({COVERAGE:.0f}x), same generator; "
              f"`bfc -k {args.k} -b {args.bf_shift} -t {threads}` FASTQ on tmpfs -> /dev/null, wall clock incl. parsing")
    return {"value": v, "unit": "Mreads/s", "cores": threads, "kind": "reference", "sample": sample,
            "ms_per_step": 1e3 * tot / len(times)}


# ----------------------------------------------------------------------------- this repo's arm

class DeviceData:
    """The synthetic read set resident in HBM: pristine seq/qual + a working copy for in-place correction."""

    def __init__(self, L, api, n_reads: int):
        self.L, self.api, self.n = L, api, n_reads
        self.G = genome_size(n_reads)
        self.nb = n_reads * (READ_LEN + 1)
        self.d_gen = L.bfcg_dev_alloc(self.G)
        self.d_seq, self.d_qual = L.bfcg_dev_alloc(self.nb), L.bfcg_dev_alloc(self.nb)
        self.d_off = L.bfcg_dev_alloc(8 * (n_reads + 1))
        if not (self.d_gen and self.d_seq and self.d_qual and self.d_off):
            raise RuntimeError("device allocation failed: " + L.bfcg_last_error().decode())
        assert L.bfcg_synth_genome(self.d_gen, self.G, SEED) == 0
        assert L.bfcg_synth_reads(self.d_gen, self.G, SEED, 0, n_reads, READ_LEN, ERR, N_RATE, self.d_seq, self.d_qual, self.d_off) == 0

    def batch(self, seq_ptr, qual_ptr, r0: int, r1: int):
        """Device batch over reads [r0, r1) of the given seq/qual buffers (absolute offsets)."""
        api = self.api
        b = api.Batch()
        b.n_reads, b.where = r1 - r0, api.DEVICE
        b.n_bytes = (r1 - r0) * (READ_LEN + 1)
        b.off = C.cast(self.d_off + 8 * r0, api.u64p)
        b.seq, b.qual = C.cast(seq_ptr, api.u8p), C.cast(qual_ptr, api.u8p)
        return b


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=int(os.environ.get("BFC_BENCH_READS", 100_000_000)))
    ap.add_argument("--k", type=int, default=33)
    ap.add_argument("--bf-shift", type=int, default=37)
    ap.add_argument("--chunk-reads", type=int, default=16_000_000, help="N > 1: reads per global chunk (one all-to-all each)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    config = {"workload": f"count+correct: {args.reads} x {READ_LEN} bp synthetic reads at {COVERAGE:.0f}x, k={args.k}, "
                          f"Bloom 2^{args.bf_shift} bits, H=4, min_cov=3 (BASELINE configs[2], `-s 3g`)",
              "reads": args.reads, "read_len": READ_LEN, "genome": genome_size(args.reads), "k": args.k,
              "bf_shift": args.bf_shift, "parallelism": f"k-mers sharded by Bloom-block prefix x{world} (all-to-all per chunk), table all-gathered, correction partitioned by reads" if world > 1 else "single GPU",
              "l2": "inputs (>= 30 GB) and filter/table (>= 32 GB) exceed the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = reference_arm(args)
        line = {"impl": "reference", "metric": "count+correct throughput", "value": r["value"], "unit": "Mreads/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": config, "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "Mreads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import bfc_b200
    from bfc_b200 import api
    L = api.lib()
    if L.bfcg_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert L.bfcg_set_device(local_rank) == 0, L.bfcg_last_error()

    def barrier():
        L.bfcg_sync()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.reads
    cudart = C.CDLL("libcudart.so.12")
    opt = bfc_b200.make_opt(k=args.k, bf_shift=args.bf_shift)
    RB = READ_LEN + 1

    if world == 1:
        data = DeviceData(L, api, n)
        r0, r1 = 0, n
        nb_mine = n * RB
        w_seq, w_qual = L.bfcg_dev_alloc(max(1, nb_mine)), L.bfcg_dev_alloc(max(1, nb_mine))
        d_aux = L.bfcg_dev_alloc(max(8, 8 * n))
        assert w_seq and w_qual and d_aux
        eng = bfc_b200.Engine(opt, timing=True)
        count_b = data.batch(data.d_seq, data.d_qual, 0, n)
        work_b = data.batch(w_seq, w_qual, 0, n)
        src_seq, src_qual, n_mine = data.d_seq, data.d_qual, n
        pieces = [(0, n)]

        def step():
            eng.reset()
            cudart.cudaMemcpy(C.c_void_p(w_seq), C.c_void_p(src_seq), C.c_size_t(nb_mine), 3)
            cudart.cudaMemcpy(C.c_void_p(w_qual), C.c_void_p(src_qual), C.c_size_t(nb_mine), 3)
            eng.count_batch(count_b)
            eng.correct_batch(work_b, d_aux)

        def get_stats():
            return eng.stats.as_dict()

        def clear_stats():
            eng.stats = api.Stats()
    else:
        # rank r holds the r-th piece of every global chunk (bfc_b200/dist.py): local reads = its pieces back to back
        from bfc_b200.dist import CudaBackend, ShardedCount, piece_bounds
        G = genome_size(n)
        pieces = [piece_bounds(lo, min(n, lo + args.chunk_reads), rank, world) for lo in range(0, n, args.chunk_reads)]
        n_mine = sum(p1 - p0 for p0, p1 in pieces)
        nb_mine = n_mine * RB
        d_gen = L.bfcg_dev_alloc(G)
        src_seq, src_qual = L.bfcg_dev_alloc(max(1, nb_mine)), L.bfcg_dev_alloc(max(1, nb_mine))
        w_seq, w_qual = L.bfcg_dev_alloc(max(1, nb_mine)), L.bfcg_dev_alloc(max(1, nb_mine))
        d_off, d_off_tmp = L.bfcg_dev_alloc(8 * (n_mine + 1)), L.bfcg_dev_alloc(8 * (max(p1 - p0 for p0, p1 in pieces) + 1))
        d_aux = L.bfcg_dev_alloc(max(8, 8 * n_mine))
        assert d_gen and src_seq and src_qual and w_seq and w_qual and d_off and d_off_tmp and d_aux, L.bfcg_last_error()
        assert L.bfcg_synth_genome(d_gen, G, SEED) == 0
        loc = 0
        piece_loc = []
        for p0, p1 in pieces:
            assert L.bfcg_synth_reads(d_gen, G, SEED, p0, p1 - p0, READ_LEN, ERR, N_RATE, src_seq + loc * RB, src_qual + loc * RB, d_off_tmp) == 0
            piece_loc.append(loc)
            loc += p1 - p0
        h_off_local = (np.arange(n_mine + 1, dtype=np.uint64) * np.uint64(RB))
        L.bfcg_h2d(d_off, h_off_local.ctypes.data, 8 * (n_mine + 1))
        L.bfcg_dev_free(d_off_tmp)
        be = CudaBackend(opt, world, local_rank)
        L.bfcg_set_timing(1)
        sc = ShardedCount(be, rank, world)

        def dev_batch(seq_ptr, qual_ptr, first, count):
            b = api.Batch()
            b.n_reads, b.where, b.n_bytes = count, api.DEVICE, count * RB
            b.off = C.cast(d_off, api.u64p)   # offsets relative to the pointers below
            b.seq, b.qual = C.cast(seq_ptr + first * RB, api.u8p), C.cast(qual_ptr + first * RB, api.u8p)
            return b

        count_pieces = [dev_batch(src_seq, src_qual, piece_loc[i], p1 - p0) for i, (p0, p1) in enumerate(pieces)]
        work_b = dev_batch(w_seq, w_qual, 0, n_mine)
        r0, r1 = 0, n_mine

        def step():
            t0 = time.perf_counter()
            be.reset()
            cudart.cudaMemcpy(C.c_void_p(w_seq), C.c_void_p(src_seq), C.c_size_t(nb_mine), 3)
            cudart.cudaMemcpy(C.c_void_p(w_qual), C.c_void_p(src_qual), C.c_size_t(nb_mine), 3)
            t1 = time.perf_counter()
            for b in count_pieces:
                sc.count_piece(b)
            L.bfcg_sync()
            t2 = time.perf_counter()
            sc.gather()
            L.bfcg_sync()
            t3 = time.perf_counter()
            if n_mine:
                be.correct_batch(work_b, d_aux)
            if sc.prof is not None and rank == 0:
                print("[dist profile] reset+copy %.3f count %.3f gather %.3f correct %.3f s; phases %s" % (
                    t1 - t0, t2 - t1, t3 - t2, time.perf_counter() - t3, {k_: round(v, 3) for k_, v in sc.prof.items()}), file=sys.stderr)
                sc.prof.clear()

        def get_stats():
            return be.stats.as_dict()

        def clear_stats():
            be.stats = api.Stats()

    for _ in range(args.warmup):
        step()
    barrier()
    clear_stats()
    api.kernel_times()
    clocks = ClockSampler(local_rank)
    clocks.start()
    L.bfcg_event_record(0)
    for _ in range(args.steps):
        step()
    L.bfcg_sync()
    L.bfcg_event_record(1)
    L.bfcg_sync()
    ms = L.bfcg_event_elapsed_ms(0, 1)
    barrier()
    clk = clocks.stop()
    ms = max_over_ranks(ms)
    st = get_stats()
    kt = api.kernel_times()
    value = n * args.steps / (ms / 1e3) / 1e6
    if world > 1:  # whole-job counters for the roofline arithmetic and the stats block
        import torch
        keys = ["n_kmers", "n_pass", "n_pending", "n_conflict", "n_lookups", "n_search_lookups", "n_redo", "n_launches"]
        t = torch.tensor([st[k_] for k_ in keys], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        launches_rank0 = st["n_launches"]
        st.update({k_: int(v) for k_, v in zip(keys, t.cpu())})
        st["n_launches"] = launches_rank0

    # ---- roofline of the dominant kernel (algorithmic bytes: DESIGN.md "Kernels")
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    count_bytes = 2 * n * RB * args.steps + 64 * st["n_kmers"] + 16 * st["n_pass"]
    # k_ec_search: a 32-byte sector per lookup it makes itself + 4 B per base of flags / planes in and edits out;
    # k_ec_lookup: one lookup per k-mer of every read (bfc_ec_kcov) + seq and qual in + a byte of planes out per position
    correct_bytes = 32 * st["n_search_lookups"] + 4 * n * RB * args.steps
    lookup_bytes = 32 * st["n_kmers"] + 3 * n * RB * args.steps
    kern = {}
    for name, alg in (("count_probe", count_bytes), ("count_part", count_bytes), ("correct", correct_bytes), ("ec_lookup", lookup_bytes)):
        t_ms, launches = kt.get(name, (0.0, 0))
        if launches:
            kern[name] = {"ms": t_ms, "launches": launches, "algorithmic_bytes": alg, "GBps": alg / (t_ms / 1e3) / 1e9,
                          "share_of_step": t_ms / ms}
    for name, (t_ms, launches) in kt.items():
        if launches and name not in kern:
            kern[name] = {"ms": t_ms, "launches": launches, "share_of_step": t_ms / ms}
    dom = max(("count_probe", "count_part", "correct"), key=lambda k_: kern.get(k_, {}).get("ms", 0.0))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom, {}).get("dram_bytes_per_launch")
    d = kern.get(dom, {"GBps": 0.0, "launches": 1, "algorithmic_bytes": 0})
    kernel_name = {"count_probe": "k_count_probe", "count_part": "k_count_part", "correct": "k_ec_search"}[dom]
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": d["GBps"], "peak": peak, "unit": "GB/s",
                "frac": d["GBps"] / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": d["algorithmic_bytes"] / max(1, d["launches"]),
                "kernels": kern}

    # ---- end to end through the C ABI with host buffers: this rank's reads start in pinned host memory, every
    # count piece and the correction copy them to the device inside the timed region, the corrected reads and
    # the per-read stats come back to host memory
    e2e = None
    if not args.no_e2e:
        e2e_steps = min(args.steps, 2)
        def pinned_u8(nbytes):  # pinned host memory (what a pipelined host would stage batches in)
            p = L.bfcg_host_alloc_pinned(max(1, nbytes))
            if not p:
                raise SystemExit("bench.py: pinned host allocation failed: " + L.bfcg_last_error().decode())
            return p, np.ctypeslib.as_array(C.cast(p, api.u8p), shape=(max(1, nbytes),))[:nbytes]
        p_hs, h_seq = pinned_u8(nb_mine)
        p_hq, h_qual = pinned_u8(nb_mine)
        p_ws, ws = pinned_u8(nb_mine)
        p_wq, wq = pinned_u8(nb_mine)
        L.bfcg_d2h(h_seq.ctypes.data, src_seq, nb_mine)
        L.bfcg_d2h(h_qual.ctypes.data, src_qual, nb_mine)
        h_off = np.arange(n_mine + 1, dtype=np.uint64) * np.uint64(RB)
        p_aux = np.empty(2 * max(1, n_mine), dtype=np.uint32)
        hb_work = api.host_batch(ws, wq, h_off)
        hb_pieces, loc = [], 0
        for p0, p1 in pieces:
            m = p1 - p0
            hb_pieces.append(api.host_batch(h_seq[loc * RB:(loc + m) * RB], h_qual[loc * RB:(loc + m) * RB], h_off[:m + 1]))
            loc += m
        tot = 0.0
        split = {"reset_s": 0.0, "count_s": 0.0, "correct_s": 0.0}
        for it in range(1 + e2e_steps):
            ws[:] = h_seq
            wq[:] = h_qual
            barrier()
            t0 = time.perf_counter()
            if world == 1:
                eng.reset()
                t1 = time.perf_counter()
                eng.count_batch(hb_pieces[0])
                t2 = time.perf_counter()
                eng.correct_batch(hb_work, p_aux.ctypes.data)
                if it > 0:
                    split["reset_s"] += t1 - t0; split["count_s"] += t2 - t1; split["correct_s"] += time.perf_counter() - t2
            else:
                be.reset()
                for hb in hb_pieces:
                    sc.count_piece(hb)
                sc.gather()
                if n_mine:
                    be.correct_batch(hb_work, p_aux.ctypes.data)
            L.bfcg_sync()
            dt = max_over_ranks(time.perf_counter() - t0)
            if it > 0:
                tot += dt
        e2e = {"value": n * e2e_steps / tot / 1e6, "unit": "Mreads/s",
               "h2d_bytes_per_step": int(4 * n * RB + 8 * (n + world)),
               "d2h_bytes_per_step": int(2 * n * RB + 8 * n),
               "steps": e2e_steps, "split": {k_: v / e2e_steps for k_, v in split.items()} if world == 1 else None, "note": "whole job, all ranks: pinned host buffers -> bfcg_count_batch (N>1: bfcg_enum_records + "
               "all-to-all + bfcg_count_records) / bfcg_correct_batch -> host buffers; wall clock, max over ranks"}
        # full-size consistency check (outside every timed region): the reads corrected through host batches (many
        # windows, copies overlapped with the kernels) equal, byte for byte, those corrected while resident in HBM
        same = True
        if nb_mine:
            L.bfcg_d2h(h_seq.ctypes.data, w_seq, nb_mine)
            L.bfcg_d2h(h_qual.ctypes.data, w_qual, nb_mine)
            d_aux_h = np.empty(2 * n_mine, dtype=np.uint32)
            L.bfcg_d2h(d_aux_h.ctypes.data, d_aux, 8 * n_mine)
            same = bool(np.array_equal(h_seq, ws) and np.array_equal(h_qual, wq) and np.array_equal(d_aux_h, p_aux[:2 * n_mine]))
        e2e["equals_resident_result"] = bool(max_over_ranks(0.0 if same else 1.0) == 0.0)
        for p_ in (p_hs, p_hq, p_ws, p_wq):
            L.bfcg_host_free_pinned(p_)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            class A: pass
            a = A(); a.steps, a.warmup, a.k, a.bf_shift = 1, 0, args.k, args.bf_shift
            os.environ.setdefault("BFC_BENCH_REF_READS", "1000000")
            r = reference_arm(a)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:  # the baseline is reported, never required
            cpu = {"value": None, "unit": "Mreads/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}

    if rank == 0:
        line = {"metric": "count+correct throughput", "value": value, "unit": "Mreads/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
                "clocks": clk, "e2e": e2e, "gpu_launches": int(st["n_launches"]), "roofline": roofline,
                "cpu_baseline": cpu,
                "stats": {"kmers_per_step": st["n_kmers"] // args.steps, "f_pass": st["n_pass"] / max(1, st["n_kmers"]),
                          "pending_frac": st["n_pending"] / max(1, st["n_kmers"]),
                          "conflict_frac": st["n_conflict"] / max(1, st["n_kmers"]),
                          "lookups_per_read": st["n_lookups"] / max(1, n * args.steps),
                          "search_lookups_per_read": st["n_search_lookups"] / max(1, n * args.steps), "redo": st["n_redo"]}}
        print(json.dumps(line))
    if world == 1:
        eng.close()
    else:
        be.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
