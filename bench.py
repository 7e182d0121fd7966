#!/usr/bin/env python
"""bench.py -- throughput of the B200 count / correct / trim engine (BASELINE.json metric: Mreads/s of 150 bp reads).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the unmodified reference on the host cores
    python bench.py --workload count|k55|trim ...            # the other single-GPU lines of BASELINE.json

Workloads (synthetic 150 bp reads at 30x, 1 % substitutions with Q-correlated qualities, rare N):
    count+correct  BASELINE configs[2]: k = 33, `-s 3g` => Bloom 2^37 bits, count + correct      (the default, the headline)
    count          BASELINE configs[1]: the count phase of the same job alone (`bfc -E`)
    k55            BASELINE configs[3] at the single-GPU size: k = 55 (l_pre -> 24, lossy key fold, 16-byte records)
    trim           BASELINE configs[4] at the single-GPU size: `-1` k = 51, two 2^37-bit filters, Bloom-only lookups
Default 100 M reads from a 500 Mb genome; `--reads` scales both (coverage stays 30x).

A *step* is one complete job: empty filter(s)/table -> count every read -> (histogram -> correct | trim) every read.
`value` times K steps with the reads already resident in HBM (CUDA events on the engine's stream, max over ranks);
`e2e` times the same job through the C ABI with HOST buffers (H2D of every batch and D2H of the results inside the
timed region); `e2e_cli` is the drop-in command line, FASTQ file -> FASTQ on stdout.  N > 1 (one process per GPU
under torchrun): the k-mers are sharded by Bloom-block prefix (bfc_b200/dist.py); total work
is fixed => "strong".

The reference arm (`--impl reference`, also the `cpu_baseline` of the default arm) runs the UNMODIFIED reference
binary (oracle/_ref/bfc, all host threads) on a bounded sample of the same generator: a fixed number of reads
(independent of --steps) from a genome that keeps the coverage, with the Bloom filter sized by the reference's own
`-s` rule for THAT genome (bfc.c:42-53) -- a 2^37-bit filter would cost a fixed ~6 s of zeroing per run and swamp a
bounded sample.  Its `config` says what it ran.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN = 150
COVERAGE = 30.0
ERR = 0.01
N_RATE = 2e-4
SEED = 2

WORKLOADS = {
    "count+correct": dict(k=33, filter_mode=0, correct=True, metric="count+correct throughput", ref_flags=[],
                          what="count+correct", baseline="BASELINE configs[2], `-s 3g`"),
    "count": dict(k=33, filter_mode=0, correct=False, metric="count throughput", ref_flags=["-E"],
                  what="count only (`bfc -E`)", baseline="BASELINE configs[1]"),
    "k55": dict(k=55, filter_mode=0, correct=True, metric="count+correct throughput (k=55)", ref_flags=[],
                what="count+correct, k=55", baseline="BASELINE configs[3] at the single-GPU size"),
    "trim": dict(k=51, filter_mode=1, correct=True, metric="trim-mode throughput (-1, k=51)", ref_flags=["-1"],
                 what="`-1` count + trim, k=51", baseline="BASELINE configs[4] at the single-GPU size"),
}


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


REF_SAMPLE_READS = 500_000     # fixed: the reference arm's rate must not depend on --steps / --warmup
REF_MAX_WARMUP = 1             # a host binary needs one run to warm the page cache, not five


def ref_bf_shift(G: int) -> int:
    """The reference's `-s` rule for the Bloom filter (bfc.c:42-53): b = floor(log2(size) + 8), at most 37."""
    return min(37, int(math.log2(G) + 8))


def reference_arm(args, wl, steps: int, warmup: int) -> dict:
    """The unmodified reference (oracle/_ref/bfc, built from /root/reference by oracle/Makefile) on the host cores,
    all threads, `steps` timed runs over one bounded sample of the workload."""
    exe = os.path.join(ROOT, "oracle", "_ref", "bfc")
    threads = os.cpu_count() or 1
    n = int(os.environ.get("BFC_BENCH_REF_READS", REF_SAMPLE_READS))
    tmpdir = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        fq = os.path.join(tmpdir, "sample.fq")
        G = write_sample_fastq(fq, n)
        b = ref_bf_shift(G)
        if not os.path.exists(exe):
            raise RuntimeError("oracle/_ref/bfc is missing (built by `make -C oracle ref` where /root/reference exists)")
        cmd = [exe] + wl["ref_flags"] + ["-k", str(wl["k"]), "-b", str(b), "-t", str(threads), fq]

        def once() -> float:
            t0 = time.time()
            with open(os.devnull, "wb") as null:
                subprocess.run(cmd, stdout=null, stderr=null, check=True)
            return time.time() - t0
        for _ in range(min(warmup, REF_MAX_WARMUP)):
            once()
        times = [once() for _ in range(steps)]
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)
    tot = sum(times)
    v = n * len(times) / tot / 1e6
    sample = (f"{n} reads x {READ_LEN} bp from a {G} bp genome ({COVERAGE:.0f}x), same generator; "
              f"`bfc {' '.join(wl['ref_flags'] + ['-k', str(wl['k']), '-b', str(b), '-t', str(threads)])}` (filter sized by the "
              f"reference's -s rule for the sample's genome), FASTQ on tmpfs -> /dev/null, wall clock incl. parsing and printing")
    return {"value": v, "unit": "Mreads/s", "cores": threads, "kind": "reference", "sample": sample,
            "ms_per_step": 1e3 * tot / len(times), "reads": n, "genome": G, "bf_shift": b}


# ----------------------------------------------------------------------------- drop-in CLI, file to stdout

def cli_e2e(L, api, wl, n_reads: int) -> dict:
    """lib/bfc on a FASTQ file (tmpfs) -> FASTQ on stdout (/dev/null): what a user of the reference's command line
    sees, parsing and printing included.  Runs in a child process, after this process has released the GPU memory."""
    exe = os.path.join(ROOT, "bfc_b200", "lib", "bfc")
    G, RB = genome_size(n_reads), READ_LEN + 1
    d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        fq = os.path.join(d, "in.fq")
        d_gen, d_off = L.bfcg_dev_alloc(G), L.bfcg_dev_alloc(8 * 1_000_001)
        assert d_gen and d_off and L.bfcg_synth_genome(d_gen, G, SEED) == 0
        with open(fq, "wb") as fp:
            for lo in range(0, n_reads, 1_000_000):
                m = min(1_000_000, n_reads - lo)
                d_s, d_q = L.bfcg_dev_alloc(m * RB), L.bfcg_dev_alloc(m * RB)
                assert L.bfcg_synth_reads(d_gen, G, SEED, lo, m, READ_LEN, ERR, N_RATE, d_s, d_q, d_off) == 0
                hs, hq = np.empty(m * RB, dtype=np.uint8), np.empty(m * RB, dtype=np.uint8)
                L.bfcg_d2h(hs.ctypes.data, d_s, m * RB); L.bfcg_d2h(hq.ctypes.data, d_q, m * RB)
                L.bfcg_dev_free(d_s); L.bfcg_dev_free(d_q)
                fp.write(fastq_fixed(hs.reshape(m, RB)[:, :-1], hq.reshape(m, RB)[:, :-1], lo))
        L.bfcg_dev_free(d_gen); L.bfcg_dev_free(d_off)
        threads = os.cpu_count() or 1
        cmd = [exe] + wl["ref_flags"] + ["-k", str(wl["k"]), "-b", "37", "-t", str(threads), fq]
        t0 = time.time()
        subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)   # warm-up: a fresh box backs
        cold = time.time() - t0                             # its memory on first touch (seconds for the first process that pins GBs)
        runs = []
        for _ in range(2):   # start-up (CUDA context + the 16 GiB filter) varies between 0.3 and 1 s on one and the same box: best of two
            t0 = time.time()
            r = subprocess.run(cmd[:1] + ["-V", "4"] + cmd[1:], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, check=True)
            runs.append(time.time() - t0)
        dt = min(runs)
        log = os.environ.get("BFC_BENCH_CLI_LOG")
        if log:
            with open(log, "wb") as fp:
                fp.write(r.stderr)
        return {"value": n_reads / dt / 1e6, "unit": "Mreads/s", "reads": n_reads, "seconds": dt, "runs_seconds": runs, "first_run_seconds": cold, "threads": threads,
                "what": f"`lib/bfc {' '.join(cmd[1:-1])}` {os.path.getsize(fq) / 1e9:.2f} GB FASTQ on tmpfs -> stdout (/dev/null), wall clock of the whole process"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


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

    def free(self):
        for p in (self.d_gen, self.d_seq, self.d_qual, self.d_off):
            self.L.bfcg_dev_free(p)


def kernel_accounting(wl, st, kt, n, RB, steps, ms, peak):
    """Per-kernel device time next to the ALGORITHMIC bytes of what it did:
    the counts come from the kernels' own counters (k-mers, passing occurrences, lookups)."""
    pos = n * RB * steps                      # stream positions (bases + terminators)
    nk, npass, npend = st["n_kmers"], st["n_pass"], st["n_pending"]
    rec = 9 if wl["k"] <= 35 else 10 if wl["k"] <= 39 else 12 if wl["k"] <= 47 else 16   # bytes of a packed record
    n_search = st["n_search_lookups"]
    n_kcov = nk if wl["correct"] and not wl["filter_mode"] else 0     # bfc_ec_kcov: one lookup per k-mer of every read
    n_ext = max(0, st["n_lookups"] - n_search - n_kcov) if not wl["filter_mode"] else 0
    alg = {
        "enum_lin": 2 * pos + rec * nk,                     # seq + qual in, one packed record per k-mer out
        "conflict_sort": 2 * rec * nk,                      # a partition reads and writes every record once
        "count_part": 64 * nk + 64 * npend + rec * nk,      # Bloom block in, written back when a bit is new (bbf.c:35-44), record in
        "tab_apply": 16 * npass + rec * nk,                 # slot read + write for every passing occurrence (htab.c:60-82)
        "ec_lookup": 32 * n_kcov + 3 * pos,                 # one 32-byte sector per lookup; seq + qual in, plane bits out
        "ec_ext": 32 * n_ext,
        "correct": 32 * n_search + 4 * pos,                 # the search's own lookups + flags / planes in, edits out
        "trim": 64 * st["n_lookups"] + pos if wl["filter_mode"] else 0,   # one 64-byte Bloom block per k-mer (bbf.c:47-63)
    }
    names = {"enum_lin": "k_enum_count + k_enum_lin", "conflict_sort": "radix partition", "count_bounds": "k_part_bounds",
             "count_part": "k_count_part", "tab_apply": "k_tab_apply_marked", "ec_lookup": "k_ec_lookup", "ec_ext": "k_ec_ext",
             "correct": "k_ec_search", "trim": "k_trim", "count_probe": "k_count_probe"}
    kern = {}
    for name, (t_ms, launches) in kt.items():
        if not launches:
            continue
        d = {"kernel": names.get(name, name), "ms": t_ms, "launches": launches, "share_of_step": t_ms / ms}
        if alg.get(name):
            d["algorithmic_bytes"] = alg[name]
            d["GBps"] = alg[name] / (t_ms / 1e3) / 1e9
            d["frac"] = d["GBps"] / peak
        kern[name] = d
    # the count phase as a whole against SURVEY section 8(d): 64 (2 - f_pass) Bloom bytes + 16 f_pass table bytes per
    # occurrence (+ 64 (2 - g) f_pass for bf_high in trim mode, g unknown: taken as 1) + 0.5 B per base of input
    count_ms = sum(kt.get(k_, (0.0, 0))[0] for k_ in ("enum_lin", "conflict_sort", "count_bounds", "count_part", "tab_apply",
                                                      "enum", "count_probe", "count_resolve", "count_replay", "bucket", "tab_rehash"))
    phase = None
    if count_ms > 0 and nk:
        b = 64 * nk + 64 * npend + (64 if wl["filter_mode"] else 16) * npass + pos // 2
        phase = {"ms": count_ms, "algorithmic_bytes": b, "GBps": b / (count_ms / 1e3) / 1e9, "frac": b / (count_ms / 1e3) / 1e9 / peak,
                 "bytes_per_occurrence": b / nk, "share_of_step": count_ms / ms}
    return kern, phase


def run(args, wl):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    n = args.reads
    config = {"workload": f"{wl['what']}: {n} x {READ_LEN} bp synthetic reads at {COVERAGE:.0f}x, k={wl['k']}, "
                          f"Bloom 2^{args.bf_shift} bits, H=4, min_cov=3 ({wl['baseline']})",
              "reads": n, "read_len": READ_LEN, "genome": genome_size(n), "k": wl["k"],
              "bf_shift": args.bf_shift, "parallelism": f"k-mers sharded by Bloom-block prefix x{world} (all-to-all per chunk), table all-gathered, correction partitioned by reads" if world > 1 else "single GPU",
              "l2": "inputs (>= 30 GB) and filter/table (>= 32 GB) exceed the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = reference_arm(args, wl, args.steps, args.warmup)
        config.update({"workload": f"{wl['what']}: bounded sample of the b200 arm's workload -- " + r["sample"],
                       "reads": r["reads"], "genome": r["genome"], "bf_shift": r["bf_shift"], "parallelism": f"{r['cores']} host threads",
                       "l2": "n/a (host)", "full_workload": f"{n} reads, Bloom 2^{args.bf_shift} bits"})
        line = {"impl": "reference", "metric": wl["metric"], "value": r["value"], "unit": "Mreads/s",
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

    cudart = C.CDLL("libcudart.so.12")
    opt = bfc_b200.make_opt(k=wl["k"], bf_shift=args.bf_shift, filter_mode=wl["filter_mode"])
    RB = READ_LEN + 1
    trim, do_correct = bool(wl["filter_mode"]), wl["correct"]

    # the command-line leg runs FIRST, while this process holds next to nothing on the GPU: device memory that one
    # process has just released is scrubbed before another process gets it, and a child started after the 100+ GB of
    # the resident run had been freed spent 1.3 s waiting for its 16 GiB filter (0.3 s otherwise)
    cli = None
    if rank == 0 and world == 1 and not args.no_cli and do_correct:
        try:
            cli = cli_e2e(L, api, wl, int(os.environ.get("BFC_BENCH_CLI_READS", min(n, 20_000_000))))
        except Exception as ex:  # reported, never required
            cli = {"value": None, "unit": "Mreads/s", "what": f"failed: {ex}"}

    if world == 1:
        data = DeviceData(L, api, n)
        n_mine, nb_mine = n, n * RB
        src_seq, src_qual = data.d_seq, data.d_qual
        w_seq = w_qual = d_aux = d_keep = d_ts = d_te = None
        if do_correct and not trim:
            w_seq, w_qual = L.bfcg_dev_alloc(max(1, nb_mine)), L.bfcg_dev_alloc(max(1, nb_mine))
            d_aux = L.bfcg_dev_alloc(max(8, 8 * n))
            assert w_seq and w_qual and d_aux
        if trim:
            d_keep, d_ts, d_te = L.bfcg_dev_alloc(max(8, n)), L.bfcg_dev_alloc(max(8, 4 * n)), L.bfcg_dev_alloc(max(8, 4 * n))
            assert d_keep and d_ts and d_te
        eng = bfc_b200.Engine(opt, timing=True)
        count_b = data.batch(data.d_seq, data.d_qual, 0, n)
        work_b = data.batch(w_seq, w_qual, 0, n) if w_seq else None
        pieces = [(0, n)]

        def step():
            eng.reset()
            if work_b is not None:
                cudart.cudaMemcpy(C.c_void_p(w_seq), C.c_void_p(src_seq), C.c_size_t(nb_mine), 3)
                cudart.cudaMemcpy(C.c_void_p(w_qual), C.c_void_p(src_qual), C.c_size_t(nb_mine), 3)
            eng.count_batch(count_b)
            if trim:
                eng.trim_batch(count_b, d_keep, d_ts, d_te)
            elif do_correct:
                eng.correct_batch(work_b, d_aux)

        def get_stats():
            return eng.stats.as_dict()

        def clear_stats():
            eng.stats = api.Stats()
    else:
        # rank r holds the r-th piece of every global chunk (bfc_b200/dist.py): local reads = its pieces back to back
        from bfc_b200.dist import CudaBackend, NativeShardedCount, ShardedCount, piece_bounds
        G = genome_size(n)
        pieces = [piece_bounds(lo, min(n, lo + args.chunk_reads), rank, world) for lo in range(0, n, args.chunk_reads)]
        n_mine = sum(p1 - p0 for p0, p1 in pieces)
        nb_mine = n_mine * RB
        d_gen = L.bfcg_dev_alloc(G)
        src_seq, src_qual = L.bfcg_dev_alloc(max(1, nb_mine)), L.bfcg_dev_alloc(max(1, nb_mine))
        w_seq = w_qual = d_aux = d_keep = d_ts = d_te = None
        if do_correct and not trim:
            w_seq, w_qual = L.bfcg_dev_alloc(max(1, nb_mine)), L.bfcg_dev_alloc(max(1, nb_mine))
            d_aux = L.bfcg_dev_alloc(max(8, 8 * n_mine))
            assert w_seq and w_qual and d_aux, L.bfcg_last_error()
        if trim:
            d_keep, d_ts, d_te = L.bfcg_dev_alloc(max(8, n_mine)), L.bfcg_dev_alloc(max(8, 4 * n_mine)), L.bfcg_dev_alloc(max(8, 4 * n_mine))
            assert d_keep and d_ts and d_te, L.bfcg_last_error()
        d_off, d_off_tmp = L.bfcg_dev_alloc(8 * (n_mine + 1)), L.bfcg_dev_alloc(8 * (max(p1 - p0 for p0, p1 in pieces) + 1))
        assert d_gen and src_seq and src_qual and d_off and d_off_tmp, L.bfcg_last_error()
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
        be = CudaBackend(opt, world, local_rank, rank=rank)
        L.bfcg_set_timing(1)
        # the exchange runs inside the library (csrc/dist.cu); BFC_DIST_PY=1 selects the round-1 path (torch all_to_all_single)
        sc = ShardedCount(be, rank, world) if os.environ.get("BFC_DIST_PY") else NativeShardedCount(be, rank, world)

        def dev_batch(seq_ptr, qual_ptr, first, count):
            b = api.Batch()
            b.n_reads, b.where, b.n_bytes = count, api.DEVICE, count * RB
            b.off = C.cast(d_off, api.u64p)   # offsets relative to the pointers below
            b.seq, b.qual = C.cast(seq_ptr + first * RB, api.u8p), C.cast(qual_ptr + first * RB, api.u8p)
            return b

        count_pieces = [dev_batch(src_seq, src_qual, piece_loc[i], p1 - p0) for i, (p0, p1) in enumerate(pieces)]
        work_b = dev_batch(w_seq, w_qual, 0, n_mine) if w_seq else None
        all_b = dev_batch(src_seq, src_qual, 0, n_mine)

        def step():
            t0 = time.perf_counter()
            be.reset()
            if work_b is not None:
                cudart.cudaMemcpy(C.c_void_p(w_seq), C.c_void_p(src_seq), C.c_size_t(nb_mine), 3)
                cudart.cudaMemcpy(C.c_void_p(w_qual), C.c_void_p(src_qual), C.c_size_t(nb_mine), 3)
            t1 = time.perf_counter()
            for b in count_pieces:
                sc.count_piece(b)
            sc.finish()
            L.bfcg_sync()
            t2 = time.perf_counter()
            if do_correct:
                sc.gather()
            L.bfcg_sync()
            t3 = time.perf_counter()
            if n_mine and trim:
                be.trim_batch(all_b, d_keep, d_ts, d_te)
            elif n_mine and do_correct:
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
    st_rank = dict(st)   # this rank's counters: what this rank's kernels (and their times) processed
    if world > 1:  # whole-job counters for the stats block
        import torch
        keys = ["n_kmers", "n_pass", "n_pending", "n_conflict", "n_lookups", "n_search_lookups", "n_redo", "n_launches"]
        t = torch.tensor([st[k_] for k_ in keys], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        launches_rank0 = st["n_launches"]
        st.update({k_: int(v) for k_, v in zip(keys, t.cpu())})
        st["n_launches"] = launches_rank0

    # ---- roofline of the dominant kernel (algorithmic bytes: DESIGN.md section 4)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    kern, count_phase = kernel_accounting(wl, st_rank, kt, n_mine, RB, args.steps, ms, peak)
    cands = [k_ for k_ in kern if "GBps" in kern[k_]]
    dom = max(cands, key=lambda k_: kern[k_]["ms"]) if cands else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if dom and os.path.exists(tpath) and args.workload == "count+correct" and world == 1:
        traffic = json.load(open(tpath)).get(dom, {}).get("dram_bytes_per_launch")
    d = kern.get(dom, {"GBps": 0.0, "launches": 1, "algorithmic_bytes": 0, "kernel": None})
    roofline = {"bound": "hbm", "kernel": d["kernel"], "achieved": d["GBps"], "peak": peak, "unit": "GB/s",
                "frac": d["GBps"] / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": d["algorithmic_bytes"] / max(1, d["launches"]),
                "count_phase": count_phase, "kernels": kern}
    if world > 1:
        roofline["note"] = "rank 0's kernels and counters"

    # ---- end to end through the C ABI with host buffers: this rank's reads start in pinned host memory, every
    # count piece and the correction copy them to the device inside the timed region, the corrected reads and
    # the per-read stats (or the trim decisions) come back to host memory
    e2e = None
    if not args.no_e2e:
        e2e_steps = min(args.steps, 2)
        pinned = []

        def pinned_u8(nbytes):  # pinned host memory (what a pipelined host would stage batches in)
            p = L.bfcg_host_alloc_pinned(max(1, nbytes))
            if not p:
                raise SystemExit("bench.py: pinned host allocation failed: " + L.bfcg_last_error().decode())
            pinned.append(p)
            return np.ctypeslib.as_array(C.cast(p, api.u8p), shape=(max(1, nbytes),))[:nbytes]
        h_seq, h_qual = pinned_u8(nb_mine), pinned_u8(nb_mine)
        L.bfcg_d2h(h_seq.ctypes.data, src_seq, nb_mine)
        L.bfcg_d2h(h_qual.ctypes.data, src_qual, nb_mine)
        h_off = np.arange(n_mine + 1, dtype=np.uint64) * np.uint64(RB)
        ws = wq = p_aux = hb_work = None
        if work_b is not None:
            ws, wq = pinned_u8(nb_mine), pinned_u8(nb_mine)
            p_aux = np.empty(2 * max(1, n_mine), dtype=np.uint32)
            hb_work = api.host_batch(ws, wq, h_off)
        h_keep = np.empty(max(1, n_mine), dtype=np.uint8)
        h_ts, h_te = np.empty(max(1, n_mine), dtype=np.int32), np.empty(max(1, n_mine), dtype=np.int32)
        hb_all = api.host_batch(h_seq, h_qual, h_off)
        hb_pieces, loc = [], 0
        for p0, p1 in pieces:
            m = p1 - p0
            hb_pieces.append(api.host_batch(h_seq[loc * RB:(loc + m) * RB], h_qual[loc * RB:(loc + m) * RB], h_off[:m + 1]))
            loc += m
        tot = 0.0
        split = {"reset_s": 0.0, "count_s": 0.0, "correct_s": 0.0}
        for it in range(1 + e2e_steps):
            if ws is not None:
                ws[:] = h_seq
                wq[:] = h_qual
            barrier()
            t0 = time.perf_counter()
            if world == 1:
                eng.reset()
                t1 = time.perf_counter()
                eng.count_batch(hb_pieces[0])
                t2 = time.perf_counter()
                if trim:
                    eng.trim_batch(hb_all, h_keep.ctypes.data, h_ts.ctypes.data, h_te.ctypes.data)
                elif do_correct:
                    eng.correct_batch(hb_work, p_aux.ctypes.data)
                if it > 0:
                    split["reset_s"] += t1 - t0; split["count_s"] += t2 - t1; split["correct_s"] += time.perf_counter() - t2
            else:
                be.reset()
                for hb in hb_pieces:
                    sc.count_piece(hb)
                sc.finish()
                if do_correct:
                    sc.gather()
                if n_mine and trim:
                    be.trim_batch(hb_all, h_keep.ctypes.data, h_ts.ctypes.data, h_te.ctypes.data)
                elif n_mine and do_correct:
                    be.correct_batch(hb_work, p_aux.ctypes.data)
            L.bfcg_sync()
            dt = max_over_ranks(time.perf_counter() - t0)
            if it > 0:
                tot += dt
        if trim:      # count: seq + qual; trim: seq + offsets in, keep / start / end out
            h2d, d2h = 2 * n * RB + n * RB + 8 * (n + world), 9 * n
        elif do_correct:  # both passes take seq + qual (+ offsets for the correction); corrected seq + qual + stats back
            h2d, d2h = 4 * n * RB + 8 * (n + world), 2 * n * RB + 8 * n
        else:
            h2d, d2h = 2 * n * RB, 0
        e2e = {"value": n * e2e_steps / tot / 1e6, "unit": "Mreads/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": e2e_steps, "split": {k_: v / e2e_steps for k_, v in split.items()} if world == 1 else None,
               "note": "whole job, all ranks: pinned host buffers -> bfcg_count_batch (N>1: bfcg_enum_records + "
               "all-to-all + bfcg_count_record_runs) / bfcg_correct_batch | bfcg_trim_batch -> host buffers; wall clock, max over ranks"}
        # full-size consistency check (outside every timed region): the results through host batches (many
        # windows, copies overlapped with the kernels) equal, byte for byte, those computed while resident in HBM
        same = None
        if nb_mine and work_b is not None:
            L.bfcg_d2h(h_seq.ctypes.data, w_seq, nb_mine)
            L.bfcg_d2h(h_qual.ctypes.data, w_qual, nb_mine)
            d_aux_h = np.empty(2 * n_mine, dtype=np.uint32)
            L.bfcg_d2h(d_aux_h.ctypes.data, d_aux, 8 * n_mine)
            same = bool(np.array_equal(h_seq, ws) and np.array_equal(h_qual, wq) and np.array_equal(d_aux_h, p_aux[:2 * n_mine]))
        elif nb_mine and trim:
            k2, s2, e2 = np.empty(n_mine, dtype=np.uint8), np.empty(n_mine, dtype=np.int32), np.empty(n_mine, dtype=np.int32)
            L.bfcg_d2h(k2.ctypes.data, d_keep, n_mine); L.bfcg_d2h(s2.ctypes.data, d_ts, 4 * n_mine); L.bfcg_d2h(e2.ctypes.data, d_te, 4 * n_mine)
            same = bool(np.array_equal(k2, h_keep[:n_mine]) and np.array_equal(s2, h_ts[:n_mine]) and np.array_equal(e2, h_te[:n_mine]))
            e2e["kept_frac"] = float(k2.mean())
        if same is not None or world > 1:
            e2e["equals_resident_result"] = bool(max_over_ranks(0.0 if same in (True, None) else 1.0) == 0.0)
        for p_ in pinned:
            L.bfcg_host_free_pinned(p_)

    exchange = None
    if world > 1 and hasattr(sc, "stats"):
        exchange = sc.stats()
        exchange["bytes_per_record"] = 9 if wl["k"] <= 35 else 10 if wl["k"] <= 39 else 12 if wl["k"] <= 47 else 16
        exchange["note"] = "rank 0, whole run (warm-up, timed steps and the end-to-end steps): grouped ncclSend/ncclRecv on the library's exchange stream"
    n_distinct = None
    if world == 1:
        if eng.ch:
            n_distinct = eng.n_distinct()
        eng.close()
        for p_ in (w_seq, w_qual, d_aux, d_keep, d_ts, d_te):
            if p_:
                L.bfcg_dev_free(p_)
        data.free()
    else:
        be.close()
        NativeShardedCount.finalize(L)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = reference_arm(args, wl, 1, 0)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:  # the baseline is reported, never required
            cpu = {"value": None, "unit": "Mreads/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}

    if rank == 0:
        line = {"metric": wl["metric"], "value": value, "unit": "Mreads/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
                "clocks": clk, "e2e": e2e, "e2e_cli": cli, "exchange": exchange, "gpu_launches": int(st["n_launches"]), "roofline": roofline,
                "cpu_baseline": cpu,
                "stats": {"kmers_per_step": st["n_kmers"] // args.steps, "f_pass": st["n_pass"] / max(1, st["n_kmers"]),
                          "pending_frac": st["n_pending"] / max(1, st["n_kmers"]),
                          "conflict_frac": st["n_conflict"] / max(1, st["n_kmers"]),
                          "lookups_per_read": st["n_lookups"] / max(1, n * args.steps),
                          "search_lookups_per_read": st["n_search_lookups"] / max(1, n * args.steps), "redo": st["n_redo"],
                          "distinct_kmers_in_table": n_distinct}}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="count+correct", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=int(os.environ.get("BFC_BENCH_READS", 100_000_000)))
    ap.add_argument("--bf-shift", type=int, default=37)
    ap.add_argument("--chunk-reads", type=int, default=16_000_000, help="N > 1: reads per global chunk (one all-to-all each)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cli", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    try:
        run(args, WORKLOADS[args.workload])
    except BaseException as ex:  # torchrun's summary swallows a rank's traceback: leave a rank-tagged last line and a file
        if isinstance(ex, SystemExit) and ex.code in (0, None):
            raise
        msg = f"[rank {rank}] {type(ex).__name__}: {ex}"
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", f"bench_error_rank{rank}.txt"), "w") as fp:
                fp.write(msg + "\n" + traceback.format_exc())
        except OSError:
            pass
        traceback.print_exc()
        sys.stderr.write(msg + "\n")
        sys.stderr.flush()
        os._exit(1)   # do not hang in a collective's destructor while the other ranks wait


if __name__ == "__main__":
    main()
