"""Synthetic read generator (SURVEY.md §8(d) / Appendix C recipe).

genome  = i.i.d. uniform ACGT of length G from numpy PCG64(seed)
reads   = N reads of length L, start uniform, strand uniform, per-base substitution
          error p (uniform over the 3 other bases), quality: error bases Q in U[2,19],
          others Q in U[25,40] (Phred+33), N-rate `n_rate`
records = "@r<i>\n<seq>\n+\n<qual>\n"

Optionally a fraction of the genome is made of diverged repeat copies
(`repeat_frac`), which gives the correction search real branching.

Pure host code (numpy).  Used by the tests, by tools/make_golden.py and by the
reference arm of bench.py.  The device-side generator used for the HBM-resident
bench lives in csrc/ (bfcg_synth_*), it follows the same distribution but not the
same PRNG stream.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def make_genome(G: int, seed: int, repeat_frac: float = 0.0, repeat_div: float = 0.015) -> np.ndarray:
    """Return the genome as base codes 0..3 (uint8)."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, size=G, dtype=np.uint8)
    if repeat_frac > 0:
        # copy random 2 kb segments elsewhere with `repeat_div` divergence
        seg = 2000 if G >= 20000 else max(50, G // 20)
        n_seg = int(G * repeat_frac / seg)
        for _ in range(n_seg):
            a = int(rng.integers(0, G - seg))
            b = int(rng.integers(0, G - seg))
            s = g[a:a + seg].copy()
            m = rng.random(seg) < repeat_div
            s[m] = (s[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
            g[b:b + seg] = s
    return g


def make_reads(genome: np.ndarray, N: int, L: int, seed: int, err: float = 0.01, n_rate: float = 2e-4):
    """Return (seq, qual): two uint8 arrays of shape (N, L) holding ASCII."""
    rng = np.random.default_rng(seed + 0x5eed)
    G = genome.shape[0]
    start = rng.integers(0, G - L + 1, size=N)
    strand = rng.integers(0, 2, size=N, dtype=np.uint8)
    idx = start[:, None] + np.arange(L)[None, :]
    codes = genome[idx]
    rc = (3 - codes[:, ::-1])
    codes = np.where(strand[:, None] == 1, rc, codes).astype(np.uint8)
    emask = rng.random((N, L)) < err
    sub = rng.integers(1, 4, size=(N, L), dtype=np.uint8)
    codes = np.where(emask, (codes + sub) & 3, codes).astype(np.uint8)
    q_ok = rng.integers(25, 41, size=(N, L), dtype=np.uint8)
    q_err = rng.integers(2, 20, size=(N, L), dtype=np.uint8)
    qual = (np.where(emask, q_err, q_ok) + 33).astype(np.uint8)
    seq = _ACGT[codes]
    nmask = rng.random((N, L)) < n_rate
    seq = np.where(nmask, np.uint8(ord("N")), seq).astype(np.uint8)
    return seq, qual


def fastq_bytes(seq: np.ndarray, qual: np.ndarray | None, first_index: int = 0, prefix: str = "r") -> bytes:
    """Serialise to FASTQ (or FASTA when qual is None)."""
    N, L = seq.shape
    out = []
    hdr = b"@" if qual is not None else b">"
    for i in range(N):
        out.append(hdr + f"{prefix}{first_index + i}\n".encode())
        out.append(seq[i].tobytes())
        if qual is not None:
            out.append(b"\n+\n")
            out.append(qual[i].tobytes())
        out.append(b"\n")
    return b"".join(out)


def write_fastq(path: str, G: int, N: int, L: int, seed: int, err: float = 0.01, n_rate: float = 2e-4,
                repeat_frac: float = 0.0, chunk: int = 200_000) -> None:
    """Write a synthetic FASTQ file (streamed in chunks so that large N stays cheap)."""
    genome = make_genome(G, seed, repeat_frac)
    with open(path, "wb") as fp:
        done = 0
        while done < N:
            n = min(chunk, N - done)
            seq, qual = make_reads(genome, n, L, seed + 7919 * (done // chunk), err, n_rate)
            fp.write(fastq_bytes(seq, qual, first_index=done))
            done += n


def concat_batch(seq: np.ndarray, qual: np.ndarray | None):
    """(N, L) ASCII arrays -> the flat host batch the C-ABI takes:
    `seq`/`qual` byte streams where every read is followed by one 0 byte, and
    `offsets` (N+1 uint64) giving each read's start in the stream."""
    N, L = seq.shape
    s = np.zeros((N, L + 1), dtype=np.uint8)
    s[:, :L] = seq
    q = None
    if qual is not None:
        q = np.zeros((N, L + 1), dtype=np.uint8)
        q[:, :L] = qual
        q = q.reshape(-1)
    offsets = (np.arange(N + 1, dtype=np.uint64) * np.uint64(L + 1))
    return s.reshape(-1), q, offsets


# ----------------------------------------------------------------------------- counter-based twin of csrc/synth.cu

def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def cb_genome(G: int, seed: int, lo: int = 0, hi: int | None = None) -> np.ndarray:
    """genome[lo:hi] as codes 0..3 -- same values as bfcg_synth_genome."""
    hi = G if hi is None else hi
    with np.errstate(over="ignore"):
        i = np.arange(lo, hi, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x632BE59BD9B4E019)
    return (_splitmix64(i) >> np.uint64(62)).astype(np.uint8)


def cb_reads(G: int, seed: int, first_read: int, n_reads: int, L: int, err: float = 0.01, n_rate: float = 2e-4,
             genome_seed: int | None = None):
    """(seq, qual) ASCII arrays of shape (n_reads, L) -- same values as bfcg_synth_reads.
    The genome is evaluated lazily at the positions the reads cover, so a 3 Gb genome needs no memory."""
    gseed = seed if genome_seed is None else genome_seed
    with np.errstate(over="ignore"):
        r = (np.arange(n_reads, dtype=np.uint64) + np.uint64(first_read)) * np.uint64(0xD6E8FEB86659FD93)
        h = _splitmix64(np.uint64(seed) ^ r)
        start = h % np.uint64(G - L + 1)
        strand = (_splitmix64(h) & np.uint64(1)).astype(bool)
        j = np.arange(L, dtype=np.uint64)
        pos = np.where(strand[:, None], start[:, None] + np.uint64(L - 1) - j[None, :], start[:, None] + j[None, :])
        g = (_splitmix64(pos + np.uint64(gseed) * np.uint64(0x632BE59BD9B4E019)) >> np.uint64(62)).astype(np.int64)
        c = np.where(strand[:, None], 3 - g, g)
        u = _splitmix64(h[:, None] ^ ((j[None, :] + np.uint64(1)) * np.uint64(0xA24BAED4963EE407)))
    is_err = (u & np.uint64(0xFFFF)).astype(np.int64) < int(err * 65536.0 + 0.5)
    c = np.where(is_err, (c + 1 + ((u >> np.uint64(16)) & np.uint64(0xFF)).astype(np.int64) % 3) & 3, c)
    q = np.where(is_err, 2 + ((u >> np.uint64(28)) & np.uint64(0xFF)).astype(np.int64) % 18,
                 25 + ((u >> np.uint64(24)) & np.uint64(15)).astype(np.int64))
    is_n = ((u >> np.uint64(40)) & np.uint64(0xFFFFF)).astype(np.int64) < int(n_rate * 1048576.0 + 0.5)
    seq = np.where(is_n, np.uint8(ord("N")), _ACGT[c]).astype(np.uint8)
    qual = (q + 33).astype(np.uint8)
    return seq, qual
