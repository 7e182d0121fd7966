"""CPU (gloo, world_size 2 and 4): the sharded count protocol of bfc_b200/dist.py -- owner bucketing,
all-to-all of k-mer records, ordered cascade on the owner, all-gather of the table / bf_high -- driven
with the oracle as compute backend, against the oracle's single-process run on the same reads.
The shards must reassemble into exactly the single-process filter and table (the `-t1` result)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402

from bfc_b200 import synth  # noqa: E402
from bfc_b200.dist import ShardedCount, owner_bits, piece_bounds  # noqa: E402


class OracleBackend:
    """ShardedCount backend on the CPU oracle.  Every rank keeps a FULL-size filter of which it only ever
    touches the blocks it owns; the test checks the owned ranges."""
    device = torch.device("cpu")

    def __init__(self, opt, world):
        self.L, self.opt, self.world = orc.lib(), opt, world
        self.filter_mode = bool(opt.filter_mode)
        self.run = orc.OracleRun(opt)
        self.full_table = None
        self.full_bf_high = None

    def empty(self, n, dtype):
        return torch.empty(int(n), dtype=dtype)

    def enum_records(self, piece, world):
        seq, qual, off = piece
        b = orc.Batch(len(off) - 1, orc.as_u64p(off), orc.as_u8p(seq), orc.as_u8p(qual) if qual is not None else None)
        n = int(self.L.orc_enum_records(C.byref(self.opt), C.byref(b), None, None))
        y0, y1 = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
        if n:
            self.L.orc_enum_records(C.byref(self.opt), C.byref(b), orc.as_u64p(y0), orc.as_u64p(y1))
        k, x = self.opt.k, self.opt.bf_shift - 9
        hashes = np.array([self.L.orc_hash_from_y(k, int(a) & ~(1 << 63), int(c)) for a, c in zip(y0, y1)], dtype=np.uint64)
        blk = hashes & np.uint64((1 << x) - 1)
        owner = (blk >> np.uint64(x - owner_bits(world))).astype(np.int64) if world > 1 else np.zeros(n, dtype=np.int64)
        order = np.argsort(owner, kind="stable")
        counts = [int((owner == o).sum()) for o in range(world)]
        return torch.from_numpy(y0[order].view(np.int64)), torch.from_numpy(y1[order].view(np.int64)), counts

    def count_records(self, y0, y1, n, world):
        a, b = y0.numpy().view(np.uint64), y1.numpy().view(np.uint64)
        r = self.run
        self.L.orc_count_records(C.byref(self.opt), r.bf, r.bf_high, r.ch, n, orc.as_u64p(a), orc.as_u64p(b), orc.as_u64p(r.stats))

    def export_table(self):
        sub, key = self.run.table()
        return torch.from_numpy(sub.view(np.int32)), torch.from_numpy(key.view(np.int64))

    def import_table(self, parts):
        sub = np.concatenate([s.numpy().view(np.uint32) for s, _ in parts])
        key = np.concatenate([k.numpy().view(np.uint64) for _, k in parts])
        order = np.lexsort((key, sub))
        self.full_table = (sub[order], key[order])

    def bf_high_shard(self):
        full = self.run.bloom_bytes(high=True)
        n = len(full) // self.world
        return torch.from_numpy(full[self.rank * n:(self.rank + 1) * n].copy())

    def set_bf_high_full(self, full):
        self.full_bf_high = full.numpy().copy()


def make_reads(seed, n_reads=3000):
    genome = synth.make_genome(20000, seed, 0.2)
    seq, qual = synth.make_reads(genome, n_reads, 100, seed)
    return seq, qual


def _worker(rank, world, port, k, b, trim, chunk, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        opt = orc.make_opt(k=k, bf_shift=b, filter_mode=1 if trim else 0)
        seq, qual = make_reads(seed=k + b)
        n = len(seq)
        be = OracleBackend(opt, world)
        be.rank = rank
        sc = ShardedCount(be, rank, world)
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            p0, p1 = piece_bounds(lo, hi, rank, world)
            sc.count_piece(synth.concat_batch(seq[p0:p1], qual[p0:p1]))
        sc.gather()
        res = {"bloom": be.run.bloom_bytes(), "stats": be.run.stats.copy()}
        if trim:
            res["bf_high"] = be.full_bf_high
        else:
            res["sub"], res["key"] = be.full_table
        np.savez(os.path.join(out, f"rank{rank}.npz"), **res)
        be.run.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,k,b,trim,chunk", [(2, 21, 18, False, 1000), (2, 31, 20, True, 3000), (4, 33, 16, False, 700)])
def test_sharded_count_matches_single_process(tmp_path, world, k, b, trim, chunk):
    port = 29500 + (os.getpid() * 7 + world * 13 + k) % 2000
    mp.spawn(_worker, args=(world, port, k, b, trim, chunk, str(tmp_path)), nprocs=world, join=True)
    # the single-process oracle on the whole input, in read order
    opt = orc.make_opt(k=k, bf_shift=b, filter_mode=1 if trim else 0)
    seq, qual = make_reads(seed=k + b)
    ref = orc.OracleRun(opt)
    ref.count(*synth.concat_batch(seq, qual))
    bloom = ref.bloom_bytes()
    shard = len(bloom) // world
    tot = np.zeros(2, dtype=np.uint64)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        # a rank touched only the blocks it owns, and they hold exactly the single-process bytes
        assert np.array_equal(got["bloom"][r * shard:(r + 1) * shard], bloom[r * shard:(r + 1) * shard])
        mask = np.ones(len(bloom), dtype=bool)
        mask[r * shard:(r + 1) * shard] = False
        assert not got["bloom"][mask].any()
        tot += got["stats"]
        if trim:
            assert np.array_equal(got["bf_high"], ref.bloom_bytes(high=True))
        else:
            sub, key = ref.table()
            assert np.array_equal(got["sub"], sub) and np.array_equal(got["key"], key)
    assert np.array_equal(tot, ref.stats)
    ref.close()


def test_piece_bounds_cover_the_chunk_in_rank_order():
    for world in (1, 2, 4, 8):
        for lo, hi in ((0, 10), (5, 6), (7, 7), (100, 1037)):
            b = [piece_bounds(lo, hi, r, world) for r in range(world)]
            assert b[0][0] == lo and b[-1][1] == hi
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        owner_bits(3)
