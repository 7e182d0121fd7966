"""Multi-GPU count phase: one process per GPU, k-mers sharded by Bloom-block prefix.

Every ordering constraint of the count phase is local to one 64-byte Bloom block
(reference bbf.c:25-45, count.c:54-70), so rank r of N owns the blocks whose index has
the top log2(N) bits equal to r -- a contiguous 1/N of the first filter (and of bf_high
in trim mode) -- and the table entries of exactly those k-mers.

The reads are consumed in GLOBAL chunks; rank r holds the r-th piece of every chunk.
Per chunk:
    1. enumerate the k-mers of the local piece and bucket them by owner   (bfcg_enum_records)
    2. all-to-all of the 16-byte records over NVLink                      (torch.distributed / NCCL)
       -- the receiver gets the pieces concatenated in rank order = the global read order
    3. the Bloom -> table cascade over the received records, in order     (bfcg_count_records)
so the result is that of the reference's `-t1` run on the whole input, bit for bit.
After the last chunk the table shards (or the bf_high shards) are all-gathered: every
rank ends with the complete table and corrects its own reads with no communication.

The protocol is written against a small backend interface so that the same driver runs
on the CUDA library (CudaBackend, below) and, in the CPU tests, on the oracle with the
gloo backend (tests/test_dist_cpu.py).  torch is plumbing here: device buffers for the
exchange and the collectives; every byte of compute is in libbfc_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def owner_bits(world: int) -> int:
    b = world.bit_length() - 1
    if world < 1 or (1 << b) != world or world > 8:
        raise ValueError("the number of ranks must be 1, 2, 4 or 8")
    return b


def piece_bounds(chunk_lo: int, chunk_hi: int, rank: int, world: int):
    """Reads [lo, hi) of the global chunk [chunk_lo, chunk_hi) that rank `rank` holds: contiguous, in rank order."""
    n = chunk_hi - chunk_lo
    return chunk_lo + n * rank // world, chunk_lo + n * (rank + 1) // world


class ShardedCount:
    """The count phase of one rank.  `backend` supplies the compute and the buffers."""

    def __init__(self, backend, rank: int, world: int, group=None):
        owner_bits(world)
        self.be, self.rank, self.world, self.group = backend, rank, world, group
        self.sent_records = 0
        self.recv_records = 0
        import os
        self.prof = {} if os.environ.get("BFC_DIST_PROFILE") else None

    # -- collectives (world == 1 short-circuits so a single process needs no process group)
    def _all_gather_counts(self, counts):
        t = torch.tensor(counts, dtype=torch.int64, device=self.be.device)
        if self.world == 1:
            return t.view(1, -1).cpu()
        out = torch.empty(self.world * self.world, dtype=torch.int64, device=self.be.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return out.view(self.world, self.world).cpu()

    def _all_to_all(self, src, in_splits, out_splits):
        out = self.be.empty(int(sum(out_splits)), torch.int64)
        if self.world == 1:
            out.copy_(src[:out.numel()])
            return out
        dist.all_to_all_single(out, src[:int(sum(in_splits))], list(out_splits), list(in_splits), group=self.group)
        return out

    def _tick(self, name, t0):
        """Phase boundary: the library works on its own stream and torch / NCCL on theirs, so every phase ends with a
        device synchronisation (the count path is host-synchronous anyway).  BFC_DIST_PROFILE=1 also attributes the
        wall clock to the phases."""
        import time
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        if self.prof is None:
            return 0.0
        t = time.perf_counter()
        if t0:
            self.prof[name] = self.prof.get(name, 0.0) + (t - t0)
        return t

    def count_piece(self, piece):
        """Count this rank's piece of the current global chunk (every rank must call this once per chunk)."""
        t = self._tick("other", 0.0)
        y0, y1, counts = self.be.enum_records(piece, self.world)
        t = self._tick("enum+bucket", t)
        m = self._all_gather_counts(counts)              # m[src][dst]
        in_splits = [int(v) for v in m[self.rank]]
        out_splits = [int(m[src][self.rank]) for src in range(self.world)]
        t = self._tick("counts", t)
        r0 = self._all_to_all(y0, in_splits, out_splits)
        r1 = self._all_to_all(y1, in_splits, out_splits)
        t = self._tick("all_to_all", t)
        self.sent_records += sum(in_splits) - in_splits[self.rank]
        self.recv_records += sum(out_splits)
        if hasattr(self.be, "count_record_runs"):     # pieces as delivered: the sender may have partitioned them
            self.be.count_record_runs(r0, r1, out_splits, self.world)
        else:
            self.be.count_records(r0, r1, int(sum(out_splits)), self.world)
        self._tick("count_records", t)

    def finish(self):
        """(Nothing is left in flight: every count_piece completes its own chunk.)"""

    def gather(self):
        """Replicate the result on every rank: the complete table (normal mode) or bf_high (trim mode)."""
        t = self._tick("other", 0.0)
        if self.be.filter_mode:
            shard = self.be.bf_high_shard()
            if self.world == 1:
                full = shard
            else:
                full = self.be.empty(shard.numel() * self.world, torch.uint8)
                dist.all_gather_into_tensor(full, shard, group=self.group)
            self.be.set_bf_high_full(full)
            return
        sub, key = self.be.export_table()
        t = self._tick("export", t)
        n = int(sub.numel())
        if self.world == 1:
            self.be.import_table([(sub, key)])
            return
        sizes = torch.empty(self.world, dtype=torch.int64, device=self.be.device)
        dist.all_gather_into_tensor(sizes, torch.tensor([n], dtype=torch.int64, device=self.be.device), group=self.group)
        sizes = [int(v) for v in sizes.cpu()]
        cap = max(max(sizes), 1)
        psub, pkey = self.be.empty(cap, torch.int32), self.be.empty(cap, torch.int64)
        psub[:n].copy_(sub)
        pkey[:n].copy_(key)
        gsub, gkey = self.be.empty(cap * self.world, torch.int32), self.be.empty(cap * self.world, torch.int64)
        dist.all_gather_into_tensor(gsub, psub, group=self.group)
        dist.all_gather_into_tensor(gkey, pkey, group=self.group)
        t = self._tick("all_gather", t)
        self.be.import_table([(gsub[r * cap:r * cap + sizes[r]], gkey[r * cap:r * cap + sizes[r]]) for r in range(self.world)])
        self._tick("import", t)


class CudaBackend:
    """The C ABI of libbfc_b200.so (include/bfc_b200.h) behind the ShardedCount backend interface."""

    def __init__(self, opt, world: int, device_index: int = 0, rank: int | None = None):
        from . import api
        self.api, self.L, self.opt, self.world = api, api.lib(), opt, world
        self.filter_mode = bool(opt.filter_mode)
        self.device = torch.device("cuda", device_index)
        L = self.L
        self.stats = api.Stats()
        self.bf = L.bfcg_bf_init_shard(opt.bf_shift, opt.n_hashes, world)
        self.bf_high = L.bfcg_bf_init_shard(opt.bf_shift, opt.n_hashes, world) if self.filter_mode else None
        self.ch = None if self.filter_mode else L.bfc_ch_init(opt.k, opt.l_pre)
        if not self.bf or (self.filter_mode and not self.bf_high) or (not self.filter_mode and not self.ch):
            raise api.BfcError("allocation failed: " + L.bfcg_last_error().decode())
        if self.ch and world > 1 and rank is not None:  # only 1/world of the sub-table regions can receive keys
            self._check(L.bfcg_ch_set_shard(self.ch, world, rank), "bfcg_ch_set_shard")
        self._native_full_bf = None  # complete bf_high owned by the library (NativeShardedCount.gather)
        self.full_ch = None          # complete table after gather()
        self.full_bf_high = None     # complete bf_high (torch tensor keeps the memory) after gather()
        self._bf_high_view = None
        self._buf = {}

    def empty(self, n, dtype):
        return torch.empty(max(int(n), 0), dtype=dtype, device=self.device)

    def _scratch(self, name, n):
        t = self._buf.get(name)
        if t is None or t.numel() < n:
            t = self._buf[name] = torch.empty(int(n), dtype=torch.int64, device=self.device)
        return t

    def _check(self, rc, what):
        if rc != 0:
            raise self.api.BfcError(f"{what} failed ({rc}): {self.L.bfcg_last_error().decode()}")

    def enum_records(self, batch, world):
        """batch: api.Batch (host or device pointers).  Returns (y0, y1, counts) bucketed by owner."""
        n = int(batch.n_bytes)
        y0, y1 = self._scratch("y0", n), self._scratch("y1", n)
        counts = (C.c_uint64 * 8)()
        torch.cuda.current_stream(self.device).synchronize()
        self._check(self.L.bfcg_enum_records(C.byref(self.opt), C.byref(batch), world, C.c_void_p(y0.data_ptr()),
                                             C.c_void_p(y1.data_ptr()), counts), "bfcg_enum_records")
        return y0, y1, [int(counts[i]) for i in range(world)]

    def count_records(self, y0, y1, n, world):
        torch.cuda.current_stream(self.device).synchronize()
        self._check(self.L.bfcg_count_records(C.byref(self.opt), self.bf, self.bf_high, self.ch, n,
                                              C.c_void_p(y0.data_ptr()), C.c_void_p(y1.data_ptr()), world,
                                              C.byref(self.stats)), "bfcg_count_records")

    def count_record_runs(self, y0, y1, run_counts, world):
        torch.cuda.current_stream(self.device).synchronize()
        rc = (C.c_uint64 * len(run_counts))(*[int(v) for v in run_counts])
        self._check(self.L.bfcg_count_record_runs(C.byref(self.opt), self.bf, self.bf_high, self.ch, len(run_counts), rc,
                                                  C.c_void_p(y0.data_ptr()), C.c_void_p(y1.data_ptr()), world,
                                                  C.byref(self.stats)), "bfcg_count_record_runs")

    def shard_bytes(self) -> int:
        return (1 << (self.opt.bf_shift - 3)) // self.world

    def _shard_tensor(self, bf):
        t = torch.empty(self.shard_bytes(), dtype=torch.uint8, device=self.device)
        cudart = C.CDLL("libcudart.so.12")
        torch.cuda.current_stream(self.device).synchronize()
        rc = cudart.cudaMemcpy(C.c_void_p(t.data_ptr()), C.c_void_p(bf.contents.b), C.c_size_t(t.numel()), 3)
        if rc != 0:
            raise self.api.BfcError(f"cudaMemcpy failed: {rc}")
        return t

    def bf_shard(self):
        return self._shard_tensor(self.bf)

    def bf_high_shard(self):
        return self._shard_tensor(self.bf_high)

    def set_bf_high_full(self, full):
        torch.cuda.current_stream(self.device).synchronize()
        self.full_bf_high = full
        self._bf_high_view = self.api.BF(self.opt.bf_shift, self.opt.n_hashes, full.data_ptr())

    def bf_high_full(self):
        """bfc_bf_t* over the gathered filter (for bfcg_trim_batch)."""
        return C.pointer(self._bf_high_view)

    def export_table(self):
        n = int(self.L.bfcg_ch_export_device(self.ch, None, None))
        sub, key = self.empty(n, torch.int32), self.empty(n, torch.int64)
        if n:
            got = int(self.L.bfcg_ch_export_device(self.ch, C.c_void_p(sub.data_ptr()), C.c_void_p(key.data_ptr())))
            if got != n:
                raise self.api.BfcError("bfcg_ch_export_device failed: " + self.L.bfcg_last_error().decode())
        return sub, key

    def import_table(self, parts):
        L = self.L
        total = sum(int(s.numel()) for s, _ in parts)
        if self.world == 1:
            self.full_ch = self.ch   # a single rank already holds everything
            return
        full = self.full_ch if self.full_ch and self.full_ch != self.ch else L.bfc_ch_init(self.opt.k, self.opt.l_pre)
        if not full:
            raise self.api.BfcError("bfc_ch_init failed: " + L.bfcg_last_error().decode())
        self._check(L.bfcg_ch_reserve(full, total), "bfcg_ch_reserve")
        torch.cuda.current_stream(self.device).synchronize()
        for sub, key in parts:
            if sub.numel():
                self._check(L.bfcg_ch_import_device(full, int(sub.numel()), C.c_void_p(sub.data_ptr()),
                                                    C.c_void_p(key.data_ptr())), "bfcg_ch_import_device")
        self.full_ch = full

    def reset(self):
        """Empty shards and tables (between bench steps); the allocations are kept, as Engine.reset() keeps them."""
        L = self.L
        cudart = C.CDLL("libcudart.so.12")
        torch.cuda.synchronize(self.device)
        for bf in (self.bf, self.bf_high):
            if bf:
                rc = cudart.cudaMemset(C.c_void_p(bf.contents.b), 0, C.c_size_t(self.shard_bytes()))
                if rc != 0:
                    raise self.api.BfcError(f"cudaMemset failed: {rc}")
        for ch in {self.ch, self.full_ch}:
            if ch:
                self._check(L.bfcg_ch_clear(ch), "bfcg_ch_clear")
        self.full_bf_high = self._bf_high_view = None
        if self._native_full_bf:
            self._check(L.bfcg_bf_clear(self._native_full_bf), "bfcg_bf_clear")
        if self.world > 1 and self.full_ch == self.ch:
            self.full_ch = None

    def mode(self) -> int:
        """bfc_ch_hist of the gathered table (reference correct.c:633)."""
        cnt = (C.c_uint64 * 256)()
        high = (C.c_uint64 * 64)()
        return int(self.L.bfc_ch_hist(self.full_ch, cnt, high))

    def correct_batch(self, batch, aux_ptr, mode=None):
        """bfc_ec1 on every read of `batch` against the gathered table."""
        torch.cuda.current_stream(self.device).synchronize()
        self._check(self.L.bfcg_correct_batch(C.byref(self.opt), self.full_ch, self.mode() if mode is None else mode,
                                              C.byref(batch), aux_ptr, C.byref(self.stats)), "bfcg_correct_batch")

    def trim_batch(self, batch, keep_ptr, ts_ptr, te_ptr):
        torch.cuda.current_stream(self.device).synchronize()
        self._check(self.L.bfcg_trim_batch(C.byref(self.opt), self.bf_high_full(), C.byref(batch), keep_ptr, ts_ptr, te_ptr,
                                           C.byref(self.stats)), "bfcg_trim_batch")

    def release_first_filter(self):
        """The first filter never outlives the count phase (reference count.c:155)."""
        if self.bf:
            self.L.bfc_bf_destroy(self.bf)
            self.bf = None

    def close(self):
        L = self.L
        self.release_first_filter()
        if self.bf_high:
            L.bfc_bf_destroy(self.bf_high)
            self.bf_high = None
        if self._native_full_bf:
            L.bfc_bf_destroy(self._native_full_bf)
            self._native_full_bf = None
        if self.full_ch and self.full_ch != self.ch:
            L.bfc_ch_destroy(self.full_ch)
        if self.ch:
            L.bfc_ch_destroy(self.ch)
        self.ch = self.full_ch = None
        self._buf.clear()


class NativeShardedCount:
    """The count phase of one rank with the exchange INSIDE the library (csrc/dist.cu: grouped ncclSend / ncclRecv on
    the library's own stream, overlapped with the cascade of the previous chunk and the enumeration of the next piece).
    Same interface as ShardedCount; torch.distributed only carries the 128-byte NCCL id to the ranks."""

    _inited = None

    def __init__(self, backend: "CudaBackend", rank: int, world: int, group=None):
        owner_bits(world)
        self.be, self.rank, self.world = backend, rank, world
        self.prof = None
        L = backend.L
        if NativeShardedCount._inited is None:
            idb = (C.c_uint8 * 128)()
            if rank == 0:
                backend._check(L.bfcg_dist_unique_id(idb), "bfcg_dist_unique_id")
            if world > 1:
                t = torch.tensor(list(idb), dtype=torch.uint8, device=backend.device)
                dist.broadcast(t, 0, group=group)
                idb = (C.c_uint8 * 128)(*[int(v) for v in t.cpu()])
            backend._check(L.bfcg_dist_init(rank, world, idb), "bfcg_dist_init")
            NativeShardedCount._inited = (rank, world)
        elif NativeShardedCount._inited != (rank, world):
            raise ValueError("the library's communicator was created for another (rank, world)")

    def count_piece(self, piece):
        be = self.be
        be._check(be.L.bfcg_dist_count_piece(C.byref(be.opt), be.bf, be.bf_high, be.ch, C.byref(piece), C.byref(be.stats)),
                  "bfcg_dist_count_piece")

    def finish(self):
        be = self.be
        be._check(be.L.bfcg_dist_count_finish(C.byref(be.opt), be.bf, be.bf_high, be.ch, C.byref(be.stats)), "bfcg_dist_count_finish")

    def gather(self):
        """Drain the pipeline, then replicate the complete table (normal mode) or bf_high (trim mode) on every rank."""
        be, L = self.be, self.be.L
        self.finish()
        if be.filter_mode:
            if be._native_full_bf is None:
                be._native_full_bf = L.bfc_bf_init(be.opt.bf_shift, be.opt.n_hashes)
                if not be._native_full_bf:
                    raise be.api.BfcError("bfc_bf_init failed: " + L.bfcg_last_error().decode())
            be._check(L.bfcg_dist_gather_filter(be.bf_high, be._native_full_bf), "bfcg_dist_gather_filter")
            be._bf_high_view = be._native_full_bf.contents
            return
        if self.world == 1:
            be.full_ch = be.ch
            return
        full = be.full_ch if be.full_ch and be.full_ch != be.ch else L.bfc_ch_init(be.opt.k, be.opt.l_pre)
        if not full:
            raise be.api.BfcError("bfc_ch_init failed: " + L.bfcg_last_error().decode())
        be.full_ch = full
        be._check(L.bfcg_dist_gather_table(be.ch, full), "bfcg_dist_gather_table")

    def stats(self):
        sent, recv, ms = C.c_uint64(), C.c_uint64(), C.c_double()
        self.be.L.bfcg_dist_stats(C.byref(sent), C.byref(recv), C.byref(ms))
        return {"sent_records": sent.value, "received_records": recv.value, "exchange_ms": ms.value}

    @staticmethod
    def finalize(L):
        if NativeShardedCount._inited is not None:
            L.bfcg_dist_finalize()
            NativeShardedCount._inited = None
