"""GPU, >= 2 devices (skipped otherwise): the sharded count over NCCL on real ranks, one process per GPU,
then correction partitioned by reads -- against the single-process oracle."""
import os
import sys

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, k, b, trim, chunk, out, native=False):
    import ctypes as C
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bfc_b200
    from bfc_b200 import api, synth
    from bfc_b200.dist import CudaBackend, NativeShardedCount, ShardedCount, piece_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    assert api.lib().bfcg_set_device(rank) == 0
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        genome = synth.make_genome(80000, k + b, 0.2)
        seq, qual = synth.make_reads(genome, 24000, 120, k + b)
        n = len(seq)
        opt = bfc_b200.make_opt(k=k, bf_shift=b, filter_mode=1 if trim else 0)
        be = CudaBackend(opt, world, rank, rank=rank)
        sc = NativeShardedCount(be, rank, world) if native else ShardedCount(be, rank, world)
        mine = []
        for lo in range(0, n, chunk):
            p0, p1 = piece_bounds(lo, min(n, lo + chunk), rank, world)
            s, q, off = synth.concat_batch(seq[p0:p1], qual[p0:p1])
            sc.count_piece(api.host_batch(s, q, off))
            mine.append((p0, p1))
        sc.finish()
        res = {"bloom": be.bf_shard().cpu().numpy(), "n_kmers": int(be.stats.n_kmers), "n_pass": int(be.stats.n_pass)}
        sc.gather()
        for i, (p0, p1) in enumerate(mine):
            s, q, off = synth.concat_batch(seq[p0:p1], qual[p0:p1])
            if trim:
                m = p1 - p0
                keep, ts, te = np.zeros(m, dtype=np.uint8), np.zeros(m, dtype=np.int32), np.zeros(m, dtype=np.int32)
                be.trim_batch(api.host_batch(s, None, off), keep.ctypes.data, ts.ctypes.data, te.ctypes.data)
                res[f"keep{i}"], res[f"ts{i}"], res[f"te{i}"] = keep, ts, te
            else:
                aux = np.zeros(2 * (p1 - p0), dtype=np.uint32)
                be.correct_batch(api.host_batch(s, q, off), aux.ctypes.data)
                res[f"seq{i}"], res[f"qual{i}"], res[f"aux{i}"] = s, q, aux
            res[f"range{i}"] = np.array([p0, p1])
        np.savez(os.path.join(out, f"rank{rank}.npz"), **res)
        be.close()
        NativeShardedCount.finalize(api.lib())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("native", [False, True], ids=["torch-exchange", "library-exchange"])
@pytest.mark.parametrize("k,b,trim,chunk", [(31, 24, False, 7000), (33, 22, True, 24000), (55, 26, False, 5000)])
def test_sharded_count_and_partitioned_correct_on_real_ranks(tmp_path, k, b, trim, chunk, native):
    import torch
    import torch.multiprocessing as mp
    world = 1
    while world * 2 <= min(torch.cuda.device_count(), 8):
        world *= 2
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    sys.path.insert(0, ROOT)
    from bfc_b200 import synth
    port = 29500 + (os.getpid() * 7 + k + (1000 if native else 0)) % 2000
    mp.spawn(_worker, args=(world, port, k, b, trim, chunk, str(tmp_path), native), nprocs=world, join=True)
    genome = synth.make_genome(80000, k + b, 0.2)
    seq, qual = synth.make_reads(genome, 24000, 120, k + b)
    s, q, off = synth.concat_batch(seq, qual)
    o = orc.OracleRun(orc.make_opt(k=k, bf_shift=b, filter_mode=1 if trim else 0))
    try:
        o.count(s, q, off)
        bloom = o.bloom_bytes()
        shard = len(bloom) // world
        if trim:
            keep, ts, te = o.trim(s, off)
        else:
            so, qo, ao, _ = o.correct(s, q, off)
        nk = npass = 0
        for r in range(world):
            got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
            assert np.array_equal(got["bloom"], bloom[r * shard:(r + 1) * shard])
            nk, npass = nk + int(got["n_kmers"]), npass + int(got["n_pass"])
            i = 0
            while f"range{i}" in got:
                p0, p1 = (int(v) for v in got[f"range{i}"])
                if trim:
                    assert np.array_equal(got[f"keep{i}"], keep[p0:p1]) and np.array_equal(got[f"ts{i}"], ts[p0:p1])
                    assert np.array_equal(got[f"te{i}"], te[p0:p1])
                else:
                    a, e = int(off[p0]), int(off[p1])
                    assert np.array_equal(got[f"seq{i}"], so[a:e]) and np.array_equal(got[f"qual{i}"], qo[a:e])
                    assert np.array_equal(got[f"aux{i}"], ao[2 * p0:2 * p1])
                i += 1
        assert nk == int(o.stats[0]) and npass == int(o.stats[1])
    finally:
        o.close()
