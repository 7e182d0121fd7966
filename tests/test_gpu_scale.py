"""GPU (B200), larger than the oracle can follow: size-independent cross-checks.  The two count paths are
independent implementations of the same exact order semantics (count_part.cu: stable partition + Bloom slices
replayed in shared memory; count.cu: probe / resolve / replay against the filter in HBM) -- on 12 M synthetic reads,
several count windows each, they must leave the same filter bytes and the same table; and the correction must give
the same bytes whether the reads sit in host memory (many windows, overlapped copies) or are resident in HBM."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, L_READ, K, B = 12_000_000, 150, 33, 33


def test_count_paths_and_batch_kinds_agree_at_scale(monkeypatch):
    import torch
    import bfc_b200
    from bfc_b200 import api
    L = api.lib()
    assert L.bfcg_device_count() > 0
    RB = L_READ + 1
    G = N * L_READ // 30
    nb = N * RB
    d_gen, d_seq, d_qual, d_off = L.bfcg_dev_alloc(G), L.bfcg_dev_alloc(nb), L.bfcg_dev_alloc(nb), L.bfcg_dev_alloc(8 * (N + 1))
    assert d_gen and d_seq and d_qual and d_off
    assert L.bfcg_synth_genome(d_gen, G, 5) == 0
    assert L.bfcg_synth_reads(d_gen, G, 5, 0, N, L_READ, 0.01, 2e-4, d_seq, d_qual, d_off) == 0
    b = api.Batch()
    b.n_reads, b.n_bytes, b.where = N, nb, api.DEVICE
    b.off, b.seq, b.qual = C.cast(d_off, api.u64p), C.cast(d_seq, api.u8p), C.cast(d_qual, api.u8p)
    monkeypatch.setenv("BFC_B200_COUNT_WINDOW", str(1 << 29))   # 4 count windows
    monkeypatch.setenv("BFC_B200_EC_BATCH", str(1 << 28))       # 7 correction windows
    engines, blooms, tables = {}, {}, {}
    try:
        for path in ("part", "probe"):
            monkeypatch.setenv("BFC_B200_COUNT", path)
            e = engines[path] = bfc_b200.Engine(bfc_b200.make_opt(k=K, bf_shift=B))
            e.count_batch(b)
            blooms[path] = e.bloom_bytes()
            n = int(L.bfcg_ch_export_device(e.ch, None, None))
            sub = torch.empty(n, dtype=torch.int32, device="cuda")
            key = torch.empty(n, dtype=torch.int64, device="cuda")
            assert int(L.bfcg_ch_export_device(e.ch, C.c_void_p(sub.data_ptr()), C.c_void_p(key.data_ptr()))) == n
            order = torch.argsort(key)  # (the export order depends on the table layout; keys are unique per sub-table...
            k2, s2 = key[order], sub[order]
            order2 = torch.argsort(s2, stable=True)  # ... so sort by (sub, key))
            tables[path] = (s2[order2].cpu().numpy(), k2[order2].cpu().numpy())
            del sub, key, order, k2, s2, order2
        assert int(engines["part"].stats.n_kmers) == int(engines["probe"].stats.n_kmers) > 0
        assert int(engines["part"].stats.n_pass) == int(engines["probe"].stats.n_pass)
        assert np.array_equal(blooms["part"], blooms["probe"])
        assert np.array_equal(tables["part"][0], tables["probe"][0]) and np.array_equal(tables["part"][1], tables["probe"][1])
        engines["probe"].close()
        del engines["probe"]
        # correction: a slice of the reads through host batches vs resident in HBM
        e = engines["part"]
        m = 3_000_000
        hs, hq = np.empty(m * RB, dtype=np.uint8), np.empty(m * RB, dtype=np.uint8)
        assert L.bfcg_d2h(hs.ctypes.data, d_seq, m * RB) == 0 and L.bfcg_d2h(hq.ctypes.data, d_qual, m * RB) == 0
        off = np.arange(m + 1, dtype=np.uint64) * np.uint64(RB)
        s1, q1, a1 = e.correct(hs, hq, off)
        d_aux = L.bfcg_dev_alloc(8 * m)
        bd = api.Batch()
        bd.n_reads, bd.n_bytes, bd.where = m, m * RB, api.DEVICE
        bd.off, bd.seq, bd.qual = C.cast(d_off, api.u64p), C.cast(d_seq, api.u8p), C.cast(d_qual, api.u8p)
        e.correct_batch(bd, d_aux)
        s2, q2, a2 = np.empty_like(hs), np.empty_like(hq), np.empty(2 * m, dtype=np.uint32)
        assert L.bfcg_d2h(s2.ctypes.data, d_seq, m * RB) == 0 and L.bfcg_d2h(q2.ctypes.data, d_qual, m * RB) == 0
        assert L.bfcg_d2h(a2.ctypes.data, d_aux, 8 * m) == 0
        L.bfcg_dev_free(d_aux)
        assert np.array_equal(a1, a2) and np.array_equal(s1, s2) and np.array_equal(q1, q2)
        assert int((s1 != hs).sum()) > 0  # something was corrected
    finally:
        for e in engines.values():
            e.close()
        for p in (d_gen, d_seq, d_qual, d_off):
            L.bfcg_dev_free(p)


def test_eight_shards_at_full_filter_size():
    """Round 1's N = 8 failure, reproduced on one GPU: a shard of 8 receives all its keys in 1/8 of the table's
    sub-table regions (the owner bits are sub-table index bits), so a table sized for uniform use of every region
    overflowed at the first 16 M-read chunk of the bench (k = 33, Bloom 2^37 bits).  Here the 8 ranks of one such chunk
    are run one after the other on one GPU through the same calls bench.py makes (bfcg_enum_records -> exchange ->
    bfcg_count_record_runs on 1/8 filters and shard tables); the union of the shard tables must equal the table of
    the single-GPU count of the same reads, and the shard filters its filter."""
    import torch
    import bfc_b200
    from bfc_b200 import api
    from bfc_b200.dist import CudaBackend, piece_bounds
    L = api.lib()
    world, n, k, b = 8, 16_000_000, 33, 37
    RB = L_READ + 1
    G = n * L_READ // 30
    nb = n * RB
    d_gen, d_seq, d_qual, d_off = L.bfcg_dev_alloc(G), L.bfcg_dev_alloc(nb), L.bfcg_dev_alloc(nb), L.bfcg_dev_alloc(8 * (n + 1))
    assert d_gen and d_seq and d_qual and d_off
    assert L.bfcg_synth_genome(d_gen, G, 11) == 0
    assert L.bfcg_synth_reads(d_gen, G, 11, 0, n, L_READ, 0.01, 2e-4, d_seq, d_qual, d_off) == 0
    opt = bfc_b200.make_opt(k=k, bf_shift=b)
    ranks = [CudaBackend(opt, world, 0, rank=r) for r in range(world)]
    single = None
    try:
        sent = []
        for r in range(world):
            p0, p1 = piece_bounds(0, n, r, world)
            pb = api.Batch()
            pb.n_reads, pb.n_bytes, pb.where = p1 - p0, (p1 - p0) * RB, api.DEVICE
            pb.off = C.cast(d_off, api.u64p)
            pb.seq, pb.qual = C.cast(d_seq + p0 * RB, api.u8p), C.cast(d_qual + p0 * RB, api.u8p)
            y0, y1, counts = ranks[r].enum_records(pb, world)
            starts = np.concatenate([[0], np.cumsum(counts)])
            sent.append([(y0[int(starts[d]):int(starts[d + 1])].clone(), y1[int(starts[d]):int(starts[d + 1])].clone()) for d in range(world)])
            ranks[r]._buf.clear()
        for d in range(world):
            r0 = torch.cat([sent[r][d][0] for r in range(world)])
            r1 = torch.cat([sent[r][d][1] for r in range(world)])
            ranks[d].count_record_runs(r0, r1, [int(sent[r][d][0].numel()) for r in range(world)], world)
            del r0, r1
        del sent
        parts = [ranks[r].export_table() for r in range(world)]
        sizes = [int(s.numel()) for s, _ in parts]
        assert min(sizes) > 0.8 * max(sizes) > 1_000_000       # every shard took its share of the keys ...
        cap = [int(L.bfcg_ch_capacity_log2(ranks[r].ch)) for r in range(world)]
        assert max(cap) <= 29, cap                              # ... in a table of its own size, not the whole job's
        key = torch.cat([kk for _, kk in parts]); sub = torch.cat([ss for ss, _ in parts])
        order = torch.argsort(key); key, sub = key[order], sub[order]
        order = torch.argsort(sub, stable=True); key, sub = key[order].cpu().numpy(), sub[order].cpu().numpy()
        bloom = torch.cat([ranks[r].bf_shard() for r in range(world)])
        n_k, n_p = sum(int(r.stats.n_kmers) for r in ranks), sum(int(r.stats.n_pass) for r in ranks)
        for r in ranks:
            r.close()
        ranks = []
        del parts, order
        single = bfc_b200.Engine(opt)
        fb = api.Batch()
        fb.n_reads, fb.n_bytes, fb.where = n, nb, api.DEVICE
        fb.off, fb.seq, fb.qual = C.cast(d_off, api.u64p), C.cast(d_seq, api.u8p), C.cast(d_qual, api.u8p)
        single.count_batch(fb)
        assert int(single.stats.n_kmers) == n_k and int(single.stats.n_pass) == n_p
        m = int(L.bfcg_ch_export_device(single.ch, None, None))
        s1, k1 = torch.empty(m, dtype=torch.int32, device="cuda"), torch.empty(m, dtype=torch.int64, device="cuda")
        assert int(L.bfcg_ch_export_device(single.ch, C.c_void_p(s1.data_ptr()), C.c_void_p(k1.data_ptr()))) == m == len(key)
        order = torch.argsort(k1); k1, s1 = k1[order], s1[order]
        order = torch.argsort(s1, stable=True); k1, s1 = k1[order].cpu().numpy(), s1[order].cpu().numpy()
        assert np.array_equal(s1, sub) and np.array_equal(k1, key)
        full = torch.empty(1 << (b - 3), dtype=torch.uint8, device="cuda")
        cudart = C.CDLL("libcudart.so.12")
        assert cudart.cudaMemcpy(C.c_void_p(full.data_ptr()), C.c_void_p(single.bf.contents.b), C.c_size_t(full.numel()), 3) == 0
        assert bool(torch.equal(full, bloom))
    finally:
        for r in ranks:
            r.close()
        if single:
            single.close()
        for p in (d_gen, d_seq, d_qual, d_off):
            L.bfcg_dev_free(p)
