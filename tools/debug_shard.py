import sys, os, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bfc_b200, orc
from bfc_b200 import synth, api
from bfc_b200.dist import CudaBackend, piece_bounds, owner_bits
os.environ["BFC_B200_SUBBATCH"] = str(1 << 17)
world, k, b = 4, 31, 22
genome = synth.make_genome(60000, k + world, 0.2)
seq, qual = synth.make_reads(genome, 16000, 120, k + world, 0.01)
seq, qual, off = synth.concat_batch(seq, qual)
N = len(off) - 1
oopt = orc.make_opt(k=k, bf_shift=b)
L = orc.lib()
o = orc.OracleRun(oopt); o.count(seq, qual, off)
print("oracle", o.stats)
e = bfc_b200.Engine(bfc_b200.make_opt(k=k, bf_shift=b)); e.count(seq, qual, off)
print("gpu unsharded", e.stats.n_kmers, e.stats.n_pass)
# oracle records
bt = orc.Batch(N, orc.as_u64p(off), orc.as_u8p(seq), orc.as_u8p(qual))
n = int(L.orc_enum_records(C.byref(oopt), C.byref(bt), None, None))
oy0, oy1 = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
L.orc_enum_records(C.byref(oopt), C.byref(bt), orc.as_u64p(oy0), orc.as_u64p(oy1))
for w in (1, 4):
    be = CudaBackend(bfc_b200.make_opt(k=k, bf_shift=b), w)
    y0, y1, counts = be.enum_records(api.host_batch(seq, qual, off), w)
    g0, g1 = y0[:n].cpu().numpy().view(np.uint64), y1[:n].cpu().numpy().view(np.uint64)
    x = b - 9
    hs = np.array([L.orc_hash_from_y(k, int(a) & ~(1 << 63), int(c)) for a, c in zip(oy0, oy1)], dtype=np.uint64)
    owner = ((hs & np.uint64((1 << x) - 1)) >> np.uint64(x - owner_bits(w))).astype(np.int64) if w > 1 else np.zeros(n, dtype=np.int64)
    order = np.argsort(owner, kind="stable")
    print("world", w, "counts", counts, "sum", sum(counts), "n", n, "bucket y0 equal", np.array_equal(g0, oy0[order]), "y1 equal", np.array_equal(g1, oy1[order]))
    if not np.array_equal(g0, oy0[order]):
        bad = np.nonzero(g0 != oy0[order])[0]
        print("  first mismatches", bad[:10], len(bad))
    be.close()

def run(w, chunk, sub):
    os.environ["BFC_B200_SUBBATCH"] = str(sub)
    ranks = [CudaBackend(bfc_b200.make_opt(k=k, bf_shift=b), w) for _ in range(w)]
    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        sent = []
        for r in range(w):
            p0, p1 = piece_bounds(lo, hi, r, w)
            s, q, f = seq[int(off[p0]):int(off[p1])], qual[int(off[p0]):int(off[p1])], off[p0:p1 + 1] - off[p0]
            y0, y1, counts = ranks[r].enum_records(api.host_batch(s, q, f), w)
            st = np.concatenate([[0], np.cumsum(counts)])
            sent.append([(y0[int(st[d]):int(st[d + 1])].clone(), y1[int(st[d]):int(st[d + 1])].clone()) for d in range(w)])
        for d in range(w):
            r0 = torch.cat([sent[r][d][0] for r in range(w)]); r1 = torch.cat([sent[r][d][1] for r in range(w)])
            ranks[d].count_records(r0, r1, int(r0.numel()), w)
    bloom = np.concatenate([ranks[r].bf_shard().cpu().numpy() for r in range(w)])
    print(f"world {w} chunk {chunk} sub {sub}: n_pass {sum(int(r.stats.n_pass) for r in ranks)} pending {sum(int(r.stats.n_pending) for r in ranks)} conflict {sum(int(r.stats.n_conflict) for r in ranks)} bloom_equal {np.array_equal(bloom, o.bloom_bytes())}")
    for r in ranks: r.close()

for w, chunk, sub in ((1, 16000, 1 << 26), (1, 16000, 1 << 17), (1, 5000, 1 << 26), (4, 16000, 1 << 26), (4, 5000, 1 << 26), (4, 5000, 1 << 17)):
    run(w, chunk, sub)

for sub in (1 << 26, 1 << 20, 1 << 19, 1 << 18):
    os.environ["BFC_B200_SUBBATCH"] = str(sub)
    for rep in range(2):
        e2 = bfc_b200.Engine(bfc_b200.make_opt(k=k, bf_shift=b)); e2.count(seq, qual, off)
        print("unsharded sub", sub, "n_pass", e2.stats.n_pass, "pending", e2.stats.n_pending, "conflict", e2.stats.n_conflict)
        e2.close()
