"""CPU: a model of the replay rule of k_count_part (csrc/count_part.cu) against the sequential cascade it must equal
(reference bbf.c:25-45 + count.c:54-70 with -t1).  Rounds of T records; a record whose bits are all set passes and
writes nothing; of the others the earliest of every block applies test-then-set at once; the ones that lost a claim
are replayed afterwards in stream order.  The model runs that rule literally (numpy, tiny filter = many collisions)
and compares, occurrence by occurrence, with one-at-a-time processing -- the argument in DESIGN.md section 4, executable."""
import numpy as np
import pytest


def probes(h, n_hashes):
    """bit positions inside a 512-bit block for hash h (bbf.c:27-42: positions < 8 belong to the lock byte)"""
    h1, h2 = int(h >> 10) & 511, int(h >> 30) & 511
    if (h2 & 31) == 0:
        h2 = (h2 + 1) & 511
    out, z = [], h1
    while len(out) < n_hashes:
        if z >= 8:
            out.append(z)
        z = (z + h2) & 511
    return out


def sequential(blocks, hashes, n_blocks, H):
    bits = np.zeros((n_blocks, 512), dtype=bool)
    passed = []
    for blk, h in zip(blocks, hashes):
        p = probes(h, H)
        c = 0
        for z in p:
            c += bool(bits[blk, z])
            bits[blk, z] = True
        passed.append(c == H)
    return bits, np.array(passed)


def rounds(blocks, hashes, n_blocks, H, T):
    bits = np.zeros((n_blocks, 512), dtype=bool)
    passed = np.zeros(len(blocks), dtype=bool)
    for base in range(0, len(blocks), T):
        idx = list(range(base, min(len(blocks), base + T)))
        pr = {i: probes(hashes[i], H) for i in idx}
        pend = []
        for i in idx:                               # phase 1: against the state before the round, nothing written
            if all(bits[blocks[i], z] for z in pr[i]):
                passed[i] = True
            else:
                pend.append(i)
        claim = {}
        for i in pend:                              # the claim: lowest thread per block
            claim.setdefault(blocks[i], i)
        losers = []
        for i in pend:                              # winners, all "at once" (different blocks: order is irrelevant)
            if claim[blocks[i]] == i:
                c = 0
                for z in pr[i]:
                    c += bool(bits[blocks[i], z])
                    bits[blocks[i], z] = True
                passed[i] = c == H
            else:
                losers.append(i)
        for i in losers:                            # the replay: stream order
            c = 0
            for z in pr[i]:
                c += bool(bits[blocks[i], z])
                bits[blocks[i], z] = True
            passed[i] = c == H
    return bits, passed


@pytest.mark.parametrize("n_blocks,n_keys,H,T", [(4, 40, 4, 128), (16, 300, 4, 128), (64, 2000, 7, 32), (2, 10, 4, 256)])
def test_round_replay_equals_sequential(n_blocks, n_keys, H, T):
    rng = np.random.default_rng(n_blocks * 1000 + n_keys)
    keys_h = rng.integers(0, 1 << 62, size=n_keys, dtype=np.uint64)
    keys_b = rng.integers(0, n_blocks, size=n_keys)
    occ = rng.integers(0, n_keys, size=6000)        # every key occurs many times, in random order
    blocks, hashes = keys_b[occ], keys_h[occ]
    b1, p1 = sequential(blocks, hashes, n_blocks, H)
    b2, p2 = rounds(blocks, hashes, n_blocks, H, T)
    assert np.array_equal(b1, b2) and np.array_equal(p1, p2)
    assert 0 < p1.sum() < len(p1)
