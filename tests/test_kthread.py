"""CPU: the two host schedulers of the phase drivers (csrc/kthread.c; interface and ordering rules of the reference's
kthread.c:85-146): kt_for covers every index exactly once; kt_pipeline never runs a step concurrently with itself,
hands batches to every step in order, and stops when step 0 returns NULL."""
import ctypes as C
import threading

import bfc_b200


def test_kt_for_covers_every_index_once():
    L = C.CDLL(bfc_b200.lib_path())
    FN = C.CFUNCTYPE(None, C.c_void_p, C.c_long, C.c_int)
    n = 1000
    hits, tids, lock = [0] * n, set(), threading.Lock()

    def work(_data, i, tid):
        with lock:
            hits[i] += 1
            tids.add(tid)

    cb = FN(work)
    L.kt_for.argtypes = [C.c_int, FN, C.c_void_p, C.c_long]
    for threads in (1, 4):
        for i in range(n):
            hits[i] = 0
        tids.clear()
        L.kt_for(threads, cb, None, n)
        assert hits == [1] * n
        assert tids <= set(range(threads))


def test_kt_pipeline_order_and_exclusion():
    L = C.CDLL(bfc_b200.lib_path())
    FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)
    n_batches, n_steps = 25, 3
    lock = threading.Lock()
    running = [0] * n_steps
    seen = [[] for _ in range(n_steps)]
    made = [0]
    errors = []

    def step(_shared, s, data):
        with lock:
            running[s] += 1
            if running[s] > 1:
                errors.append(("step runs concurrently with itself", s))
        try:
            if s == 0:
                if made[0] == n_batches:
                    return None
                made[0] += 1
                b = made[0]          # batch ids 1..n (a non-NULL pointer value)
            else:
                b = data
            seen[s].append(b)
            for _ in range(2000):    # some work, so that the steps of neighbouring batches overlap
                pass
            return b if s < n_steps - 1 else None
        finally:
            with lock:
                running[s] -= 1

    cb = FN(step)
    L.kt_pipeline.argtypes = [C.c_int, FN, C.c_void_p, C.c_int]
    for threads in (1, 2, 3):
        made[0] = 0
        for s in range(n_steps):
            seen[s].clear()
        L.kt_pipeline(threads, cb, None, n_steps)
        assert not errors
        for s in range(n_steps):
            assert seen[s] == list(range(1, n_batches + 1)), (threads, s)
