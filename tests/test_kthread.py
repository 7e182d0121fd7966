"""CPU: the two host schedulers of the phase drivers (csrc/kthread.c; interface and ordering rules of the reference's
kthread.c:85-146): kt_for covers every index exactly once; kt_pipeline never runs a step concurrently with itself,
hands batches to every step in order, and stops when step 0 returns NULL."""
import ctypes as C
import threading

import pytest

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


@pytest.mark.parametrize("n_steps", [2, 3, 4])
def test_kt_pipeline_order_and_exclusion(n_steps):
    """Batches pass every step in order, no step runs concurrently with itself, and never more than n_threads batches
    are in flight -- what the drivers' rings of pinned buffers (count_host.c, correct_host.c: N_FLAT) rely on."""
    L = C.CDLL(bfc_b200.lib_path())
    FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)
    n_batches = 25
    lock = threading.Lock()
    in_flight, max_in_flight = [0], [0]
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
                with lock:
                    in_flight[0] += 1
                    max_in_flight[0] = max(max_in_flight[0], in_flight[0])
            else:
                b = data
            if s == n_steps - 1:
                with lock:
                    in_flight[0] -= 1
            seen[s].append(b)
            for _ in range(2000):    # some work, so that the steps of neighbouring batches overlap
                pass
            return b if s < n_steps - 1 else None
        finally:
            with lock:
                running[s] -= 1

    cb = FN(step)
    L.kt_pipeline.argtypes = [C.c_int, FN, C.c_void_p, C.c_int]
    for threads in (1, 2, 3, 4):
        made[0] = 0
        in_flight[0] = max_in_flight[0] = 0
        for s in range(n_steps):
            seen[s].clear()
        L.kt_pipeline(threads, cb, None, n_steps)
        assert not errors
        assert in_flight[0] == 0 and 1 <= max_in_flight[0] <= threads, (threads, max_in_flight[0])
        for s in range(n_steps):
            assert seen[s] == list(range(1, n_batches + 1)), (threads, s)
