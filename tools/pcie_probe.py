"""Host<->device copy bandwidth through the library's own pinned allocator (what bench.py's e2e leg uses)."""
import ctypes as C, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bfc_b200
L = bfc_b200.lib()
n = 1 << 30
d = L.bfcg_dev_alloc(n); h = L.bfcg_host_alloc_pinned(n)
C.memset(h, 1, n)
for name, f, a, b in (("h2d", L.bfcg_h2d, d, h), ("d2h", L.bfcg_d2h, h, d)):
    f(a, b, n)
    t = time.perf_counter()
    for _ in range(4):
        f(a, b, n)
    dt = (time.perf_counter() - t) / 4
    print(f"{name}: {n / dt / 1e9:.1f} GB/s")
