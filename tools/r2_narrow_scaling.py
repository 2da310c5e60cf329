"""Host-side index narrowing: how does the AVX-512 form scale over threads on this box (memory-bound or not)?
Python threads call trs_host_narrow_indices on slices (ctypes drops the GIL)."""
import ctypes, sys, threading, time
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
from torecsys_b200 import _cabi
lib = _cabi.load()
n = 65536 * 39
ring = [torch.randint(0, 5_000_000, (n,), dtype=torch.int64).pin_memory() for _ in range(8)]
dst = torch.empty(n, dtype=torch.int32).pin_memory()
def run(threads, reps=30):
    per = (n // threads + 15) // 16 * 16
    def work(t, src):
        lo, hi = t * per, min(n, (t + 1) * per)
        if lo < hi:
            lib.trs_host_narrow_indices(ctypes.c_void_p(src.data_ptr() + lo * 8), ctypes.c_void_p(dst.data_ptr() + lo * 4), hi - lo, 0)
    best = 1e9
    for r in range(reps):
        src = ring[r % 8]
        ts = [threading.Thread(target=work, args=(t, src)) for t in range(threads)]
        t0 = time.perf_counter()
        for t in ts: t.start()
        for t in ts: t.join()
        best = min(best, time.perf_counter() - t0)
    return best
for th in (1, 2, 4, 8, 12, 14, 16):
    dt = run(th)
    print(f'{th:2d} threads: {dt * 1e6:7.1f} us per 65536 x 39 batch  ({n * 12 / dt / 1e9:6.1f} GB/s read+write)', flush=True)
