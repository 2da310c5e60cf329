"""Host-side index narrowing: how the sessions' narrowing pool (csrc/session.cu, AVX-512 / AVX2 + streaming stores) scales
over the host's cores on one 65 536 x 39 batch (trs_host_narrow_pool_ns: best of 40 passes)."""
import ctypes, sys
sys.path.insert(0, '/root/repo')
import torch
from torecsys_b200 import _cabi
lib = _cabi.load()
n = 65536 * 39
src = torch.randint(0, 5_000_000, (n,), dtype=torch.int64).pin_memory()
dst = torch.empty(n, dtype=torch.int32).pin_memory()
for th in (1, 2, 4, 6, 8, 10, 12, 14, 16):
    ns = lib.trs_host_narrow_pool_ns(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()), n, th, 40)
    print(f'{th:2d} threads: {ns / 1e3:7.1f} us per batch  ({n * 12 / ns:6.1f} GB/s read+write, {65536 / ns * 1e3:6.1f} M samples/s)', flush=True)
