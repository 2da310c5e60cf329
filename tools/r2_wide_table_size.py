"""DeepFM-[400,400,400] (the fused wide deep branch) against the table size: is the gathering layer bound by DRAM or by
its producers?  usage: python tools/r2_wide_table_size.py"""
import sys, torch
sys.path.insert(0, '/root/repo')
from torecsys_b200 import ops
dev, N, B = 'cuda', 39, 65536
dims = [N * 16, 400, 400, 400, 1]
ws = [torch.randn(dims[i + 1], dims[i], device=dev) * dims[i] ** -0.5 for i in range(4)]
bs = [torch.randn(dims[i + 1], device=dev) * 0.1 for i in range(4)]
pack = ops.MlpPack(ws, bs, ops.activation_id('relu'))
for rpf in (1000, 100_000, 5_128_192):
    rows = N * rpf
    w16 = torch.randn(rows, 16, device=dev)
    w1 = torch.randn(rows, 1, device=dev)
    off = (torch.arange(N) * rpf).to(dev)
    ring = [torch.randint(0, rpf, (B, N), device=dev) for _ in range(4)]
    out = torch.empty(B, 1, device=dev)
    for i in range(3): ops.deepfm(ring[i % 4], off, w1, w16, pack, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10): ops.deepfm(ring[i % 4], off, w1, w16, pack, out=out)
    e1.record(); torch.cuda.synchronize()
    print(f'rows/field {rpf}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per {B} samples', flush=True)
    del w16, w1
