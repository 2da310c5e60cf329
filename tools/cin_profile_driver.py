import sys, torch
sys.path.insert(0, '/root/repo')
from torecsys_b200 import ops
dev = 'cuda'
N, B = 39, int(sys.argv[1]) if len(sys.argv) > 1 else 65536
x = torch.randn(B, N, 16, device=dev)
sizes = [128, 128]
cw, sc, sh = [], [], []
hp = N
for h in sizes:
    cw.append(torch.randn(2 * h, N * hp, device=dev) * (N * hp) ** -0.5); sc.append(torch.rand(2 * h, device=dev) + 0.5); sh.append(torch.randn(2 * h, device=dev) * 0.1); hp = h
fw = torch.randn(1, 256, device=dev); fb = torch.randn(1, device=dev)
pack = ops.CinPack(cw, sc, sh, sizes, False, ops.activation_id('relu'), fw, fb)
for _ in range(2): ops.cin(x, pack, 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): ops.cin(x, pack, 1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
fl = 2 * 16 * (N * N * 256 + N * 128 * 128)
print(f'B={B} {ms:.3f} ms  {B/ms/1e3:.2f} Msamp/s  {fl*B/ms/1e9:.1f} algorithmic TFLOP/s')
