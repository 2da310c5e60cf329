#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py -x -q -p no:cacheprovider -k "deepfm" > gpurun_out/r2_tests_pw.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r2_tests_pw.log)"
grep -E "^FAILED|Error" gpurun_out/r2_tests_pw.log | head -5
python - <<'PY'
import sys, torch
sys.path.insert(0, '/root/repo')
from torecsys_b200 import ops
dev, N, B, rpf = 'cuda', 39, 65536, 5_128_192
dims = [N * 16, 400, 400, 400, 1]
ws = [torch.randn(dims[i + 1], dims[i], device=dev) * dims[i] ** -0.5 for i in range(4)]
bs = [torch.randn(dims[i + 1], device=dev) * 0.1 for i in range(4)]
pack = ops.MlpPack(ws, bs, ops.activation_id('relu'))
rows = N * rpf
w16 = torch.randn(rows, 16, device=dev); w1 = torch.randn(rows, 1, device=dev)
pk = ops.fm_pack_table(w16, w1)
off = (torch.arange(N) * rpf).to(dev)
ring = [torch.randint(0, rpf, (B, N), device=dev) for _ in range(4)]
out = torch.empty(B, 1, device=dev)
def t(f):
    for i in range(3): f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10): f(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1e3
print('split  %.1f us' % t(lambda i: ops.deepfm(ring[i % 4], off, w1, w16, pack, out=out)))
print('packed %.1f us' % t(lambda i: ops.deepfm_packed(ring[i % 4], off, pk, pack, out=out, kernel='auto')))
a = ops.deepfm(ring[0], off, w1, w16, pack); b2 = ops.deepfm_packed(ring[0], off, pk, pack, kernel='auto')
print('max rel diff', float(((a - b2).abs() / (a.abs() + 1)).max()))
PY
