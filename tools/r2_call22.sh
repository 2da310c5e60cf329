#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -x -q -p no:cacheprovider > gpurun_out/r2_tests_train.log 2>&1
echo "training tests rc=$? $(tail -1 gpurun_out/r2_tests_train.log)"
grep -E "^FAILED|^ERROR|Error|assert " gpurun_out/r2_tests_train.log | head -12
python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
from torecsys_b200 import ops
B = 65536
dims = [624, 16, 16, 16, 1]
ws = [torch.randn(dims[i + 1], dims[i], device='cuda') * dims[i] ** -0.5 for i in range(4)]
bs = [torch.randn(dims[i + 1], device='cuda') * 0.1 for i in range(4)]
pack = ops.MlpPack(ws, bs, ops.activation_id('relu'))
x = torch.randn(B, 624, device='cuda'); g = torch.randn(B, 1, device='cuda')
for _ in range(3): ops.mlp_backward(x, pack, g)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.mlp_backward(x, pack, g)
e1.record(); torch.cuda.synchronize()
t_k = e0.elapsed_time(e1) / 10
xr = x.clone().requires_grad_(True); wr = [w.clone().requires_grad_(True) for w in ws]; br = [b.clone().requires_grad_(True) for b in bs]
def f():
    h = xr
    for i in range(4):
        h = torch.nn.functional.linear(h, wr[i], br[i])
        if i < 3: h = torch.relu(h)
    return torch.autograd.grad(h, [xr] + wr + br, g)
torch.backends.cuda.matmul.allow_tf32 = False
for _ in range(3): f()
torch.cuda.synchronize(); e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
print(f'mlp backward 65536 x [624,16,16,16,1]: kernel {t_k*1e3:.1f} us, torch recompute (fwd+bwd, fp32 cuBLAS) {e0.elapsed_time(e1)/10*1e3:.1f} us')
PY
