#!/bin/bash
# cross_tc5.cu: tiles in flight per CTA (slots; 1 and 2 also allow several CTAs per SM) -- the time does not scale with
# either, i.e. the chain is bound by the tensor pipe's per-instruction cost, not by latency.
python - <<'PY'
import os, subprocess, sys
for n in (1, 2, 3, 4):
    env = dict(os.environ, TRS_TC5_SLOTS=str(n))
    code = ("import sys, torch; sys.path.insert(0, '.'); from torecsys_b200 import ops; "
            "x = torch.randn(32768 * 39, 32, device='cuda'); w = torch.randn(6, 32, 32, device='cuda') * 0.17; "
            "b = torch.randn(6, 32, device='cuda') * 0.1; "
            "[ops.cross(x, w, b, tc5=True) for _ in range(3)]; torch.cuda.synchronize(); "
            "e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record(); "
            "[ops.cross(x, w, b, tc5=True) for _ in range(10)]; e1.record(); torch.cuda.synchronize(); "
            "t5 = e0.elapsed_time(e1) / 10; e0.record(); [ops.cross(x, w, b) for _ in range(10)]; e1.record(); "
            "torch.cuda.synchronize(); print('slots', %d, 'tcgen05 chain us', round(t5 * 1e3, 1), "
            "'mma.sync chain us', round(e0.elapsed_time(e1) / 10 * 1e3, 1))" % n)
    subprocess.run([sys.executable, '-c', code], env=env)
PY
