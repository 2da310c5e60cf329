#!/usr/bin/env python
"""Per-role timeline of CTA 0 of deepfm_tc5_kernel (debug hook trs_debug_tc5_trace): where does a stage spend its time?"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import BATCH, EMBED, NUM_FIELDS, make_mlp_params  # noqa: E402
from torecsys_b200 import _cabi, ops  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
rpf = int(sys.argv[2]) if len(sys.argv) > 2 else 5_128_192
dev = torch.device('cuda', 0)
gen = torch.Generator().manual_seed(0)
ws, bs = make_mlp_params(torch, gen, dev)
pack = ops.MlpPack(ws, bs, ops.activation_id('relu'))
rows = NUM_FIELDS * rpf
packed = torch.randn(rows, 32, device=dev)
offsets = (torch.arange(NUM_FIELDS, dtype=torch.int64) * rpf).to(dev)
idx = [torch.randint(0, rpf, (BATCH, NUM_FIELDS), dtype=torch.int64).to(dev) for _ in range(3)]
ops.set_index_check('deferred')
lib = _cabi.load()
trace = torch.zeros(7 * 512 * 4, dtype=torch.int64, device=dev)
for i in range(3):
    ops.deepfm_packed(idx[i], offsets, packed, pack, kernel='tc5', variant=variant)
torch.cuda.synchronize()
lib.trs_debug_tc5_trace(trace.data_ptr())
ops.deepfm_packed(idx[0], offsets, packed, pack, kernel='tc5', variant=variant)
torch.cuda.synchronize()
lib.trs_debug_tc5_trace(None)
t = trace.cpu().view(7, 512, 4)
t0 = int(t[6, 0, 0])
rel = lambda v: (int(v) - t0) if int(v) else -1
print(f'variant {variant}: kernel entry 0, setup done {rel(t[6,0,1])}, exit {rel(t[6,0,2])} (cycles)')
print('tile: index-warp convert start/end, epilogue acc_full/stored')
for tl in range(6):
    print(tl, rel(t[4, tl, 0]), rel(t[4, tl, 1]), '|', rel(t[5, tl, 0]), rel(t[5, tl, 1]))
print('q: producer[slot free, issued] consumer[v_full seen, lo_full arrived] mma[lo_full seen, hi inputs seen, hi issued] slab[slot free]')
nq = 0
for q in range(512):
    if int(t[0, q, 0]) == 0:
        break
    nq = q + 1
    if q < 30 or q % 13 == 0 or q > nq - 3:
        print(q, [rel(t[0, q, 0]), rel(t[0, q, 1])], [rel(t[1, q, 0]), rel(t[1, q, 1])],
              [rel(t[2, q, 0]), rel(t[2, q, 1]), rel(t[2, q, 2])], [rel(t[3, q, 0])])
print('stages', nq)
