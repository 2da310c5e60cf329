#!/usr/bin/env python
"""Accuracy / timing probe of the tcgen05 dense chain (dense.cu mlp_chain_run): per-layer and whole-stack error against
fp64, next to torch's own fp32 matmul, for the wide DeepFM deep branch.  Run under ncu for the per-kernel times.

    python tools/mlp_chain_probe.py [--rows 65536] [--dims 624,400,400,400,1]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torecsys_b200 import ops  # noqa: E402


def nerr(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs() / (b.abs() + b.abs().mean())).max().item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows', type=int, default=65536)
    ap.add_argument('--dims', default='624,400,400,400,1')
    args = ap.parse_args()
    dims = [int(v) for v in args.dims.split(',')]
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device('cuda', 0)
    x = torch.rand(args.rows, dims[0], device=dev) * 2 - 1
    ws = [(torch.rand(dims[i + 1], dims[i], device=dev) * 2 - 1) * dims[i] ** -0.5 for i in range(len(dims) - 1)]
    bs = [torch.rand(dims[i + 1], device=dev) - 0.5 for i in range(len(dims) - 1)]
    # single layers, no activation in the way: the last layer of a one-layer pack has none
    for i in range(len(dims) - 1):
        xi = torch.rand(args.rows, dims[i], device=dev) * 2 - 1
        pack = ops.MlpPack([ws[i]], [bs[i]], ops.activation_id('relu'))
        got = ops.mlp(xi, pack)
        ref64 = xi.double() @ ws[i].double().t() + bs[i].double()
        ref32 = torch.addmm(bs[i], xi, ws[i].t())
        print(f'layer {dims[i]}->{dims[i + 1]}: ours {nerr(got, ref64):.3e}  torch fp32 {nerr(ref32, ref64):.3e}')
    pack = ops.MlpPack(ws, bs, ops.activation_id('relu'))
    got = ops.mlp(x, pack)
    h64, h32 = x.double(), x
    for i in range(len(ws)):
        h64 = h64 @ ws[i].double().t() + bs[i].double()
        h32 = torch.addmm(bs[i], h32, ws[i].t())
        if i < len(ws) - 1:
            h64, h32 = h64.relu(), h32.relu()
    print(f'stack: ours {nerr(got, h64):.3e}  torch fp32 {nerr(h32, h64):.3e}')
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()
