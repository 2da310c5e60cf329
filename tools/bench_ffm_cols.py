"""One rank's share of the embedding-dimension-sharded configs[4] on ONE GPU: the interleaved FFM kernel on `cols`
columns of all 39 tables (25.6 M row ids), batch = the rank's part of the global 262 144.
    python tools/bench_ffm_cols.py [--cols 4] [--batch 131072]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from torecsys_b200 import ops
    ap = argparse.ArgumentParser()
    ap.add_argument('--cols', type=int, default=4)
    ap.add_argument('--batch', type=int, default=131072)
    args = ap.parse_args()
    n, rpf = 39, 657472
    rows = n * rpf
    dev = torch.device('cuda', 0)
    pitch = int(ops._cabi.load().trs_ffm_interleaved_pitch(n, args.cols))
    packed = torch.empty(rows, pitch, device=dev).uniform_(-0.01, 0.01)
    off = (torch.arange(n, dtype=torch.int64) * rpf).to(dev)
    idx = [torch.randint(0, rpf, (args.batch, n), device=dev) for _ in range(4)]
    bias = torch.zeros(1, device=dev)
    out = torch.empty(args.batch, 1, device=dev)
    ops.set_index_check('deferred')
    for i in range(3):
        ops.ffm_model_interleaved(idx[i % 4], off, packed, n, args.cols, bias, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        ops.ffm_model_interleaved(idx[i % 4], off, packed, n, args.cols, bias, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    chunk = n * args.cols * 4 + 4
    print(json.dumps({'cols': args.cols, 'batch': args.batch, 'ms': ms, 'samples_per_s': args.batch / ms * 1e3,
                      'useful_gbs': args.batch * n * chunk / ms / 1e6, 'chunk_bytes': chunk}))


if __name__ == '__main__':
    main()
