"""One GPU, `world` VIRTUAL ranks: all interleaved shards of configs[4] live in this GPU's HBM, so the chunk copies of
trs_ffm_shard_blocks are all local.  Times the block kernel of virtual rank 0 at the full global batch: what the kernel
STRUCTURE sustains when NVLink is out of the picture (compare with bench.py --only-sharded on the real box).
    python tools/bench_ffm_blocks_local.py [--world 8] [--batch 262144] [--rows-per-field 657472]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from torecsys_b200 import ops
    ap = argparse.ArgumentParser()
    ap.add_argument('--world', type=int, default=8)
    ap.add_argument('--batch', type=int, default=262144)
    ap.add_argument('--rows-per-field', type=int, default=657472)
    args = ap.parse_args()
    n, e, w = 39, 16, args.world
    rows = n * args.rows_per_field
    slots = (n + w - 1) // w
    dev = torch.device('cuda', 0)
    shards = [torch.empty(rows, slots, e, device=dev).uniform_(-0.01, 0.01) for _ in range(w)]
    rows_all = torch.randint(0, rows, (args.batch, n), dtype=torch.int32, device=dev)
    first = torch.zeros(args.batch // w, device=dev)
    out = torch.empty(args.batch, device=dev)
    res = {}
    for k in (0, w - 1):
        plan = ops.FfmShardPlan(n, w, k, e)
        ptrs = [s.data_ptr() for s in shards]
        for _ in range(3):
            ops.ffm_shard_blocks(rows_all, plan, ptrs, first, (0, args.batch // w), out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.ffm_shard_blocks(rows_all, plan, ptrs, first, (0, args.batch // w), out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        total = (plan.tx_bytes[0] + plan.tx_bytes[1]) / 2 * args.batch
        res[f'rank{k}'] = {'ms': ms, 'samples_per_s': args.batch / ms * 1e3, 'chunk_gbs': total / ms / 1e6,
                           'remote_equiv_gbs': (plan.remote_bytes(0) + plan.remote_bytes(1)) / 2 * args.batch / ms / 1e6,
                           'copies': plan.n_copies, 'stage_bytes': plan.stage_bytes}
    print(json.dumps({'world': w, 'batch': args.batch, 'all_local': res}))


if __name__ == '__main__':
    main()
