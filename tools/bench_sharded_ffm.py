#!/usr/bin/env python
"""BASELINE.json configs[4]: FieldAwareFactorizationMachine, 39 fields, tables row-sharded over the GPUs of one box.

    python -m torch.distributed.run --nproc-per-node N tools/bench_sharded_ffm.py [--rows-per-field R] [--batch B]

Each rank owns ~39/N whole tables (table-wise sharding), keeps batch/N samples and runs the fused FFM kernel with
peer-mapped table pointers: looked-up vectors cross NVLink as 128-bit peer loads inside the kernel (no collective).
Prints one JSON line (rank 0): whole-job samples/s, device-timed, max over ranks, plus the NVLink volume per GPU.
Default sizes: 39 x 657 472 rows per table x 39 tables x 64 B = 64 GB total (the 1 B-row configuration); use
--rows-per-field to scale down on fewer GPUs.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N, E = 39, 16


def main():
    from torecsys_b200 import ops
    from torecsys_b200.sharded import ShardedFFM, ShardedFieldAwareTables
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows-per-field', type=int, default=657_472)
    ap.add_argument('--batch', type=int, default=262_144, help='global batch')
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--scheme', default='both', choices=['sample', 'owner', 'both'],
                    help='sample: pairs computed at the rank of the sample (peer loads only); owner: pairs reduced at a '
                         'table owner (all-gather of indices + reduce-scatter of logits, half the NVLink volume)')
    args = ap.parse_args()
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    fs = [args.rows_per_field] * N
    rows = sum(fs)
    tables = ShardedFieldAwareTables(E, fs)
    gen = torch.Generator(device=dev).manual_seed(rank)
    tables.init_(lambda t, dst: dst.uniform_(-0.05, 0.05, generator=gen))      # xavier-like scale
    w_feat = torch.randn(rows, 1, device=dev)
    bias = torch.rand(1, device=dev)
    model = ShardedFFM(tables, w_feat, bias)
    per = args.batch // world
    igen = torch.Generator(device=dev).manual_seed(1234 + rank)
    ring = [torch.randint(0, args.rows_per_field, (per, N), device=dev, generator=igen) for _ in range(4)]
    def measure(fn):
        for i in range(args.warmup):
            fn(ring[i % 4])
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            fn(ring[i % 4])
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ops.check_index_errors()
        return float(t.item()) / args.steps

    schemes = ['sample', 'owner'] if args.scheme == 'both' else [args.scheme]
    for scheme in schemes:
        ms = measure(model.forward if scheme == 'sample' else model.forward_owner_side)
        if rank == 0:
            if scheme == 'sample':
                nv_rows = per * N * (N - 1) * tables.plan.remote_fraction()
            else:
                nv_rows = per * world * max(tables.plan.remote_rows_per_sample(r) for r in range(world))
            nv_bytes = nv_rows * 64
            print(json.dumps({
                'metric': 'ctr_forward_samples_per_sec', 'value': per * world / (ms * 1e-3), 'unit': 'samples/s',
                'n_gpus': world, 'ms_per_step': ms, 'scaling': 'strong (global batch fixed)', 'scheme': scheme,
                'config': {'workload': f'configs[4]: FFM {N} fields, {N} tables x {rows} rows x {E} '
                                       f'(={N * rows * E * 4 / 1e9:.1f} GB), global batch {per * world}, tables sharded '
                                       f'table-wise over {world} GPUs',
                           'collectives': 'none (in-kernel peer loads)' if scheme == 'sample' else
                                          'NCCL all-gather of indices + reduce-scatter of partial logits, in-kernel peer loads'},
                'nvlink_bytes_in_per_gpu_per_step': nv_bytes,
                'nvlink_GBps_per_gpu': nv_bytes / (ms * 1e-3) / 1e9,
                'hbm_algorithmic_GBps_per_gpu': per * 95320 / (ms * 1e-3) / 1e9}), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
